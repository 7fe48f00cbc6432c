"""Pins oracle/cvshim -- the OpenCV algorithms underneath oracle/_ref (the reference's own translation units) -- bit for
bit against the real library (cv2): cv::FAST, cv::resize, cv::GaussianBlur, cv::copyMakeBorder, cv::fastAtan2, the
small-matrix gemm policy and cv::norm.  CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402
from hyslam_b200 import synth  # noqa: E402

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and reference tree absent")
cv2.setNumThreads(1)


def _img(shape, kind, seed):
    return (synth.noise_image if kind == "noise" else synth.blocks_image)(shape[0], shape[1], seed)


@pytest.mark.parametrize("shape,kind,seed", [((480, 752), "noise", 0), ((480, 752), "blocks", 1), ((42, 36), "noise", 2), ((36, 36), "blocks", 9),
                                             ((7, 7), "noise", 3), ((6, 30), "noise", 4), ((36, 71), "noise", 5), ((33, 32), "noise", 6),
                                             ((40, 39), "noise", 7), ((376, 1241), "noise", 8)])
def test_fast_matches_cv2(shape, kind, seed):
    img = _img(shape, kind, seed)
    fd = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = np.array([[k.pt[0], k.pt[1], k.response] for k in fd.detect(img)], np.float32).reshape(-1, 3)
    x, y, r = R.shim_fast(img)
    assert np.array_equal(ref, np.stack([x, y, r], 1))


def test_fast_on_cell_views_matches_cv2():
    """the reference calls cv::FAST on ~36x36 views of a level (ORBExtractor.cpp:450): strided ROIs, vector tails"""
    img = synth.noise_image(200, 300, 11)
    fd = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    rng = np.random.default_rng(0)
    for _ in range(60):
        w, h = int(rng.integers(7, 80)), int(rng.integers(7, 60))
        x0, y0 = int(rng.integers(0, 300 - w)), int(rng.integers(0, 200 - h))
        view = img[y0:y0 + h, x0:x0 + w]
        ref = np.array([[k.pt[0], k.pt[1], k.response] for k in fd.detect(np.ascontiguousarray(view))], np.float32).reshape(-1, 3)
        n = R.lib().cvshim_fast  # strided call: pass the parent's stride
        cap = w * h
        xs, ys, rs = (np.empty(cap, np.float32) for _ in range(3))
        cnt = n(view.ctypes.data_as(R.C.c_void_p), w, h, img.strides[0], 20, 1, xs.ctypes.data_as(R.C.c_void_p), ys.ctypes.data_as(R.C.c_void_p),
                rs.ctypes.data_as(R.C.c_void_p), cap)
        assert cnt == len(ref)
        assert np.array_equal(ref, np.stack([xs[:cnt], ys[:cnt], rs[:cnt]], 1))


def test_fast_thresholds_and_no_nms():
    img = synth.noise_image(120, 160, 8)
    for th in (7, 20, 40):
        fd = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        ref = np.array([[k.pt[0], k.pt[1], k.response] for k in fd.detect(img)], np.float32).reshape(-1, 3)
        x, y, r = R.shim_fast(img, threshold=th)
        assert np.array_equal(ref, np.stack([x, y, r], 1))
    fd = cv2.FastFeatureDetector_create(20, False, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = np.array([[k.pt[0], k.pt[1]] for k in fd.detect(img)], np.float32).reshape(-1, 2)
    x, y, _ = R.shim_fast(img, nms=False)        # without NMS cv::FAST reports no response; positions only
    assert np.array_equal(ref, np.stack([x, y], 1))


@pytest.mark.parametrize("shape,seed", [((480, 752), 0), ((376, 1241), 1), ((97, 131), 2), ((240, 320), 3)])
def test_resize_chain_matches_cv2(shape, seed):
    img = synth.noise_image(shape[0], shape[1], seed)
    cur = img
    for (w, h) in O.level_sizes(O.default_params(), shape[1], shape[0])[1:]:
        if w < 8 or h < 8:
            break
        b = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(R.shim_resize(cur, w, h), b)
        cur = b


@pytest.mark.parametrize("sw,sh,dw,dh", [(100, 80, 50, 40), (64, 64, 63, 31), (33, 47, 40, 60), (200, 10, 77, 9), (17, 19, 17, 19), (50, 50, 120, 7)])
def test_resize_odd_sizes(sw, sh, dw, dh):
    img = synth.noise_image(sh, sw, sw * 1000 + sh)
    assert np.array_equal(R.shim_resize(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape,kind,seed", [((480, 752), "noise", 0), ((134, 210), "noise", 4), ((100, 130), "blocks", 3), ((9, 40), "noise", 5), ((40, 7), "blocks", 6)])
def test_gaussian7_matches_cv2(shape, kind, seed):
    img = _img(shape, kind, seed)
    assert np.array_equal(R.shim_blur(img), cv2.GaussianBlur(img.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


def test_copy_make_border_matches_cv2():
    img = synth.noise_image(40, 57, 1)
    for bt in (cv2.BORDER_REFLECT_101, cv2.BORDER_REPLICATE):
        assert np.array_equal(R.shim_border(img, 19, bt), cv2.copyMakeBorder(img, 19, 19, 19, 19, bt))


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    ys = rng.integers(-300000, 300000, 20000)
    xs = rng.integers(-300000, 300000, 20000)
    ys[:50] = 0
    xs[25:75] = 0
    for yv, xv in zip(ys, xs):
        assert np.float32(R.shim_fast_atan2(float(yv), float(xv))) == np.float32(cv2.fastAtan2(float(yv), float(xv)))


def test_gemm_policy_matches_cv2():
    """A*B, A*B + C (fused, as cv::MatExpr evaluates `Rcw*P + tcw`), A.t()*B for the shapes the path multiplies"""
    rng = np.random.default_rng(3)
    for _ in range(300):
        for (m, k, n) in ((3, 3, 1), (3, 3, 3), (4, 4, 4), (4, 4, 1), (3, 4, 1), (2, 3, 3)):
            A = rng.normal(0, 3, (m, k)).astype(np.float32)
            B = rng.normal(0, 50, (k, n)).astype(np.float32)
            Cm = rng.normal(0, 5, (m, n)).astype(np.float32)
            assert np.array_equal(R.shim_gemm(A, B).view(np.uint32), cv2.gemm(A, B, 1.0, None, 0.0).view(np.uint32)), (m, k, n)
            assert np.array_equal(R.shim_gemm(A, B, Cm).view(np.uint32), cv2.gemm(A, B, 1.0, Cm, 1.0).view(np.uint32)), (m, k, n)
        A = rng.normal(0, 1, (3, 3)).astype(np.float32)
        b = rng.normal(0, 5, (3, 1)).astype(np.float32)
        assert np.array_equal(R.shim_gemm(A, b, ta=True).view(np.uint32), cv2.gemm(A, b, 1.0, None, 0.0, flags=cv2.GEMM_1_T).view(np.uint32))


def test_norm_matches_cv2():
    rng = np.random.default_rng(4)
    for _ in range(500):
        v = rng.normal(0, 30, 3).astype(np.float32)
        assert R.shim_norm(v) == cv2.norm(v.reshape(3, 1))
