"""Pins the C oracle's restatements of the OpenCV primitives the reference calls
(cv::resize, cv::GaussianBlur, cv::FAST, cv::fastAtan2 -- SURVEY.md A.7) bit-for-bit against the
real library (cv2).  CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as O  # noqa: E402
from hyslam_b200 import synth  # noqa: E402

cv2.setNumThreads(1)


@pytest.mark.parametrize("shape,seed", [((480, 752), 0), ((376, 1241), 1), ((97, 131), 2), ((240, 320), 3)])
def test_resize_chain_matches_cv2(shape, seed):
    img = synth.noise_image(shape[0], shape[1], seed)
    p = O.default_params()
    cur = img
    for (w, h) in O.level_sizes(p, shape[1], shape[0])[1:]:
        if w < 8 or h < 8:
            break
        a = O.resize_linear(cur, w, h)
        b = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(a, b)
        cur = b


@pytest.mark.parametrize("sw,sh,dw,dh", [(100, 80, 50, 40), (64, 64, 63, 31), (33, 47, 40, 60), (200, 10, 77, 9), (17, 19, 17, 19)])
def test_resize_odd_sizes(sw, sh, dw, dh):
    img = synth.noise_image(sh, sw, sw * 1000 + sh)
    assert np.array_equal(O.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape,kind,seed", [((480, 752), "noise", 0), ((134, 210), "noise", 4), ((100, 130), "blocks", 3),
                                             ((9, 40), "noise", 5), ((40, 7), "blocks", 6)])
def test_gaussian7_matches_cv2(shape, kind, seed):
    img = (synth.noise_image if kind == "noise" else synth.blocks_image)(shape[0], shape[1], seed)
    b = cv2.GaussianBlur(img.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(O.gaussian7(img), b)


@pytest.mark.parametrize("shape,kind,seed", [((480, 752), "noise", 0), ((480, 752), "blocks", 1), ((42, 36), "noise", 2),
                                             ((36, 36), "blocks", 9), ((7, 7), "noise", 3), ((6, 30), "noise", 4)])
def test_fast_matches_cv2(shape, kind, seed):
    img = (synth.noise_image if kind == "noise" else synth.blocks_image)(shape[0], shape[1], seed)
    fd = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ks = fd.detect(img)
    ref = np.array([[k.pt[0], k.pt[1], k.response] for k in ks], np.float32).reshape(-1, 3)
    x, y, r = O.fast9(img)
    assert np.array_equal(ref, np.stack([x, y, r], 1))


def test_fast_no_nms_matches_cv2():
    img = synth.noise_image(120, 160, 8)
    fd = cv2.FastFeatureDetector_create(20, False, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    # without NMS cv::FAST reports response 0; only positions are comparable (the path always uses NMS)
    ref = np.array([[k.pt[0], k.pt[1]] for k in fd.detect(img)], np.float32).reshape(-1, 2)
    x, y, r = O.fast9(img, nms=False)
    assert np.array_equal(ref, np.stack([x, y], 1))


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    ys = rng.integers(-300000, 300000, 30000)
    xs = rng.integers(-300000, 300000, 30000)
    ys[:50] = 0
    xs[25:75] = 0
    for yv, xv in zip(ys, xs):
        assert np.float32(O.fast_atan2(float(yv), float(xv))) == np.float32(cv2.fastAtan2(float(yv), float(xv)))


def test_umax_table():
    assert O.umax().tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
