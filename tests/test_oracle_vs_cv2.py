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


def _scene(seed, n=400, stereo=True):
    """a frame pose, a KITTI-like camera and landmarks in front of / behind / beside it"""
    rng = np.random.default_rng(seed)
    a, b, c = rng.normal(0, 0.2, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    Rcw = (Rx @ Ry @ Rz).astype(np.float32)
    tcw = rng.normal(0, 2, 3).astype(np.float32)
    Ow = (-cv2.gemm(np.ascontiguousarray(Rcw.T), tcw.reshape(3, 1), 1.0, None, 0.0)).reshape(3)      # Frame: mOw = -mRcw.t()*mtcw
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.22], [0, 0, 1]], np.float32)
    lms = np.zeros(n, O.LM_DTYPE)
    Pc = np.stack([rng.uniform(-30, 30, n), rng.uniform(-10, 10, n), rng.uniform(-5, 60, n)], 1)
    lms["Pw"] = ((Pc - tcw) @ Rcw).astype(np.float32)         # Rcw^T (Pc - tcw)
    lms["size"] = rng.uniform(0.05, 1.5, n).astype(np.float32)
    lms["min_dist"] = rng.uniform(0.5, 10, n).astype(np.float32)
    lms["max_dist"] = rng.uniform(20, 80, n).astype(np.float32)
    lms["assoc_idx"] = -1
    return Rcw, tcw, Ow.astype(np.float32), K, 386.1448, stereo, (0.0, 1241.0, 0.0, 376.0), lms


def _project_cv2(Rcw, tcw, K, mbf, stereo, bounds, P):
    """Frame::ProjectLandMark + Camera::Project written with the cv2 calls the reference's cv::Mat expressions turn into"""
    Pc = cv2.gemm(Rcw, P.reshape(3, 1).astype(np.float32), 1.0, tcw.reshape(3, 1), 1.0)
    PcZ = np.float32(Pc[2, 0])
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = np.float32(1.0) / PcZ
        Pch = np.array([[Pc[0, 0] / PcZ], [Pc[1, 0] / PcZ], [Pc[2, 0] / PcZ]], np.float32)
        uv = cv2.gemm(K, Pch, 1.0, None, 0.0)
        u, v = np.float32(uv[0, 0]), np.float32(uv[1, 0])
        ur = np.float32(u - np.float32(np.float32(mbf) * invz)) if stereo else np.float32(-1.0)
    valid = bool(PcZ > 0 and bounds[0] <= u <= bounds[1] and bounds[2] <= v <= bounds[3])
    return u, v, ur, valid


@pytest.mark.parametrize("seed,stereo", [(0, True), (1, False), (2, True)])
def test_landmark_projection_matches_cv2_expressions(seed, stereo):
    """pins the fp32 policy of orc_project_landmarks (gemm small-matrix path, norm in double) against cv2.gemm / cv2.norm"""
    Rcw, tcw, Ow, K, mbf, st, bounds, lms = _scene(seed, stereo=stereo)
    kps = np.zeros(5, O.KP_DTYPE); kps["size"] = [31, 37, 44, 53, 64]
    lms["assoc_idx"][::9] = np.arange(len(lms[::9])) % 5
    pr = O.make_projection(Rcw, tcw, Ow, K, mbf, st, bounds)
    th = 3.0
    q, passed = O.project_landmarks(pr, lms, kps, th)
    f = np.float32
    for i, lm in enumerate(lms):
        u, v, ur, valid = _project_cv2(Rcw, tcw, K, mbf, st, bounds, lm["Pw"])
        PO = (lm["Pw"] - Ow).astype(np.float32)
        dist = f(cv2.norm(PO.reshape(3, 1)))
        ok = valid and not (dist < lm["min_dist"] or dist > lm["max_dist"])
        if lm["assoc_idx"] >= 0:
            size_px = f(kps["size"][lm["assoc_idx"]])
        else:
            half = f(lm["size"] / f(2))
            Lp = lm["Pw"].copy(); Lp[0] = f(Lp[0] - half)
            Rp = lm["Pw"].copy(); Rp[0] = f(Rp[0] + half)
            size_px = f(_project_cv2(Rcw, tcw, K, mbf, st, bounds, Rp)[0] - _project_cv2(Rcw, tcw, K, mbf, st, bounds, Lp)[0])
        radius = f(f(f(th) * size_px) / f(31.0))
        want = (u, v, radius, f(f(0.5) * size_px), f(f(1.5) * size_px), ur, radius if st else f(-1.0))
        got = tuple(q[i][k] for k in ("u", "v", "r", "size_lo", "size_hi", "ur", "ur_radius"))
        same = all((np.isnan(a) and np.isnan(b)) or a == b for a, b in zip(got, want))
        assert same and bool(passed[i]) == ok, (i, got, want, passed[i], ok)
    assert 0.1 < passed.mean() < 0.9


@pytest.mark.parametrize("shape", [(37, 53), (36, 52), (376, 1241), (480, 752)])
@pytest.mark.parametrize("cn", [1, 3, 4])
def test_preprocess_matches_cv2(shape, cn):
    """ImageProcessing::PreProcessImg (ImageProcessing.cpp:118-138) restated: cv::resize at fscale 1.0 / 0.5 then cvtColor to gray."""
    h, w = shape
    rng = np.random.default_rng(h * 7 + cn)
    img = rng.integers(0, 256, (h, w, cn), dtype=np.uint8) if cn > 1 else rng.integers(0, 256, (h, w), dtype=np.uint8)
    for rgb in (True, False):
        for half in (False, True):
            ref = cv2.resize(img, None, fx=0.5, fy=0.5) if half else img
            if cn == 3:
                ref = cv2.cvtColor(ref, cv2.COLOR_RGB2GRAY if rgb else cv2.COLOR_BGR2GRAY)
            elif cn == 4:
                ref = cv2.cvtColor(ref, cv2.COLOR_RGBA2GRAY if rgb else cv2.COLOR_BGRA2GRAY)
            assert np.array_equal(O.preprocess(img, rgb, half), ref), (shape, cn, rgb, half)


def test_preprocess_rejects_sizes_outside_the_box_path():
    with pytest.raises(ValueError):
        O.preprocess(np.zeros((35, 51, 3), np.uint8), True, True)       # cvRound(17.5) = 18, 2*18 > 35
