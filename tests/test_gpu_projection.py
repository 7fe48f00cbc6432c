"""FeatureMatcher::SearchByProjection(Frame&, landmarks, th) on the device: landmark projection /
landmark criteria against the oracle (itself pinned against cv2.gemm / cv2.norm in tests/test_oracle_vs_cv2.py), and the
fused projection -> window -> view criteria -> best score call against the oracle's composition of the same stages."""
import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import _ffi as F, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _scene(seed, n, stereo, kps, uR=None):
    """landmarks scattered around back-projected keypoints of a real extraction (so that windows contain candidates),
    plus landmarks behind / beside the camera and outside their distance range"""
    rng = np.random.default_rng(seed)
    a, b, c = rng.normal(0, 0.1, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    Rcw = (Rx @ Ry @ Rz).astype(np.float32)
    tcw = rng.normal(0, 1, 3).astype(np.float32)
    Ow = (-(Rcw.T.astype(np.float64) @ tcw.astype(np.float64))).astype(np.float32)
    fx, cx, cy = 718.856, 607.19, 185.22
    K = np.array([[fx, 0, cx], [0, fx, cy], [0, 0, 1]], np.float32)
    pick = rng.integers(0, len(kps), n)
    z = rng.uniform(2, 40, n)
    if uR is not None:                      # depth consistent with the keypoint's own stereo match where it has one
        has = uR[pick] > 0
        z[has] = 386.1448 / np.maximum(kps["x"][pick][has] - uR[pick][has], 0.5)
    u = kps["x"][pick] + rng.normal(0, 3, n); v = kps["y"][pick] + rng.normal(0, 3, n)
    Pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fx, z], 1)
    Pc[::11, 2] *= -1                       # behind the camera
    Pc[5::13, 0] += 80                      # far outside the image
    lms = np.zeros(n, F.LM_DTYPE)
    lms["Pw"] = ((Pc - tcw.astype(np.float64)) @ Rcw.astype(np.float64)).astype(np.float32)
    lms["size"] = (kps["size"][pick] * z / fx * rng.uniform(0.6, 1.6, n)).astype(np.float32)
    lms["min_dist"] = rng.uniform(0.5, 6, n).astype(np.float32)
    lms["max_dist"] = rng.uniform(15, 120, n).astype(np.float32)
    lms["assoc_idx"] = -1
    lms["assoc_idx"][3::17] = pick[3::17]
    return (Rcw, tcw, Ow, K, 386.1448, stereo, (0.0, 1241.0, 0.0, 376.0)), lms, pick


@pytest.mark.parametrize("seed,stereo,th", [(0, True, 3.0), (1, False, 5.0), (2, True, 1.0)])
def test_search_by_projection_matches_oracle(seed, stereo, th):
    L, R = synth.stereo_pair(376, 1241, 40 + seed)
    p = O.default_params(2000)
    kl, dl = O.extract(L, p)
    kr, dr = O.extract(R, p)
    uR, _, _, _ = O.stereo_match(O.StereoParams(386.1448, 718.856, 376, 100.0, 50.0, 31.0), kl, dl, kr, dr)
    cam, lms, pick = _scene(seed, 1500, stereo, kl, uR if stereo else None)
    rng = np.random.default_rng(100 + seed)
    lm_desc = dl[pick].copy()
    flip = rng.integers(0, 256, (len(pick), 6))
    for i in range(len(pick)):                               # a few flipped bits: realistic distances, ties in best/second
        for bpos in flip[i][: rng.integers(0, 7)]:
            lm_desc[i, bpos >> 3] ^= 1 << (bpos & 7)
    t_matched = (rng.random(len(kl)) < 0.1).astype(np.uint8)

    m = hb.FeatureMatcher()
    pr = m.make_projection(*cam)
    opr = O.make_projection(*cam)
    q, passed = m.ProjectLandMarks(pr, lms, kl, th)
    oq, opassed = O.project_landmarks(opr, lms, kl, th)
    assert q.tobytes() == oq.tobytes(), "window queries (u, v, r, size bounds, ur) differ from the oracle"
    assert np.array_equal(passed, opassed)
    assert 0.3 < passed.mean() < 0.95

    bi, b, s, acc, passed2 = m.SearchByProjectionLandMarks(pr, lms, lm_desc, kl, dl, th, t_uR=uR if stereo else None, t_matched=t_matched,
                                                           thr=100.0, ratio=0.9)
    assert np.array_equal(passed2, opassed)
    bounds = O.Bounds(*cam[6])
    off, idx = O.grid_build(kl, bounds)
    obi, ob, osd, oacc = O.match_window(kl, dl, uR if stereo else None, t_matched, bounds, off, idx, oq, lm_desc, thr=100.0, ratio=0.9)
    dead = opassed == 0                                      # landmark criteria failed: never reaches the view criteria
    obi[dead] = -1; ob[dead] = 65535; osd[dead] = 65535; oacc[dead] = 0
    for g, w, name in zip((bi, b, s, acc), (obi, ob, osd, oacc), ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, w), name
    assert acc.sum() > 20, acc.sum()


def test_projection_edge_cases():
    m = hb.FeatureMatcher()
    cam = (np.eye(3), np.zeros(3), np.zeros(3), np.array([[500, 0, 320], [0, 500, 240], [0, 0, 1]]), 40.0, True, (0, 640, 0, 480))
    pr = m.make_projection(*cam)
    lms = np.zeros(3, F.LM_DTYPE)
    lms["Pw"] = [[0, 0, 0], [0, 0, 5], [0, 0, 5]]            # z = 0 (division by zero in the reference too), fine, fine
    lms["size"] = 0.2; lms["min_dist"] = [0, 1, 6]; lms["max_dist"] = [10, 10, 10]; lms["assoc_idx"] = -1
    kps = np.zeros(1, F.KP_DTYPE)
    q, passed = m.ProjectLandMarks(pr, lms, kps, 3.0)
    oq, opassed = O.project_landmarks(O.make_projection(*cam), lms, kps, 3.0)
    assert np.array_equal(passed, opassed) and passed.tolist() == [0, 1, 0]
    assert q[1:].tobytes() == oq[1:].tobytes()
    assert np.array_equal(np.isnan(q["u"]), np.isnan(oq["u"]))
    lms["assoc_idx"][1] = 7                                  # out of range index is reported
    with pytest.raises(hb.HyorbError):
        m.ProjectLandMarks(pr, lms, kps, 3.0)


@pytest.mark.parametrize("seed,stereo", [(3, False), (4, True)])
def test_motion_model_variant_with_rotation_consistency(seed, stereo):
    """SearchByProjection(CurrentFrame, LastFrame, th) (FeatureMatcher.cc:145-176): no DistanceCriterion, rotation histogram over the
    matches, duplicates on one keypoint resolved to the landmark listed last (the reference's map order)"""
    L, R = synth.stereo_pair(376, 1241, 60 + seed)
    p = O.default_params(2000)
    kl, dl = O.extract(L, p)
    kr, dr = O.extract(R, p)
    uR, _, _, _ = O.stereo_match(O.StereoParams(386.1448, 718.856, 376, 100.0, 50.0, 31.0), kl, dl, kr, dr)
    cam, lms, pick = _scene(seed, 2500, stereo, kl, uR if stereo else None)        # 2500 landmarks on 2000 keypoints: duplicates guaranteed
    rng = np.random.default_rng(200 + seed)
    lm_desc = dl[pick].copy()
    prev_angle = (kl["angle"][pick] + rng.choice([0.0, 0.0, 0.0, 25.0, 170.0], len(pick)) + rng.normal(0, 2, len(pick))).astype(np.float32) % np.float32(360)
    flags = F.SBP_STEREO | F.SBP_ROTATION
    m = hb.FeatureMatcher()
    pr = m.make_projection(*cam)
    bi, b, s, acc, passed = m.SearchByProjectionLandMarks(pr, lms, lm_desc, kl, dl, 7.0, t_uR=uR if stereo else None, thr=100.0, ratio=0.9,
                                                          flags=flags, lm_prev_angle=prev_angle)
    # oracle composition: projection without the distance criterion, window scan, then the rotation criterion
    opr = O.make_projection(*cam)
    far = lms.copy(); far["min_dist"] = 0; far["max_dist"] = np.float32(3e38)
    oq, opassed = O.project_landmarks(opr, far, kl, 7.0)
    if not stereo:
        pass
    bounds = O.Bounds(*cam[6])
    off, idx = O.grid_build(kl, bounds)
    obi, ob, osd, oacc = O.match_window(kl, dl, uR if stereo else None, None, bounds, off, idx, oq, lm_desc, thr=100.0, ratio=0.9)
    dead = opassed == 0
    obi[dead] = -1; ob[dead] = 65535; osd[dead] = 65535; oacc[dead] = 0
    before = int(oacc.sum())
    oacc = O.projection_rotation(obi, oacc, prev_angle, kl)
    assert np.array_equal(passed, opassed)
    for g, w, name in zip((bi, b, s, acc), (obi, ob, osd, oacc), ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, w), name
    kept = int(acc.sum())
    assert 50 < kept < before                              # the histogram and the de-duplication both removed something
    assert len(np.unique(bi[acc > 0])) == kept             # one landmark per keypoint survives
