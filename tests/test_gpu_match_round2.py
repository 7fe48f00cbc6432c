"""GPU parity of the matcher entry points added in round 2 -- Fuse, SearchBySim3 (per direction) and SearchForInitialization -- against
the C oracle, and, where oracle/_ref travelled to this box, against the REFERENCE'S OWN FeatureMatcher code end to end."""
import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import _ffi as F
from oracle import oracle as O
from oracle import ref as R

import test_oracle_match_vs_ref as T          # scene builders shared with the CPU pin

pytestmark = pytest.mark.gpu
HAVE_REF = R.available()


@pytest.fixture(scope="module")
def frames():
    return _frames()


def _frames():
    from hyslam_b200 import synth
    out = []
    p = O.default_params(2000)
    sp = O.StereoParams(T.MBF, T.FX, 376, 100.0, 50.0, 31.0)
    for seed in (40, 41):
        L, Rt = synth.stereo_pair(376, 1241, seed)
        kl, dl = O.extract(L, p)
        kr, dr = O.extract(Rt, p)
        uR, depth, _, _ = O.stereo_match(sp, kl, dl, kr, dr)
        out.append((kl, dl, uR, depth))
    return out


@pytest.mark.parametrize("seed,window,ratio", [(0, 100, 0.9), (1, 40, 0.9), (2, 100, 0.6), (3, 15, 0.9)])
def test_search_for_initialization(frames, seed, window, ratio):
    k1, d1, _, _ = frames[0]
    rng = np.random.default_rng(seed)
    n2 = 1700
    src = rng.integers(0, len(k1), n2)
    k2 = k1[src].copy()
    k2["x"] += rng.normal(0, 12, n2).astype(np.float32); k2["y"] += rng.normal(0, 8, n2).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.choice([0.0, 0.0, 0.0, 45.0], n2) + rng.normal(0, 2, n2)).astype(np.float32) % np.float32(360)
    d2 = T.noisy_desc(rng, d1[src], 20)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    m = hb.FeatureMatcher()
    n, m12, pm = m.SearchForInitialization(k1, d1, k2, d2, T.BOUNDS, prev, window, 50.0, ratio)
    on, om12, opm = O.search_for_initialization(k1, d1, k2, d2, O.Bounds(*T.BOUNDS), prev, window, 50.0, ratio)
    assert n == on and n > 50
    assert np.array_equal(m12, om12) and pm.tobytes() == opm.tobytes()
    if HAVE_REF:
        sc = R.Scene(1)
        f1 = sc.add_frame(k1, d1, T.K, np.eye(4), T.BOUNDS)
        f2 = sc.add_frame(k2, d2, T.K, np.eye(4), T.BOUNDS)
        rn, rm12, rpm = sc.search_for_initialization(f1, f2, prev, window, R.settings(nnratio=ratio, th_low=50.0))
        assert n == rn and np.array_equal(m12, rm12) and pm.tobytes() == rpm.tobytes()
        sc.close()


def test_search_for_initialization_edge_cases():
    m = hb.FeatureMatcher()
    k = np.zeros(3, F.KP_DTYPE); k["x"] = [10, 20, 30]; k["y"] = 10
    d = np.zeros((3, 32), np.uint8)
    n, m12, pm = m.SearchForInitialization(k, d, k[:0], d[:0], (0, 100, 0, 100), np.zeros((3, 2), np.float32), 10)
    assert n == 0 and m12.tolist() == [-1, -1, -1]
    # three identical features fight for one target: the FIRST keeps it (a later one needs a strictly smaller distance)
    t = k[:1].copy(); td = d[:1].copy()
    prev = np.tile(np.array([[10.0, 10.0]], np.float32), (3, 1))
    n, m12, pm = m.SearchForInitialization(k, d, t, td, (0, 100, 0, 100), prev, 50, 50.0, 0.9)
    on, om12, _ = O.search_for_initialization(k, d, t, td, O.Bounds(0, 100, 0, 100), prev, 50, 50.0, 0.9)
    assert n == on and np.array_equal(m12, om12) and m12.tolist() == [0, -1, -1]


@pytest.mark.parametrize("seed,stereo", [(0, True), (1, False), (2, True)])
def test_fuse(frames, seed, stereo):
    kk, dk, uRk, depk = frames[seed % 2]
    rng = np.random.default_rng(seed)
    Rcw, tcw, Tm = T.pose(rng)
    n = 1500
    pick, Pw, size, raw_min, raw_max = T.landmarks_around(rng, kk, uRk, Rcw, tcw, n, stereo)
    lm_desc = T.noisy_desc(rng, dk[pick])
    Ow_true = -(Rcw.T.astype(np.float64) @ tcw.astype(np.float64))
    normal = Pw.astype(np.float64) - Ow_true
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    normal = normal + rng.normal(0, 0.9, (n, 3)) * (rng.random((n, 1)) < 0.5)
    normal = (normal / np.linalg.norm(normal, axis=1, keepdims=True)).astype(np.float32)
    Ow = (-(Rcw.T.astype(np.float64) @ tcw.astype(np.float64))).astype(np.float32)
    sc = None
    if HAVE_REF:
        sc = R.Scene(n)
        sc.add_mappoints(Pw, lm_desc, normal=normal, size=size, min_dist=raw_min, max_dist=raw_max)
        kf = sc.add_frame(kk, dk, T.K, Tm, T.BOUNDS, mbf=T.MBF, stereo=stereo, uR=uRk if stereo else None, depth=depk if stereo else None, keyframe=True)
        Ow = sc.camera_center(kf)
    lms = np.zeros(n, F.LM_DTYPE)
    lms["Pw"] = Pw; lms["size"] = size; lms["min_dist"] = np.float32(0.8) * raw_min; lms["max_dist"] = np.float32(1.2) * raw_max; lms["assoc_idx"] = -1
    m = hb.FeatureMatcher()
    pr = m.make_projection(Rcw, tcw, Ow, T.K, T.MBF, stereo, T.BOUNDS)
    bi, b, s, acc, passed = m.Fuse(pr, lms, normal, lm_desc, kk, dk, 3.0, 5.99, t_uR=uRk if stereo else None, thr=50.0, ratio=1.0)
    opr = O.make_projection(Rcw, tcw, Ow, T.K, T.MBF, stereo, T.BOUNDS)
    q, opassed = O.project_landmarks(opr, lms, kk, 3.0)
    q["ur_radius"] = -1
    opassed &= O.viewing_angle(Ow, Pw, normal, 1.047)
    bounds = O.Bounds(*T.BOUNDS)
    off, idx = O.grid_build(kk, bounds)
    obi, ob, osd, oacc = O.match_window_ex(kk, dk, uRk if stereo else None, None, bounds, off, idx, q, lm_desc, 50.0, 1.0, rule=0, q_active=opassed, reproj_thr=5.99)
    assert np.array_equal(passed, opassed) and 0.2 < passed.mean() < 0.9
    for g, w, name in zip((bi, b, s, acc), (obi, ob, osd, oacc), ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, w), name
    assert acc.sum() > 30
    if sc is not None:
        gidx, glm = sc.fuse(kf, np.arange(n), 3.0, 5.99, R.settings(th_low=50.0))
        want = {}
        for j in range(n):
            if acc[j] and int(bi[j]) not in want:
                want[int(bi[j])] = j
        assert gidx.tolist() == sorted(want) and glm.tolist() == [want[k] for k in sorted(want)]
        sc.close()


@pytest.mark.parametrize("seed,s12", [(0, 1.0), (1, 1.07)])
def test_search_by_sim3_direction(frames, seed, s12):
    k1, d1, uR1, dep1 = frames[0]
    k2, d2, uR2, dep2 = frames[1]
    rng = np.random.default_rng(seed)
    R1, t1, T1 = T.pose(rng, 0.05, 0.5)
    R2, t2, T2 = T.pose(rng, 0.05, 0.5)
    n = 1200
    pick, Pw, size, raw_min, raw_max = T.landmarks_around(rng, k2, uR2, R2, t2, n, True)      # landmarks of KF1 that fall on KF2's features
    lm_desc = T.noisy_desc(rng, d2[pick])
    T12 = T1.astype(np.float64) @ np.linalg.inv(T2.astype(np.float64))
    R12 = T12[:3, :3].astype(np.float32); t12 = T12[:3, 3].astype(np.float32)
    f32 = np.float32
    sR21 = (f32(1.0 / s12) * R12.T).astype(f32)
    t21 = -np.array([(sR21[i, 0] * t12[0] + sR21[i, 1] * t12[1]) + sR21[i, 2] * t12[2] for i in range(3)], f32)
    Ow2 = (-(R2.T.astype(np.float64) @ t2.astype(np.float64))).astype(np.float32)
    lms = np.zeros(n, F.LM_DTYPE)
    lms["Pw"] = Pw; lms["size"] = size; lms["min_dist"] = np.float32(0.8) * raw_min * 0.3; lms["max_dist"] = np.float32(1.2) * raw_max * 2; lms["assoc_idx"] = -1
    lms["assoc_idx"][5::19] = pick[5::19]
    m = hb.FeatureMatcher()
    pr2 = m.make_projection(R2, t2, Ow2, T.K, T.MBF, True, T.BOUNDS)
    bi, b, acc, passed = m.SearchBySim3Direction(R1, t1, sR21, t21, pr2, lms, lm_desc, k2, d2, 7.5, thr=100.0)
    opr2 = O.make_projection(R2, t2, Ow2, T.K, T.MBF, True, T.BOUNDS)
    q, opassed = O.project_sim3(R1, t1, sR21, t21, opr2, lms, k2, 7.5)
    bounds = O.Bounds(*T.BOUNDS)
    off, idx = O.grid_build(k2, bounds)
    obi, ob, _, oacc = O.match_window_ex(k2, d2, None, None, bounds, off, idx, q, lm_desc, 100.0, np.inf, rule=0, q_active=opassed)
    assert np.array_equal(passed, opassed) and passed.mean() > 0.3
    assert np.array_equal(bi, obi) and np.array_equal(b, ob) and np.array_equal(acc, oacc)
    assert acc.sum() > 100
