"""The C++ drop-in matcher (include/hyorb_hyslam_matcher.hpp: HYSLAM::CudaFeatureMatcher, the public signatures of FeatureMatcher.h:105-176
over the C ABI) against HYSLAM::FeatureMatcher ITSELF: twin scenes of the reference's real Frame / KeyFrame / MapPoint objects (oracle/_ref),
one driven by the reference's matcher, one by the drop-in on the GPU; everything a hySLAM caller observes must be identical."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import ref as R
import test_oracle_match_vs_ref as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "cpp", "_build", "libmatcher_shim_test.so")
if not (R.available() and os.path.exists(SHIM)):
    pytest.skip("oracle/_ref or the matcher shim test library did not travel to this box", allow_module_level=True)


@pytest.fixture(scope="module")
def shim():
    R.lib()
    return C.CDLL(SHIM)


@pytest.fixture(scope="module")
def frames():
    import test_gpu_match_round2 as G
    return G._frames()


def twin(n, shim):
    return R.Scene(n), R.Scene(n, matcher_lib=shim, prefix="shimm_")


@pytest.mark.parametrize("seed,stereo,th,ratio", [(0, True, 3.0, 0.9), (1, False, 5.0, 0.9), (3, True, 7.0, 0.75)])
def test_search_by_projection_local_map(frames, shim, seed, stereo, th, ratio):
    kl, dl, uR, depth = frames[seed % 2]
    rng = np.random.default_rng(seed)
    Rcw, tcw, Tm = T.pose(rng)
    n = 1500
    pick, Pw, size, raw_min, raw_max = T.landmarks_around(rng, kl, uR, Rcw, tcw, n, stereo)
    lm_desc = T.noisy_desc(rng, dl[pick])
    assoc0 = np.full(len(kl), -1, np.int32)
    extra = np.nonzero(rng.random(len(kl)) < 0.08)[0]
    assoc0[extra] = n + np.arange(len(extra))
    for i in range(3, n, 17):
        if assoc0[pick[i]] < 0:
            assoc0[pick[i]] = i
    order = rng.permutation(n)                           # the caller's landmark list is in no particular order, with null entries and repeats
    lm_list = np.concatenate([order, [-1, -1], order[:50]])
    res = []
    for sc in twin(n + len(extra), shim):
        sc.add_mappoints(Pw, lm_desc, size=size, min_dist=raw_min, max_dist=raw_max)
        sc.add_mappoints(np.zeros((len(extra), 3), np.float32) + [0, 0, 5], dl[extra], size=np.full(len(extra), 0.1, np.float32))
        for j in range(len(extra)):
            sc.set_observation_count(n + j, 1)
        f = sc.add_frame(kl, dl, T.K, Tm, T.BOUNDS, mbf=T.MBF, stereo=stereo, uR=uR if stereo else None, depth=depth if stereo else None, assoc=assoc0)
        nm = sc.search_by_projection(f, lm_list, th, R.settings(nnratio=ratio))
        res.append((nm, sc.assoc(f, len(kl))))
        sc.close()
    assert res[0][0] == res[1][0] and res[0][0] > 10
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("seed,stereo", [(4, True), (5, False)])
def test_search_by_projection_motion_model(frames, shim, seed, stereo):
    kc, dc, uRc, depc = frames[0]
    kp, dp, uRp, depp = frames[1]
    rng = np.random.default_rng(seed)
    Rcw, tcw, Tm = T.pose(rng)
    n = 1200
    pick, Pw, size, raw_min, raw_max = T.landmarks_around(rng, kc, uRc, Rcw, tcw, n, stereo)
    lm_desc = T.noisy_desc(rng, dc[pick])
    prev_idx = rng.permutation(len(kp))[:n]
    kp = kp.copy()
    kp["angle"][prev_idx] = (kc["angle"][pick] + rng.choice([0.0, 0.0, 0.0, 25.0, 170.0], n) + rng.normal(0, 2, n)).astype(np.float32) % np.float32(360)
    assoc_prev = np.full(len(kp), -1, np.int32)
    assoc_prev[prev_idx] = np.arange(n)
    res = []
    for sc in twin(n, shim):
        sc.add_mappoints(Pw, lm_desc, size=size, min_dist=raw_min, max_dist=raw_max)
        cur = sc.add_frame(kc, dc, T.K, Tm, T.BOUNDS, mbf=T.MBF, stereo=stereo, uR=uRc if stereo else None, depth=depc if stereo else None)
        last = sc.add_frame(kp, dp, T.K, np.eye(4), T.BOUNDS, mbf=T.MBF, stereo=stereo, uR=uRp if stereo else None, depth=depp if stereo else None, assoc=assoc_prev)
        nm = sc.search_by_projection_motion(cur, last, 7.0, R.settings(nnratio=0.9), mono=not stereo)
        res.append((nm, sc.assoc(cur, len(kc))))
        sc.close()
    assert res[0][0] == res[1][0] and res[0][0] > 30
    assert np.array_equal(res[0][1], res[1][1])


def test_search_by_projection_relocalisation(frames, shim):
    kc, dc, uRc, depc = frames[0]
    kk, dk, uRk, depk = frames[1]
    rng = np.random.default_rng(9)
    Rcw, tcw, Tm = T.pose(rng)
    n = 1000
    pick, Pw, size, raw_min, raw_max = T.landmarks_around(rng, kc, uRc, Rcw, tcw, n, True)
    lm_desc = T.noisy_desc(rng, dc[pick])
    assoc_kf = np.full(len(kk), -1, np.int32)
    assoc_kf[rng.permutation(len(kk))[:n]] = np.arange(n)
    found = np.arange(0, n, 7)
    res = []
    for sc in twin(n, shim):
        sc.add_mappoints(Pw, lm_desc, size=size, min_dist=raw_min, max_dist=raw_max)
        cur = sc.add_frame(kc, dc, T.K, Tm, T.BOUNDS, mbf=T.MBF, stereo=True, uR=uRc, depth=depc)
        kf = sc.add_frame(kk, dk, T.K, np.eye(4), T.BOUNDS, mbf=T.MBF, stereo=True, uR=uRk, depth=depk, assoc=assoc_kf, keyframe=True)
        nm = sc.search_by_projection_reloc(cur, kf, found, 10.0, 64, R.settings())
        res.append((nm, sc.assoc(cur, len(kc))))
        sc.close()
    assert res[0][0] == res[1][0] and res[0][0] > 30
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("seed,stereo", [(0, True), (1, False)])
def test_fuse(frames, shim, seed, stereo):
    kk, dk, uRk, depk = frames[seed % 2]
    rng = np.random.default_rng(seed)
    Rcw, tcw, Tm = T.pose(rng)
    n = 1500
    pick, Pw, size, raw_min, raw_max = T.landmarks_around(rng, kk, uRk, Rcw, tcw, n, stereo)
    lm_desc = T.noisy_desc(rng, dk[pick])
    Ow = -(Rcw.T.astype(np.float64) @ tcw.astype(np.float64))
    normal = Pw.astype(np.float64) - Ow
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    normal = normal + rng.normal(0, 0.9, (n, 3)) * (rng.random((n, 1)) < 0.5)
    normal = (normal / np.linalg.norm(normal, axis=1, keepdims=True)).astype(np.float32)
    bad = (rng.random(n) < 0.05).astype(np.uint8)
    prot = (rng.random(n) < 0.05).astype(np.int32)
    assoc = np.full(len(kk), -1, np.int32)
    for i in range(7, n, 23):
        if assoc[pick[i]] < 0:
            assoc[pick[i]] = i
    lm_list = np.concatenate([np.arange(n), [-1]])
    res = []
    for sc in twin(n, shim):
        sc.add_mappoints(Pw, lm_desc, normal=normal, size=size, min_dist=raw_min, max_dist=raw_max, bad=bad, n_protected=prot)
        kf = sc.add_frame(kk, dk, T.K, Tm, T.BOUNDS, mbf=T.MBF, stereo=stereo, uR=uRk if stereo else None, depth=depk if stereo else None, assoc=assoc, keyframe=True)
        res.append(sc.fuse(kf, lm_list, 3.0, 5.99, R.settings(th_low=50.0)))
        sc.close()
    assert len(res[0][0]) > 30
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("seed,window,ratio", [(0, 100, 0.9), (2, 40, 0.6)])
def test_search_for_initialization(frames, shim, seed, window, ratio):
    k1, d1, _, _ = frames[0]
    rng = np.random.default_rng(seed)
    n2 = 1700
    src = rng.integers(0, len(k1), n2)
    k2 = k1[src].copy()
    k2["x"] += rng.normal(0, 12, n2).astype(np.float32); k2["y"] += rng.normal(0, 8, n2).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.choice([0.0, 0.0, 0.0, 45.0], n2) + rng.normal(0, 2, n2)).astype(np.float32) % np.float32(360)
    d2 = T.noisy_desc(rng, d1[src], 20)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    res = []
    for sc in twin(1, shim):
        f1 = sc.add_frame(k1, d1, T.K, np.eye(4), T.BOUNDS)
        f2 = sc.add_frame(k2, d2, T.K, np.eye(4), T.BOUNDS)
        res.append(sc.search_for_initialization(f1, f2, prev, window, R.settings(nnratio=ratio, th_low=50.0)))
        sc.close()
    assert res[0][0] == res[1][0] and res[0][0] > 50
    assert np.array_equal(res[0][1], res[1][1]) and res[0][2].tobytes() == res[1][2].tobytes()


def _sim3_scene(frames, rng):
    k1, d1, uR1, dep1 = frames[0]
    R1, t1, T1 = T.pose(rng, 0.05, 0.5)
    R2, t2, T2 = T.pose(rng, 0.05, 0.5)
    N = len(k1)
    z = rng.uniform(4, 30, N)
    Pc1 = np.stack([(k1["x"] - T.CX) * z / T.FX, (k1["y"] - T.CY) * z / T.FX, z], 1)
    Pw = (Pc1 - t1.astype(np.float64)) @ R1.astype(np.float64)
    Pc2 = Pw @ R2.astype(np.float64).T + t2.astype(np.float64)
    u2 = T.FX * Pc2[:, 0] / Pc2[:, 2] + T.CX; v2 = T.FX * Pc2[:, 1] / Pc2[:, 2] + T.CY
    src = np.nonzero((Pc2[:, 2] > 0) & (u2 > 20) & (u2 < 1220) & (v2 > 20) & (v2 < 356))[0]
    src = src[rng.permutation(len(src))]
    k2 = k1[src].copy()
    k2["x"] = (u2[src] + rng.normal(0, 1.0, len(src))).astype(np.float32); k2["y"] = (v2[src] + rng.normal(0, 1.0, len(src))).astype(np.float32)
    d2 = T.noisy_desc(rng, d1[src], 10)
    uR2 = (k2["x"] - T.MBF / Pc2[src, 2]).astype(np.float32); dep2 = Pc2[src, 2].astype(np.float32)
    f1 = np.nonzero(rng.random(N) < 0.5)[0]; f2 = np.nonzero(rng.random(len(src)) < 0.6)[0]
    n1, n2 = len(f1), len(f2)
    a1 = np.full(N, -1, np.int32); a2 = np.full(len(src), -1, np.int32)
    a1[f1] = np.arange(n1); a2[f2] = n1 + np.arange(n2)
    mp = dict(Pw=np.concatenate([(Pw[f1] + rng.normal(0, 0.02, (n1, 3))), (Pw[src[f2]] + rng.normal(0, 0.02, (n2, 3)))]).astype(np.float32),
              desc=np.concatenate([T.noisy_desc(rng, d1[f1], 8), T.noisy_desc(rng, d2[f2], 8)]),
              size=np.concatenate([k1["size"][f1] * z[f1] / T.FX, k1["size"][src[f2]] * z[src[f2]] / T.FX]).astype(np.float32),
              mn=np.concatenate([z[f1], z[src[f2]]]).astype(np.float32) * np.float32(0.15), mx=np.concatenate([z[f1], z[src[f2]]]).astype(np.float32) * np.float32(3.2),
              bad=(rng.random(n1 + n2) < 0.04).astype(np.uint8))
    T12 = T1.astype(np.float64) @ np.linalg.inv(T2.astype(np.float64))
    pre = np.full(N, -1, np.int32)
    pre[f1[::15]] = n1 + rng.integers(0, n2, len(f1[::15]))
    return dict(k1=k1, d1=d1, uR1=uR1, dep1=dep1, k2=k2, d2=d2, uR2=uR2, dep2=dep2, T1=T1, T2=T2, a1=a1, a2=a2, mp=mp,
                R12=T12[:3, :3].astype(np.float32), t12=T12[:3, 3].astype(np.float32), pre=pre, n=n1 + n2)


@pytest.mark.parametrize("seed,s12", [(0, 1.0), (1, 1.07)])
def test_search_by_sim3(frames, shim, seed, s12):
    S = _sim3_scene(frames, np.random.default_rng(seed))
    res = []
    for sc in twin(S["n"], shim):
        mp = S["mp"]
        sc.add_mappoints(mp["Pw"], mp["desc"], size=mp["size"], min_dist=mp["mn"], max_dist=mp["mx"], bad=mp["bad"])
        kf1 = sc.add_frame(S["k1"], S["d1"], T.K, S["T1"], T.BOUNDS, mbf=T.MBF, stereo=True, uR=S["uR1"], depth=S["dep1"], assoc=S["a1"], keyframe=True)
        kf2 = sc.add_frame(S["k2"], S["d2"], T.K, S["T2"], T.BOUNDS, mbf=T.MBF, stereo=True, uR=S["uR2"], depth=S["dep2"], assoc=S["a2"], keyframe=True)
        res.append(sc.search_by_sim3(kf1, kf2, S["pre"], s12, S["R12"], S["t12"], 7.5, R.settings(th_high=100.0)))
        sc.close()
    assert res[0][0] == res[1][0] and res[0][0] > 20
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("seed,only_stereo", [(0, False), (1, True)])
def test_search_for_triangulation_and_bow(frames, shim, seed, only_stereo):
    k1, d1, uR1, dep1 = frames[0]
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(k1))
    k2 = k1[perm].copy()
    k2["x"] -= rng.uniform(2, 30, len(k2)).astype(np.float32)
    k2["y"] += rng.choice([0, 0, 0, 0.4, 3.0], len(k2)).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.choice([0.0, 0.0, 0.0, 40.0], len(k2)) + rng.normal(0, 2, len(k2))).astype(np.float32) % np.float32(360)
    d2 = T.noisy_desc(rng, d1[perm], 12)
    uR2 = uR1[perm] - 1
    nodes1 = rng.integers(0, 60, len(k1)).astype(np.int32)
    nodes2 = nodes1[perm].copy()
    moved = rng.random(len(k2)) < 0.2
    nodes2[moved] = rng.integers(0, 60, int(moved.sum()))
    nodes1[::37] = -1
    n_mp = 400
    a1 = np.full(len(k1), -1, np.int32); a2 = np.full(len(k2), -1, np.int32)
    a1[rng.permutation(len(k1))[:200]] = np.arange(200)
    a2[rng.permutation(len(k2))[:200]] = 200 + np.arange(200)
    bad = np.zeros(n_mp, np.uint8); bad[::9] = 1
    T2 = np.eye(4, dtype=np.float32); T2[0, 3] = -0.5
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    tri, bow = [], []
    for sc in twin(n_mp, shim):
        sc.add_mappoints(np.zeros((n_mp, 3), np.float32) + [0, 0, 5], rng.integers(0, 256, (n_mp, 32), dtype=np.uint8) * 0 + 7, bad=bad)
        kf1 = sc.add_frame(k1, d1, T.K, np.eye(4), T.BOUNDS, mbf=T.MBF, stereo=True, uR=uR1, depth=dep1, assoc=a1, keyframe=True)
        kf2 = sc.add_frame(k2, d2, T.K, T2, T.BOUNDS, mbf=T.MBF, stereo=True, uR=uR2, depth=dep1[perm], assoc=a2, keyframe=True)
        fr = sc.add_frame(k2, d2, T.K, T2, T.BOUNDS)
        sc.set_feature_nodes(kf1, nodes1); sc.set_feature_nodes(kf2, nodes2); sc.set_feature_nodes(fr, nodes2)
        tri.append(sc.search_for_triangulation(kf1, kf2, F12, only_stereo, R.settings(th_low=50.0)))
        bow.append(sc.search_by_bow(kf1, fr, R.settings(nnratio=0.7, th_low=50.0)))
        sc.close()
    assert len(tri[0][0]) > 50 and len(bow[0][0]) > 20
    assert np.array_equal(tri[0][0], tri[1][0]) and np.array_equal(tri[0][1], tri[1][1])
    assert np.array_equal(bow[0][0], bow[1][0]) and np.array_equal(bow[0][1], bow[1][1])
