"""PARITY PIN, matcher rows: the C oracle's restatement of the window / projection / BoW matchers against the REFERENCE'S OWN
FeatureMatcher.cc + MatchCriteria.cpp + Frame.cc + KeyFrame.cc + MapPoint.cc + Camera.cpp + LandMarkMatches.cpp (oracle/_ref,
compiled unmodified), driven through the reference's public entry points (FeatureMatcher.h:105-176).  What is compared is what a
hySLAM caller observes: the frame's landmark associations after the call, match lists, match counts.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R
from hyslam_b200 import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and reference tree absent")

FX, CX, CY, MBF = 718.856, 607.19, 185.22, 386.1448
K = np.array([[FX, 0, CX], [0, FX, CY], [0, 0, 1]], np.float32)
BOUNDS = (0.0, 1241.0, 0.0, 376.0)


@pytest.fixture(scope="module")
def frames():
    out = []
    p = O.default_params(2000)
    sp = O.StereoParams(MBF, FX, 376, 100.0, 50.0, 31.0)
    for seed in (40, 41):
        L, Rt = synth.stereo_pair(376, 1241, seed)
        kl, dl = O.extract(L, p)
        kr, dr = O.extract(Rt, p)
        uR, depth, _, _ = O.stereo_match(sp, kl, dl, kr, dr)
        out.append((kl, dl, uR, depth))
    return out


def pose(rng, rot=0.1, trans=1.0):
    a, b, c = rng.normal(0, rot, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    Rcw = (Rx @ Ry @ Rz).astype(np.float32)
    tcw = rng.normal(0, trans, 3).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = Rcw; T[:3, 3] = tcw
    return Rcw, tcw, T


def landmarks_around(rng, kps, uR, Rcw, tcw, n, stereo):
    """landmarks near back-projected keypoints (so that windows hold candidates), some behind / beside the camera or out of range"""
    pick = rng.integers(0, len(kps), n)
    z = rng.uniform(2, 40, n)
    if stereo:
        has = uR[pick] > 0
        z[has] = MBF / np.maximum(kps["x"][pick][has] - uR[pick][has], 0.5)
    u = kps["x"][pick] + rng.normal(0, 3, n); v = kps["y"][pick] + rng.normal(0, 3, n)
    Pc = np.stack([(u - CX) * z / FX, (v - CY) * z / FX, z], 1)
    Pc[::11, 2] *= -1
    Pc[5::13, 0] += 80
    Pw = ((Pc - tcw.astype(np.float64)) @ Rcw.astype(np.float64)).astype(np.float32)
    size = (kps["size"][pick] * z / FX * rng.uniform(0.6, 1.6, n)).astype(np.float32)
    raw_min = rng.uniform(0.5, 7, n).astype(np.float32); raw_max = rng.uniform(12, 100, n).astype(np.float32)
    return pick, Pw, size, raw_min, raw_max


def noisy_desc(rng, d, max_flips=7):
    d = d.copy()
    for i in range(len(d)):
        for bpos in rng.integers(0, 256, rng.integers(0, max_flips)):
            d[i, bpos >> 3] ^= 1 << (bpos & 7)
    return d


def associate(assoc, idx, lm):
    """LandMarkMatches::associateLandMark(i, pMP, replace = true)  (LandMarkMatches.cpp:23-46)"""
    old = np.nonzero(assoc == lm)[0]
    assoc[idx] = lm
    for o in old:
        if o != idx:
            assoc[o] = -1


def oracle_search_by_projection(kps, desc, uR, stereo, Rcw, tcw, Ow, lm, lm_desc, assoc0, n_obs, th, thr, ratio, use_distance=True, prev_angle=None,
                                check_matched=True, use_stereo=True):
    """the oracle's composition of FeatureMatcher::_SearchByProjection_ (FeatureMatcher.cc:57-121) incl. the association loop"""
    pr = O.make_projection(Rcw, tcw, Ow, K, MBF, stereo, BOUNDS)
    lms = np.zeros(len(lm["Pw"]), O.LM_DTYPE)
    lms["Pw"] = lm["Pw"]; lms["size"] = lm["size"]
    lms["min_dist"] = np.float32(0.8) * lm["raw_min"] if use_distance else 0          # MapPoint::GetMinDistanceInvariance: 0.8f * mfMinDistance
    lms["max_dist"] = np.float32(1.2) * lm["raw_max"] if use_distance else np.float32(3e38)
    lms["assoc_idx"] = -1
    for i, a in enumerate(assoc0):
        if 0 <= a < len(lms):
            lms["assoc_idx"][a] = i
    q, passed = O.project_landmarks(pr, lms, kps, th)
    if not use_stereo:
        q["ur_radius"] = -1
    bounds = O.Bounds(*BOUNDS)
    off, idx = O.grid_build(kps, bounds)
    t_matched = np.array([1 if (a >= 0 and n_obs[a] > 0) else 0 for a in assoc0], np.uint8) if check_matched else None
    bi, b, s, acc = O.match_window(kps, desc, uR if stereo else None, t_matched, bounds, off, idx, q, lm_desc, thr=thr, ratio=ratio)
    acc[passed == 0] = 0
    if prev_angle is not None:
        acc = O.projection_rotation(bi, acc, prev_angle, kps)
    assoc = assoc0.copy()
    for i in range(len(lms)):                      # std::map<MapPoint*, ...> order == landmark id order
        if acc[i]:
            associate(assoc, bi[i], i)
    return assoc, int(acc.sum())


@pytest.mark.parametrize("seed,stereo,th,ratio", [(0, True, 3.0, 0.9), (1, False, 5.0, 0.9), (2, True, 1.0, 0.6), (3, True, 7.0, 0.75)])
def test_search_by_projection_local_map(frames, seed, stereo, th, ratio):
    """FeatureMatcher::SearchByProjection(Frame&, vector<MapPoint*>, th)  (FeatureMatcher.cc:123-143)"""
    kl, dl, uR, depth = frames[seed % 2]
    rng = np.random.default_rng(seed)
    Rcw, tcw, T = pose(rng)
    n = 1500
    pick, Pw, size, raw_min, raw_max = landmarks_around(rng, kl, uR, Rcw, tcw, n, stereo)
    lm_desc = noisy_desc(rng, dl[pick])
    # initial associations: some candidate landmarks already sit on a keypoint of the frame (no observations yet), and 8 % of
    # the keypoints carry OTHER landmarks that have observations (these keypoints are not available, MatchCriteria.cpp:124-144)
    assoc0 = np.full(len(kl), -1, np.int32)
    n_obs = np.zeros(n + len(kl), np.int32)
    taken = rng.random(len(kl)) < 0.08
    extra = np.nonzero(taken)[0]
    assoc0[extra] = n + np.arange(len(extra))
    n_obs[n: n + len(extra)] = 1
    for i in range(3, n, 17):
        if assoc0[pick[i]] < 0:
            assoc0[pick[i]] = i
    sc = R.Scene(n + len(extra))
    sc.add_mappoints(Pw, lm_desc, size=size, min_dist=raw_min, max_dist=raw_max)
    sc.add_mappoints(np.zeros((len(extra), 3), np.float32) + [0, 0, 5], dl[extra], size=np.full(len(extra), 0.1, np.float32))
    for j in range(len(extra)):
        sc.set_observation_count(n + j, 1)
    f = sc.add_frame(kl, dl, K, T, BOUNDS, mbf=MBF, stereo=stereo, uR=uR if stereo else None, depth=depth if stereo else None, assoc=assoc0)
    st = R.settings(nnratio=ratio, th_high=100.0, th_low=50.0)
    nref = sc.search_by_projection(f, np.arange(n), th, st)
    got = sc.assoc(f, len(kl))
    lm = dict(Pw=Pw, size=size, raw_min=raw_min, raw_max=raw_max)
    want, nor = oracle_search_by_projection(kl, dl, uR, stereo, Rcw, tcw, sc.camera_center(f), lm, lm_desc, assoc0, n_obs, th, 100.0, ratio)
    assert nref == nor and nref > 10
    assert np.array_equal(got, want)
    sc.close()


@pytest.mark.parametrize("seed,stereo", [(4, True), (5, False)])
def test_search_by_projection_motion_model(frames, seed, stereo):
    """FeatureMatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono)  (:145-176): landmarks = the last frame's
    associations, no distance criterion, RotationConsistencyCriterion over the matches"""
    kc, dc, uRc, depc = frames[0]
    kp, dp, uRp, depp = frames[1]
    rng = np.random.default_rng(seed)
    Rcw, tcw, T = pose(rng)
    n = 1200
    pick, Pw, size, raw_min, raw_max = landmarks_around(rng, kc, uRc, Rcw, tcw, n, stereo)
    lm_desc = noisy_desc(rng, dc[pick])
    # the last frame sees landmark i at keypoint prev_idx[i] (distinct keypoints); its angle drives the rotation histogram
    prev_idx = rng.permutation(len(kp))[:n]
    kp = kp.copy()
    kp["angle"][prev_idx] = (kc["angle"][pick] + rng.choice([0.0, 0.0, 0.0, 25.0, 170.0], n) + rng.normal(0, 2, n)).astype(np.float32) % np.float32(360)
    assoc_prev = np.full(len(kp), -1, np.int32)
    assoc_prev[prev_idx] = np.arange(n)
    sc = R.Scene(n)
    sc.add_mappoints(Pw, lm_desc, size=size, min_dist=raw_min, max_dist=raw_max)
    cur = sc.add_frame(kc, dc, K, T, BOUNDS, mbf=MBF, stereo=stereo, uR=uRc if stereo else None, depth=depc if stereo else None)
    last = sc.add_frame(kp, dp, K, np.eye(4), BOUNDS, mbf=MBF, stereo=stereo, uR=uRp if stereo else None, depth=depp if stereo else None, assoc=assoc_prev)
    st = R.settings(nnratio=0.9)
    nref = sc.search_by_projection_motion(cur, last, 7.0, st, mono=not stereo)
    got = sc.assoc(cur, len(kc))
    # replicatemvpMapPoints(): landmarks in the order of the last frame's keypoints; pointer order (= id order) rules the maps
    lm = dict(Pw=Pw, size=size, raw_min=raw_min, raw_max=raw_max)
    prev_angle = kp["angle"][prev_idx]
    want, nor = oracle_search_by_projection(kc, dc, uRc, stereo, Rcw, tcw, sc.camera_center(cur), lm, lm_desc, np.full(len(kc), -1, np.int32),
                                            np.zeros(n, np.int32), 7.0, 100.0, 0.9, use_distance=False, prev_angle=prev_angle)
    assert nref == nor and nref > 10
    assert np.array_equal(got, want)
    sc.close()


def test_search_by_projection_relocalisation(frames):
    """FeatureMatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)  (:180-213): the keyframe's landmarks minus
    those already found; no stereo criterion; BestScore(ORBdist, 1.0); with no previous frame the rotation criterion passes everything"""
    kc, dc, uRc, depc = frames[0]
    kk, dk, uRk, depk = frames[1]
    rng = np.random.default_rng(9)
    Rcw, tcw, T = pose(rng)
    n = 1000
    pick, Pw, size, raw_min, raw_max = landmarks_around(rng, kc, uRc, Rcw, tcw, n, True)
    lm_desc = noisy_desc(rng, dc[pick])
    kf_idx = rng.permutation(len(kk))[:n]
    assoc_kf = np.full(len(kk), -1, np.int32)
    assoc_kf[kf_idx] = np.arange(n)
    found = np.arange(0, n, 7)
    sc = R.Scene(n)
    sc.add_mappoints(Pw, lm_desc, size=size, min_dist=raw_min, max_dist=raw_max)
    cur = sc.add_frame(kc, dc, K, T, BOUNDS, mbf=MBF, stereo=True, uR=uRc, depth=depc)
    kf = sc.add_frame(kk, dk, K, np.eye(4), BOUNDS, mbf=MBF, stereo=True, uR=uRk, depth=depk, assoc=assoc_kf, keyframe=True)
    nref = sc.search_by_projection_reloc(cur, kf, found, 10.0, 64, R.settings())
    got = sc.assoc(cur, len(kc))
    keep = np.setdiff1d(np.arange(n), found)
    lm = dict(Pw=Pw[keep], size=size[keep], raw_min=raw_min[keep], raw_max=raw_max[keep])
    want, nor = oracle_search_by_projection(kc, dc, uRc, True, Rcw, tcw, sc.camera_center(cur), lm, lm_desc[keep], np.full(len(kc), -1, np.int32),
                                            np.zeros(n, np.int32), 10.0, 64.0, 1.0, use_stereo=False)
    want = np.where(want >= 0, keep[np.clip(want, 0, len(keep) - 1)], -1)
    assert nref == nor and nref > 10
    assert np.array_equal(got, want)
    sc.close()


def _bow_oracle(k1, d1, k2, d2, nodes1, nodes2, ok1, ok2, thr, ratio, F12=None):
    """the oracle's composition of FeatureMatcher::_SearchByBoW_ (FeatureMatcher.cc:281-345): per shared node, every eligible
    feature of set 1 scans the eligible features of set 2 under that node (optionally behind the epipolar gate) with
    BestMatchBoWCriterion, then RotationConsistencyBoW.  Returns the surviving (idx1, idx2) pairs in idx1 order."""
    by_node = {}
    for j in np.nonzero(ok2)[0]:
        by_node.setdefault(int(nodes2[j]), []).append(int(j))
    off, idx = [0], []
    for i in range(len(k1)):
        cand = by_node.get(int(nodes1[i]), []) if ok1[i] else []
        if cand and F12 is not None:
            keep = O.epipolar_check(k1, k2, np.full(len(cand), i, np.int32), np.array(cand, np.int32), F12).astype(bool)
            cand = [c for c, kflag in zip(cand, keep) if kflag]
        idx += cand
        off.append(len(idx))
    bi, b, s, acc = O.match_csr(d1, d2, np.array(off, np.int32), np.array(idx if idx else [0], np.int32), mode=1, thr=thr, ratio=ratio)
    i1 = np.nonzero(acc)[0]
    i2 = bi[i1]
    keep = O.rotation_consistency(k2["angle"][i2], k1["angle"][i1]).astype(bool)      # rot = angle(views2) - angle(views1)
    return i1[keep].astype(np.int32), i2[keep].astype(np.int32)


@pytest.mark.parametrize("seed,only_stereo", [(0, False), (1, True)])
def test_search_for_triangulation(frames, seed, only_stereo):
    """FeatureMatcher::SearchForTriangulation  (FeatureMatcher.cc:373-402): features WITHOUT a (good) landmark on both sides, optional
    stereo-only filter, epipolar gate (MatchCriteria.cpp:641-676), BestMatchBoW(TH_LOW, 1.0), rotation histogram"""
    k1, d1, uR1, dep1 = frames[0]
    rng = np.random.default_rng(seed)
    # second keyframe: the same features shifted along x (a rectified pair: epipolar lines are the rows), noisy descriptors, permuted
    perm = rng.permutation(len(k1))
    k2 = k1[perm].copy()
    k2["x"] -= rng.uniform(2, 30, len(k2)).astype(np.float32)
    k2["y"] += rng.choice([0, 0, 0, 0.4, 3.0], len(k2)).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.choice([0.0, 0.0, 0.0, 40.0], len(k2)) + rng.normal(0, 2, len(k2))).astype(np.float32) % np.float32(360)
    d2 = noisy_desc(rng, d1[perm], 12)
    uR2 = uR1[perm] - 1
    nodes1 = rng.integers(0, 60, len(k1)).astype(np.int32)
    nodes2 = nodes1[perm].copy()
    moved = rng.random(len(k2)) < 0.2
    nodes2[moved] = rng.integers(0, 60, int(moved.sum()))
    nodes1[::37] = -1                                                   # features the vocabulary did not place
    n_mp = 400
    a1 = np.full(len(k1), -1, np.int32); a2 = np.full(len(k2), -1, np.int32)
    a1[rng.permutation(len(k1))[:200]] = np.arange(200)
    a2[rng.permutation(len(k2))[:200]] = 200 + np.arange(200)
    bad = np.zeros(n_mp, np.uint8); bad[::9] = 1                      # an association to a BAD landmark does not count (:556-575)
    sc = R.Scene(n_mp)
    sc.add_mappoints(np.zeros((n_mp, 3), np.float32) + [0, 0, 5], rng.integers(0, 256, (n_mp, 32), dtype=np.uint8), bad=bad)
    T2 = np.eye(4, dtype=np.float32); T2[0, 3] = -0.5
    kf1 = sc.add_frame(k1, d1, K, np.eye(4), BOUNDS, mbf=MBF, stereo=True, uR=uR1, depth=dep1, assoc=a1, keyframe=True)
    kf2 = sc.add_frame(k2, d2, K, T2, BOUNDS, mbf=MBF, stereo=True, uR=uR2, depth=dep1[perm], assoc=a2, keyframe=True)
    sc.set_feature_nodes(kf1, nodes1); sc.set_feature_nodes(kf2, nodes2)
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    gi1, gi2 = sc.search_for_triangulation(kf1, kf2, F12, only_stereo, R.settings(th_low=50.0))
    free1 = np.array([a < 0 or bad[a] for a in a1]) & (nodes1 >= 0)
    free2 = np.array([a < 0 or bad[a] for a in a2])
    if only_stereo:
        free1 &= uR1 >= 0
        free2 &= uR2 >= 0
    wi1, wi2 = _bow_oracle(k1, d1, k2, d2, nodes1, nodes2, free1, free2, 50.0, 1.0, F12)
    assert len(gi1) > 50
    assert np.array_equal(gi1, wi1) and np.array_equal(gi2, wi2)
    sc.close()


def test_search_by_bow_keyframe_frame(frames):
    """FeatureMatcher::SearchByBoW(KeyFrame*, Frame&, matches)  (FeatureMatcher.cc:216-280): keyframe features WITH a good landmark
    against all frame features of the same node, BestMatchBoW(TH_LOW, nnratio), rotation histogram; result keyed by frame index"""
    k1, d1, uR1, dep1 = frames[1]
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(k1))
    k2 = k1[perm].copy()
    k2["angle"] = (k2["angle"] + rng.choice([0.0, 0.0, 0.0, 40.0], len(k2)) + rng.normal(0, 2, len(k2))).astype(np.float32) % np.float32(360)
    d2 = noisy_desc(rng, d1[perm], 12)
    nodes1 = rng.integers(0, 80, len(k1)).astype(np.int32)
    nodes2 = nodes1[perm].copy()
    n_mp = 900
    a1 = np.full(len(k1), -1, np.int32)
    a1[rng.permutation(len(k1))[:n_mp]] = np.arange(n_mp)
    bad = np.zeros(n_mp, np.uint8); bad[::11] = 1
    sc = R.Scene(n_mp)
    sc.add_mappoints(np.zeros((n_mp, 3), np.float32) + [0, 0, 5], rng.integers(0, 256, (n_mp, 32), dtype=np.uint8), bad=bad)
    kf = sc.add_frame(k1, d1, K, np.eye(4), BOUNDS, assoc=a1, keyframe=True)
    fr = sc.add_frame(k2, d2, K, np.eye(4), BOUNDS)
    sc.set_feature_nodes(kf, nodes1); sc.set_feature_nodes(fr, nodes2)
    gidx, glm = sc.search_by_bow(kf, fr, R.settings(nnratio=0.7, th_low=50.0))
    have = np.array([a >= 0 and not bad[a] for a in a1])
    wi1, wi2 = _bow_oracle(k1, d1, k2, d2, nodes1, nodes2, have, np.ones(len(k2), bool), 50.0, 0.7)
    want = {}
    for i1, i2 in zip(wi1, wi2):                                       # matches[idx_f] = lm: later keyframe features overwrite
        want[int(i2)] = int(a1[i1])
    assert len(gidx) > 100
    assert gidx.tolist() == sorted(want) and glm.tolist() == [want[k] for k in sorted(want)]
    sc.close()


@pytest.mark.parametrize("seed,window,ratio", [(0, 100, 0.9), (1, 40, 0.9), (2, 100, 0.6)])
def test_search_for_initialization(frames, seed, window, ratio):
    """FeatureMatcher::SearchForInitialization (FeatureMatcher.cc:404-462): the sequential mono-initialisation matcher
    (MonoInitScoreExceedsPrevious + MonoInitBestScore, MatchCriteria.cpp:486-549; window 100 at MonoInitializer.cpp:83)"""
    k1, d1, _, _ = frames[0]
    rng = np.random.default_rng(seed)
    n2 = 1700
    src = rng.integers(0, len(k1), n2)                  # frame 2 re-observes frame-1 features (some several times: contention)
    k2 = k1[src].copy()
    k2["x"] += rng.normal(0, 12, n2).astype(np.float32); k2["y"] += rng.normal(0, 8, n2).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.choice([0.0, 0.0, 0.0, 45.0], n2) + rng.normal(0, 2, n2)).astype(np.float32) % np.float32(360)
    d2 = noisy_desc(rng, d1[src], 20)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    sc = R.Scene(1)
    f1 = sc.add_frame(k1, d1, K, np.eye(4), BOUNDS)
    f2 = sc.add_frame(k2, d2, K, np.eye(4), BOUNDS)
    nref, m12, pm = sc.search_for_initialization(f1, f2, prev, window, R.settings(nnratio=ratio, th_low=50.0))
    nor, om12, opm = O.search_for_initialization(k1, d1, k2, d2, O.Bounds(*BOUNDS), prev, window, 50.0, ratio)
    assert nref == nor and nref > 100
    assert np.array_equal(m12, om12)
    assert pm.tobytes() == opm.tobytes()
    sc.close()


@pytest.mark.parametrize("seed,stereo", [(0, True), (1, False)])
def test_fuse(frames, seed, stereo):
    """FeatureMatcher::Fuse(pKF, landmarks, fuse_matches, th, reprojection_err)  (FeatureMatcher.cc:464-521): pre-screen (bad / already in
    the keyframe / protected), Projection + Distance + ViewingAngle(1.047) criteria, FeatureSize + ProjectionView + BestScore(TH_LOW, 1.0);
    the first landmark that reaches a keypoint keeps it (std::map::insert)"""
    kk, dk, uRk, depk = frames[seed % 2]
    rng = np.random.default_rng(seed)
    Rcw, tcw, T = pose(rng)
    n = 1500
    pick, Pw, size, raw_min, raw_max = landmarks_around(rng, kk, uRk, Rcw, tcw, n, stereo)
    lm_desc = noisy_desc(rng, dk[pick])
    Ow_true = -(Rcw.T.astype(np.float64) @ tcw.astype(np.float64))
    normal = Pw.astype(np.float64) - Ow_true
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    tilt = rng.normal(0, 0.9, (n, 3))
    normal = (normal + tilt * (rng.random((n, 1)) < 0.5))            # half of them viewed at an angle, some beyond 60 degrees
    normal = (normal / np.linalg.norm(normal, axis=1, keepdims=True)).astype(np.float32)
    bad = (rng.random(n) < 0.05).astype(np.uint8)
    prot = (rng.random(n) < 0.05).astype(np.int32)
    assoc = np.full(len(kk), -1, np.int32)
    for i in range(7, n, 23):                                         # landmarks the keyframe already observes: pre-screened out
        if assoc[pick[i]] < 0:
            assoc[pick[i]] = i
    sc = R.Scene(n)
    sc.add_mappoints(Pw, lm_desc, normal=normal, size=size, min_dist=raw_min, max_dist=raw_max, bad=bad, n_protected=prot)
    kf = sc.add_frame(kk, dk, K, T, BOUNDS, mbf=MBF, stereo=stereo, uR=uRk if stereo else None, depth=depk if stereo else None, assoc=assoc, keyframe=True)
    gidx, glm = sc.fuse(kf, np.arange(n), 3.0, 5.99, R.settings(th_low=50.0))
    # oracle composition
    in_kf = np.zeros(n, bool); in_kf[assoc[assoc >= 0]] = True
    cand = np.nonzero(~(bad.astype(bool) | in_kf | (prot > 0)))[0]
    Ow = sc.camera_center(kf)
    pr = O.make_projection(Rcw, tcw, Ow, K, MBF, stereo, BOUNDS)
    lms = np.zeros(len(cand), O.LM_DTYPE)
    lms["Pw"] = Pw[cand]; lms["size"] = size[cand]; lms["min_dist"] = np.float32(0.8) * raw_min[cand]; lms["max_dist"] = np.float32(1.2) * raw_max[cand]
    lms["assoc_idx"] = -1
    q, passed = O.project_landmarks(pr, lms, kk, 3.0)
    q["ur_radius"] = -1                                              # no StereoConsistencyCriterion in Fuse
    passed &= O.viewing_angle(Ow, Pw[cand], normal[cand], 1.047)
    bounds = O.Bounds(*BOUNDS)
    off, idx = O.grid_build(kk, bounds)
    bi, b, s, acc = O.match_window_ex(kk, dk, uRk if stereo else None, None, bounds, off, idx, q, lm_desc[cand], 50.0, 1.0, rule=0, q_active=passed,
                                      reproj_thr=5.99)
    want = {}
    for j in range(len(cand)):
        if acc[j] and int(bi[j]) not in want:
            want[int(bi[j])] = int(cand[j])
    assert len(gidx) > 30
    assert gidx.tolist() == sorted(want) and glm.tolist() == [want[k] for k in sorted(want)]
    sc.close()


@pytest.mark.parametrize("seed,s12", [(0, 1.0), (1, 1.07)])
def test_search_by_sim3(frames, seed, s12):
    """FeatureMatcher::SearchBySim3 (FeatureMatcher.cc:739-937): each keyframe's landmarks are carried into the other camera through the
    similarity [s12 R12 | t12], best Hamming in the window (<= TH_HIGH), and only mutually agreeing matches survive"""
    k1, d1, uR1, dep1 = frames[0]
    rng = np.random.default_rng(seed)
    R1, t1, T1 = pose(rng, 0.05, 0.5)
    R2, t2, T2 = pose(rng, 0.05, 0.5)
    # one physical point per KF1 feature; KF2 sees (most of) them where its own pose projects them, with a noisy descriptor.  Both
    # keyframes carry their OWN MapPoint for a point (duplicates -- the situation loop closing resolves): ids 0..n1-1 belong to KF1,
    # n1.. to KF2
    N = len(k1)
    z = rng.uniform(4, 30, N)
    Pc1 = np.stack([(k1["x"] - CX) * z / FX, (k1["y"] - CY) * z / FX, z], 1)
    Pw = (Pc1 - t1.astype(np.float64)) @ R1.astype(np.float64)
    Pc2 = Pw @ R2.astype(np.float64).T + t2.astype(np.float64)
    u2 = FX * Pc2[:, 0] / Pc2[:, 2] + CX; v2 = FX * Pc2[:, 1] / Pc2[:, 2] + CY
    vis = (Pc2[:, 2] > 0) & (u2 > 20) & (u2 < 1220) & (v2 > 20) & (v2 < 356)
    src = np.nonzero(vis)[0]
    src = src[rng.permutation(len(src))]
    k2 = k1[src].copy()
    k2["x"] = (u2[src] + rng.normal(0, 1.0, len(src))).astype(np.float32); k2["y"] = (v2[src] + rng.normal(0, 1.0, len(src))).astype(np.float32)
    d2 = noisy_desc(rng, d1[src], 10)
    uR2 = (k2["x"] - MBF / Pc2[src, 2]).astype(np.float32); dep2 = Pc2[src, 2].astype(np.float32)
    has1 = rng.random(N) < 0.5                                       # KF1 features that carry a MapPoint
    has2 = rng.random(len(src)) < 0.6
    f1 = np.nonzero(has1)[0]; f2 = np.nonzero(has2)[0]
    n1, n2 = len(f1), len(f2)
    a1 = np.full(N, -1, np.int32); a2 = np.full(len(src), -1, np.int32)
    a1[f1] = np.arange(n1); a2[f2] = n1 + np.arange(n2)
    Pw1 = (Pw[f1] + rng.normal(0, 0.02, (n1, 3))).astype(np.float32); Pw2 = (Pw[src[f2]] + rng.normal(0, 0.02, (n2, 3))).astype(np.float32)
    size1 = (k1["size"][f1] * z[f1] / FX).astype(np.float32); size2 = (k1["size"][src[f2]] * z[src[f2]] / FX).astype(np.float32)
    rmin1 = (z[f1] * 0.5).astype(np.float32); rmax1 = (z[f1] * 1.6).astype(np.float32)
    rmin2 = (z[src[f2]] * 0.5).astype(np.float32); rmax2 = (z[src[f2]] * 1.6).astype(np.float32)
    ld1 = noisy_desc(rng, d1[f1], 8); ld2 = noisy_desc(rng, d2[f2], 8)
    u1 = f1
    bad = (rng.random(n1 + n2) < 0.04).astype(np.uint8)
    sc = R.Scene(n1 + n2)
    sc.add_mappoints(np.concatenate([Pw1, Pw2]), np.concatenate([ld1, ld2]), size=np.concatenate([size1, size2]),
                     min_dist=np.concatenate([rmin1, rmin2]) * 0.3, max_dist=np.concatenate([rmax1, rmax2]) * 2, bad=bad)
    kf1 = sc.add_frame(k1, d1, K, T1, BOUNDS, mbf=MBF, stereo=True, uR=uR1, depth=dep1, assoc=a1, keyframe=True)
    kf2 = sc.add_frame(k2, d2, K, T2, BOUNDS, mbf=MBF, stereo=True, uR=uR2, depth=dep2, assoc=a2, keyframe=True)
    # relative transform camera2 -> camera1 as the loop closer would hand it over: T12 = T1 * inv(T2) (+ scale)
    T12 = T1.astype(np.float64) @ np.linalg.inv(T2.astype(np.float64))
    R12 = T12[:3, :3].astype(np.float32); t12 = T12[:3, 3].astype(np.float32)
    pre = np.full(len(k1), -1, np.int32)
    pre[u1[::15]] = n1 + rng.integers(0, n2, len(u1[::15]))          # matches found earlier (e.g. by SearchByBoW): both ends are skipped
    nref, m12 = sc.search_by_sim3(kf1, kf2, pre, s12, R12, t12, 7.5, R.settings(th_high=100.0))
    # ---- oracle composition
    f32 = np.float32
    sR12 = (f32(s12) * R12).astype(f32)                              # cv::Mat * scalar: float multiply by (float)alpha
    sR21 = (f32(1.0 / s12) * R12.T).astype(f32)
    t21 = -np.array([(sR21[i, 0] * t12[0] + sR21[i, 1] * t12[1]) + sR21[i, 2] * t12[2] for i in range(3)], f32)       # -sR21*t12: small-matrix gemm, alpha -1
    bounds = O.Bounds(*BOUNDS)
    lmA = dict(Pw=np.concatenate([Pw1, Pw2]), size=np.concatenate([size1, size2]), mn=f32(0.8) * (np.concatenate([rmin1, rmin2]) * 0.3).astype(f32),
               mx=f32(1.2) * (np.concatenate([rmax1, rmax2]) * 2).astype(f32), desc=np.concatenate([ld1, ld2]))
    matched1 = pre >= 0
    matched2 = np.zeros(len(k2), bool)
    for i in np.nonzero(matched1)[0]:
        j = np.nonzero(a2 == pre[i])[0]                              # pMP->GetIndexInKeyFrame(pKF2)
        if len(j):
            matched2[j[0]] = True

    def direction(kA, assocA, matchedA, RA, tA, sR, t, kB, dB, assocB, RB, tB, OwB):
        feats = np.array([i for i in range(len(kA)) if assocA[i] >= 0 and not matchedA[i] and not bad[assocA[i]]], np.int32)
        ids = assocA[feats]
        lms = np.zeros(len(ids), O.LM_DTYPE)
        lms["Pw"] = lmA["Pw"][ids]; lms["size"] = lmA["size"][ids]; lms["min_dist"] = lmA["mn"][ids]; lms["max_dist"] = lmA["mx"][ids]
        lms["assoc_idx"] = -1
        for j, lid in enumerate(ids):                               # pKF_B->landMarkSizePixels: B's own association, if any
            w = np.nonzero(assocB == lid)[0]
            if len(w):
                lms["assoc_idx"][j] = w[0]
        prB = O.make_projection(RB, tB, OwB, K, MBF, True, BOUNDS)
        q, passed = O.project_sim3(RA, tA, sR, t, prB, lms, kB, 7.5)
        off, idx = O.grid_build(kB, bounds)
        bi, b, s, acc = O.match_window_ex(kB, dB, None, None, bounds, off, idx, q, lmA["desc"][ids], 100.0, np.inf, rule=0, q_active=passed)
        out = np.full(len(kA), -1, np.int32)
        out[feats[acc > 0]] = bi[acc > 0]
        return out
    vn1 = direction(k1, a1, matched1, R1, t1, sR21, t21, k2, d2, a2, R2, t2, sc.camera_center(kf2))
    vn2 = direction(k2, a2, matched2, R2, t2, sR12, t12, k1, d1, a1, R1, t1, sc.camera_center(kf1))
    want = pre.copy()
    nfound = 0
    for i1 in range(len(k1)):
        if vn1[i1] >= 0 and vn2[vn1[i1]] == i1:
            want[i1] = a2[vn1[i1]]
            nfound += 1
    assert nref == nfound and nref > 20
    assert np.array_equal(m12, want)
    sc.close()
