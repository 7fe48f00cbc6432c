"""GPU parity of stereo association and descriptor matching through the C ABI against the CPU oracle and the
committed golden fixtures: indices, distances, accept flags, uR / depth all bit-exact."""
import os

import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import _ffi as F, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _eq(got, want, what):
    for g, w, n in zip(got, want, ("best_idx", "best", "second", "accepted")):
        bad = np.nonzero(np.asarray(g) != np.asarray(w))[0]
        assert len(bad) == 0, f"{what}: {n} differs at {bad[:8]} ({len(bad)}): got {np.asarray(g)[bad[:4]]} want {np.asarray(w)[bad[:4]]}"


def test_match_rules_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "match_300.npz"))
    m = hb.FeatureMatcher()
    for key in g.files:
        if not key.startswith("m"):
            continue
        mode, thr, ratio = key[1:].split("_")
        _eq(m.match(g["a"], g["b"], rule=int(mode), thr=float(thr), ratio=float(ratio)), g[key], key)
    _eq(m.match(g["a"], g["b"], g["csr_off"], g["csr_idx"], rule=1, thr=50.0, ratio=0.6), g["csr_res"], "csr")


@pytest.mark.parametrize("nq,nt", [(1, 1), (5, 300), (300, 7), (1000, 1000), (8000, 8000)])
def test_bruteforce_matches_oracle(nq, nt):
    """C4 at full size (8000 x 8000) plus ragged shapes; heavy ties (few distinct distances) included."""
    a = synth.random_descriptors(max(nq, nt, 4), 1)
    b, _ = synth.perturbed_descriptors(a, 2)
    q, t = a[:nq], b[:nt]
    m = hb.FeatureMatcher()
    for rule, thr, ratio in [(0, 100.0, 0.9), (1, 50.0, 1.0), (1, 50.0, 0.6), (2, 50.0, 0.9)]:
        _eq(m.match(q, t, rule=rule, thr=thr, ratio=ratio), O.match_csr(q, t, mode=rule, thr=thr, ratio=ratio), f"{nq}x{nt} rule {rule}")
    # ties: every target is one of 4 distinct rows -> first index must win, second == best
    t2 = a[:4][np.random.default_rng(0).integers(0, 4, nt)]
    _eq(m.match(q, t2, rule=0, thr=256.0, ratio=1.0), O.match_csr(q, t2, mode=0, thr=256.0, ratio=1.0), "ties")


def test_self_match_property():
    a = synth.random_descriptors(8000, 5)
    m = hb.FeatureMatcher()
    bi, b, s, acc = m.match(a, a, rule=0, thr=100.0, ratio=0.9)
    assert np.array_equal(bi, np.arange(8000)) and not b.any() and acc.all() and (s > 60).all()


def test_csr_matches_oracle_including_empty_lists():
    rng = np.random.default_rng(3)
    a = synth.random_descriptors(2000, 7)
    b, _ = synth.perturbed_descriptors(a, 8)
    lens = rng.integers(0, 90, 2000); lens[::17] = 0; lens[5] = 1500
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    idx = rng.integers(0, 2000, off[-1]).astype(np.int32)
    m = hb.FeatureMatcher()
    for rule, thr, ratio in [(1, 50.0, 0.6), (0, 100.0, 0.9), (2, 50.0, 0.9)]:
        _eq(m.match(a, b, off, idx, rule=rule, thr=thr, ratio=ratio), O.match_csr(a, b, off, idx, mode=rule, thr=thr, ratio=ratio), f"csr rule {rule}")
    bad = idx.copy(); bad[3] = 2000
    with pytest.raises(hb.HyorbError) as e:
        m.match(a, b, off, bad)
    assert e.value.rc == F.EINVAL


def test_grid_window_rotation_match_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_rotation.npz"))
    m = hb.FeatureMatcher()
    off, idx = m.grid_build(g["kps"], g["bounds"])
    assert np.array_equal(off, g["cell_off"]) and np.array_equal(idx, g["cell_idx"])
    assert np.array_equal(m.RotationConsistency(g["angle_prev"], g["angle_curr"]), g["keep"])
    rng = np.random.default_rng(1)
    for n in (1, 40, 3000):
        ap = rng.uniform(0, 360, n).astype(np.float32); ac = rng.uniform(0, 360, n).astype(np.float32)
        ac[: n // 2] = (ap[: n // 2] - 30 + rng.normal(0, 4, n // 2)).astype(np.float32) % np.float32(360)
        assert np.array_equal(m.RotationConsistency(ap, ac), O.rotation_consistency(ap, ac))


def test_window_matching_matches_oracle():
    img = synth.noise_image(376, 1241, 31)
    s = hb.FeatureExtractorSettings(nFeatures=2000)
    kps, desc = O.extract(img, O.default_params(2000))
    n = len(kps)
    rng = np.random.default_rng(4)
    bounds = (0.0, 1241.0, 0.0, 376.0)
    ob = O.Bounds(*bounds)
    off, idx = O.grid_build(kps, ob)
    nq = 1500
    src = rng.integers(0, n, nq)
    q = np.zeros(nq, F.WQ_DTYPE)
    q["u"] = kps["x"][src] + rng.normal(0, 3, nq).astype(np.float32)
    q["v"] = kps["y"][src] + rng.normal(0, 3, nq).astype(np.float32)
    q["r"] = rng.choice([4.0, 7.5, 15.0, 40.0], nq).astype(np.float32) * kps["size"][src] / 31
    q["size_lo"] = 0.5 * kps["size"][src]; q["size_hi"] = 1.5 * kps["size"][src]
    q["u"][:20] = -50; q["v"][20:40] = 5000                      # off-image queries
    t_uR = np.where(rng.random(n) < 0.7, kps["x"] - rng.uniform(1, 60, n), -1).astype(np.float32)
    q["ur"] = q["u"] - 20; q["ur_radius"] = np.where(rng.random(nq) < 0.5, -1.0, 25.0)
    qd = desc[src].copy()
    flip = rng.integers(0, 256, (nq, 12))
    for i in range(nq):
        for bpos in flip[i]:
            qd[i, bpos >> 3] ^= 1 << (bpos & 7)
    matched = (rng.random(n) < 0.2).astype(np.uint8)
    m = hb.FeatureMatcher()
    for tu, tm in [(t_uR, matched), (t_uR, None)]:
        want = O.match_window(kps, desc, tu, tm, ob, off, idx, q, qd, thr=100.0, ratio=0.9)
        got = m.SearchByProjection(kps, desc, bounds, q, qd, t_uR=tu, t_matched=tm, thr=100.0, ratio=0.9)
        _eq(got, want, "window")
        assert (got[0] >= 0).sum() > 800


def test_stereo_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "stereo_noise_480x240.npz"))
    cam = hb.StereoCamera(float(g["mbf"]), float(g["fx"]), float(g["h"]))
    sm = hb.Stereomatcher((g["kl"], g["dl"], g["kr"], g["dr"]), cam)
    sm.computeStereoMatches()
    uR, depth = sm.getData()
    assert np.array_equal(uR.view(np.uint32), g["uR"].view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), g["depth"].view(np.uint32))


@pytest.mark.parametrize("kind,seed", [("noise", 2), ("blocks", 6)])
def test_c2_extract_plus_stereo_matches_oracle(kind, seed):
    """C2: KITTI-shaped 1241x376 pair, 2000 features per image, extract + ComputeStereoMatches end to end."""
    L, R = synth.stereo_pair(376, 1241, seed, kind)
    s = hb.FeatureExtractorSettings(nFeatures=2000)
    ex = hb.ORBExtractor(s)
    kl, dl = ex(L, None); kr, dr = ex(R, None)
    okl, odl = O.extract(L, O.default_params(2000)); okr, odr = O.extract(R, O.default_params(2000))
    assert np.array_equal(kl, okl) and np.array_equal(kr, okr) and np.array_equal(dl, odl) and np.array_equal(dr, odr)
    cam = hb.StereoCamera(386.1448, 718.856, 376.0)
    sm = hb.Stereomatcher((kl, dl, kr, dr), cam)
    sm.computeStereoMatches()
    uR, depth = sm.getData()
    ouR, odepth, obr, obd = O.stereo_match(O.StereoParams(386.1448, 718.856, 376, 100.0, 50.0, 31.0), okl, odl, okr, odr)
    assert np.array_equal(sm.best_r, obr) and np.array_equal(sm.best_dist, obd)
    # north_star: disparity within 1e-3 px; the reference has no sub-pixel step, so it is in fact bit-exact
    assert np.max(np.abs(uR - ouR), initial=0) <= 1e-3
    assert np.array_equal(uR.view(np.uint32), ouR.view(np.uint32)) and np.array_equal(depth.view(np.uint32), odepth.view(np.uint32))
    if kind == "noise":
        assert (uR >= 0).sum() > 300


def test_process_stereo_batch_matches_oracle():
    """ImageProcessing::ProcessStereoImage over a batch: extract L, extract R, stereo match, one ABI call."""
    pairs = [synth.stereo_pair(376, 1241, 40 + i, "noise" if i != 1 else "blocks") for i in range(3)]
    imgs = np.stack([im for p in pairs for im in p])
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=2000))
    cam = hb.StereoCamera(386.1448, 718.856, 376.0)
    kps, desc, counts, uR, depth = ex.process_stereo_batch(imgs, cam, capacity=2560)
    osp = O.StereoParams(386.1448, 718.856, 376, 100.0, 50.0, 31.0)
    for p in range(3):
        okl, odl = O.extract(imgs[2 * p], O.default_params(2000)); okr, odr = O.extract(imgs[2 * p + 1], O.default_params(2000))
        nl, nr = counts[2 * p], counts[2 * p + 1]
        assert nl == len(okl) and nr == len(okr)
        assert np.array_equal(kps[2 * p, :nl], okl) and np.array_equal(kps[2 * p + 1, :nr], okr)
        assert np.array_equal(desc[2 * p, :nl], odl) and np.array_equal(desc[2 * p + 1, :nr], odr)
        ouR, odepth, _, _ = O.stereo_match(osp, okl, odl, okr, odr)
        assert np.array_equal(uR[p, :nl].view(np.uint32), ouR.view(np.uint32))
        assert np.array_equal(depth[p, :nl].view(np.uint32), odepth.view(np.uint32))
    st, calls = ex.stage_times()
    assert calls == 0                      # profiling is off by default


def test_stereo_with_large_keypoint_sizes_imaging_config():
    """config/slam_feature_config.yaml's Imaging camera: 3000 features, scale 1.4 -- level-7 keypoints are 31 * 1.4^7 = 326 px wide, their
    row band (r = 2 * size / 31 = 21 rows either side) is wider than the default row-table budget; both stereo entry points must size the
    table from the keypoints instead of refusing the frame (round-1 ADVICE)."""
    L, R = synth.stereo_pair(900, 1600, 11, "noise")
    s = hb.FeatureExtractorSettings(nFeatures=3000, fScaleFactor=1.4, nLevels=8)
    p = O.default_params(3000, 1.4, 8, 30)
    okl, odl = O.extract(L, p); okr, odr = O.extract(R, p)
    assert okr["size"].max() > 250
    osp = O.StereoParams(386.1448, 718.856, 900, 100.0, 50.0, 31.0)
    ouR, odepth, obr, obd = O.stereo_match(osp, okl, odl, okr, odr)
    cam = hb.StereoCamera(386.1448, 718.856, 900.0)
    # Stereomatcher mirror (hyorb_stereo_match_host)
    sm = hb.Stereomatcher((okl, odl, okr, odr), cam)
    sm.computeStereoMatches()
    uR, depth = sm.getData()
    assert np.array_equal(sm.best_r, obr) and np.array_equal(sm.best_dist, obd)
    assert np.array_equal(uR.view(np.uint32), ouR.view(np.uint32)) and np.array_equal(depth.view(np.uint32), odepth.view(np.uint32))
    # ProcessStereoImage in one call (hyorb_process_stereo_batch_host)
    ex = hb.ORBExtractor(s)
    kps, desc, counts, uR2, depth2 = ex.process_stereo_batch(np.stack([L, R]), cam, capacity=ex.default_capacity())
    nl = counts[0]
    assert nl == len(okl) and np.array_equal(kps[0, :nl], okl)
    assert np.array_equal(uR2[0, :nl].view(np.uint32), ouR.view(np.uint32)) and np.array_equal(depth2[0, :nl].view(np.uint32), odepth.view(np.uint32))
