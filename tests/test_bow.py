"""Bag-of-words quantisation and BoW-gated matching.  DBoW2 and its vocabulary are not part of the
reference tree (un-vendored dependency, missing ORBvoc file): the oracle restates DBoW2's PUBLISHED transform algorithm and
parity is UNPINNED by the reference for this function; what is checked is oracle == literal Python restatement (CPU) and
CUDA == oracle on seeded stand-in vocabularies (GPU), including the candidate gating of _SearchByBoW_."""
import numpy as np
import pytest

from hyslam_b200 import synth
from hyslam_b200.matcher import Vocabulary
from oracle import oracle as O


def _features(tree, n, seed):
    """descriptors near random leaves of the tree (a few flipped bits) plus pure noise"""
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(tree["word_of"] >= 0)[0]
    base = tree["node_desc"][rng.choice(leaves, n)]
    bits = np.unpackbits(base, axis=1)
    flips = rng.random(bits.shape) < 0.08
    out = np.packbits(bits ^ flips, axis=1)
    out[::9] = rng.integers(0, 256, (len(out[::9]), 32), dtype=np.uint8)
    return out


def _literal_transform(tree, d, levelsup):
    pop = lambda a, b: int(np.unpackbits(a ^ b).sum())
    node, level, nid = 0, 0, (0 if tree["L"] - levelsup <= 0 else -1)
    co, ci = tree["child_off"], tree["child_idx"]
    while co[node + 1] > co[node]:
        level += 1
        kids = ci[co[node]:co[node + 1]]
        best, best_d = kids[0], pop(d, tree["node_desc"][kids[0]])
        for k in kids[1:]:
            dd = pop(d, tree["node_desc"][k])
            if dd < best_d:
                best, best_d = k, dd
        node = int(best)
        if level == tree["L"] - levelsup:
            nid = node
    return int(tree["word_of"][node]), (nid if nid >= 0 else node), float(tree["weight_of"][node])


def test_oracle_matches_literal_descent():
    tree = Vocabulary.random_tree(5, 4, 3, shrink=0.3)
    f = _features(tree, 60, 1)
    for levelsup in (0, 2, 4, 7):
        w, nid, wt = O.bow_transform(tree, f, levelsup)
        for i in range(len(f)):
            assert (int(w[i]), int(nid[i]), float(wt[i])) == _literal_transform(tree, f[i], levelsup)
    # every feature ends in a leaf; nodes at level L - levelsup have that depth
    assert (w >= 0).all()


def _search_reference(tree, d1, d2, m1, m2, levelsup, rule, thr, ratio):
    """_SearchByBoW_ by composition of oracle pieces: quantise, group set 2 by node (index order), explicit CSR lists"""
    _, n1, _ = O.bow_transform(tree, d1, levelsup)
    _, n2, _ = O.bow_transform(tree, d2, levelsup)
    order = np.argsort(n2, kind="stable")
    if m2 is not None:
        order = order[m2[order] != 0]
    keys = n2[order]
    off, idx = [0], []
    for i in range(len(d1)):
        if m1 is None or m1[i]:
            lo, hi = np.searchsorted(keys, n1[i], "left"), np.searchsorted(keys, n1[i], "right")
            idx += order[lo:hi].tolist()
        off.append(len(idx))
    return O.match_csr(d1, d2, np.array(off, np.int32), np.array(idx if idx else [0], np.int32), mode=rule, thr=thr, ratio=ratio), n1, n2


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,shrink,levelsup", [(10, 4, 0.0, 2), (6, 5, 0.4, 4), (40, 2, 0.2, 1), (3, 3, 0.0, 5)])
def test_gpu_transform_matches_oracle(k, L, shrink, levelsup):
    import hyslam_b200 as hb
    tree = Vocabulary.random_tree(k, L, 10 * k + L, shrink=shrink)
    f = _features(tree, 3000, 2)
    v = Vocabulary(tree)
    m = hb.FeatureMatcher()
    w, nid, wt = m.BowTransform(v, f, levelsup)
    ow, onid, owt = O.bow_transform(tree, f, levelsup)
    assert np.array_equal(w, ow) and np.array_equal(nid, onid) and np.array_equal(wt, owt)
    nodes, off, order = m.feature_vector(nid)
    assert off[-1] == len(f) and (np.diff(nodes) > 0).all()
    bv = m.bow_vector(w, wt)
    assert abs(sum(bv.values()) - 1.0) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("seed,with_masks", [(0, False), (1, True)])
def test_gpu_search_by_bow_matches_oracle_composition(seed, with_masks):
    import hyslam_b200 as hb
    tree = Vocabulary.random_tree(10, 4, 77)
    rng = np.random.default_rng(seed)
    d1 = _features(tree, 2500, 3 + seed)
    d2, _ = synth.perturbed_descriptors(d1, 5 + seed, max_flips=30)
    m1 = (rng.random(len(d1)) < 0.8).astype(np.uint8) if with_masks else None
    m2 = (rng.random(len(d2)) < 0.8).astype(np.uint8) if with_masks else None
    v = Vocabulary(tree)
    m = hb.FeatureMatcher()
    bi, b, s, acc, n1, n2 = m.SearchByBoW(v, d1, d2, m1, m2, levelsup=2, rule=1, thr=50.0, ratio=0.6)
    (obi, ob, osd, oacc), on1, on2 = _search_reference(tree, d1, d2, m1, m2, 2, 1, 50.0, 0.6)
    assert np.array_equal(n1, on1) and np.array_equal(n2, on2)
    for g, w_, name in zip((bi, b, s, acc), (obi, ob, osd, oacc), ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, w_), name
    assert acc.sum() > 200


@pytest.mark.gpu
def test_vocabulary_errors():
    import hyslam_b200 as hb
    from hyslam_b200 import _ffi as F
    tree = Vocabulary.random_tree(3, 2, 1)
    bad = dict(tree); bad["child_idx"] = tree["child_idx"].copy(); bad["child_idx"][0] = 999
    with pytest.raises(hb.HyorbError) as e:
        Vocabulary(bad)
    assert e.value.rc == F.EINVAL


def _epipolar_scene(tree, n, seed, general):
    """Two keypoint sets whose descriptors match (perturbed copies) and whose positions lie near each other's epipolar lines
    (noise of a few pixels, so that the 3.84 * sigma2(size) gate accepts some candidates and rejects others)."""
    from hyslam_b200 import _ffi as F
    rng = np.random.default_rng(seed)
    d1 = _features(tree, n, 11 + seed)
    d2, perm = synth.perturbed_descriptors(d1, 13 + seed, max_flips=30)
    if general:
        Fm = (rng.normal(size=(3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-2], [1e-3, 1e-2, 1.0]])).astype(np.float32)
    else:
        Fm = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)          # rectified pair: the line of (x1, y1) is y2 = y1
    k1 = np.zeros(n, F.KP_DTYPE); k2 = np.zeros(n, F.KP_DTYPE)
    k1["x"] = rng.uniform(20, 1200, n).astype(np.float32); k1["y"] = rng.uniform(20, 350, n).astype(np.float32)
    octv = rng.integers(0, 8, n)
    k1["size"] = (31 * 1.2 ** octv).astype(np.float32); k1["octave"] = octv
    a = k1["x"].astype(np.float64) * Fm[0, 0] + k1["y"] * Fm[1, 0] + Fm[2, 0]
    b = k1["x"].astype(np.float64) * Fm[0, 1] + k1["y"] * Fm[1, 1] + Fm[2, 1]
    c = k1["x"].astype(np.float64) * Fm[0, 2] + k1["y"] * Fm[1, 2] + Fm[2, 2]
    # set-2 feature j is a perturbed copy of set-1 feature src[j]: put it near that feature's line
    src = np.empty(n, np.int64)
    src[:] = -1
    if perm is not None and len(perm) == n:
        src = np.asarray(perm)
    x2 = rng.uniform(20, 1200, n)
    y2 = rng.uniform(20, 350, n)
    ok = (src >= 0) & (np.abs(b[np.clip(src, 0, n - 1)]) > 1e-9)
    s = np.clip(src, 0, n - 1)
    y2[ok] = (-(a[s] * x2 + c[s]) / b[s])[ok] + rng.normal(0, 2.0, n)[ok]
    k2["x"] = x2.astype(np.float32); k2["y"] = y2.astype(np.float32)
    o2 = rng.integers(0, 8, n)
    k2["size"] = (31 * 1.2 ** o2).astype(np.float32); k2["octave"] = o2
    return k1, d1, k2, d2, Fm


@pytest.mark.gpu
@pytest.mark.parametrize("seed,general,with_masks", [(0, False, False), (1, True, True), (2, True, False)])
def test_gpu_search_for_triangulation_matches_oracle_composition(seed, general, with_masks):
    """FeatureMatcher::SearchForTriangulation (FeatureMatcher.cc:373-402): BoW gating + EpipolarConsistencyBoWCriterion +
    BestMatchBoWCriterion(TH_LOW, 1.0), against the same chain composed from oracle pieces."""
    import hyslam_b200 as hb
    tree = Vocabulary.random_tree(10, 4, 78)
    k1, d1, k2, d2, Fm = _epipolar_scene(tree, 2500, seed, general)
    rng = np.random.default_rng(100 + seed)
    m1 = (rng.random(len(d1)) < 0.8).astype(np.uint8) if with_masks else None
    m2 = (rng.random(len(d2)) < 0.8).astype(np.uint8) if with_masks else None
    v = Vocabulary(tree)
    m = hb.FeatureMatcher()
    bi, b, s, acc, n1, n2 = m.SearchForTriangulation(v, k1, d1, k2, d2, Fm, m1, m2, levelsup=2, thr=50.0, ratio=1.0)
    # reference chain: node lists (index order) -> epipolar filter -> best match
    _, on1, _ = O.bow_transform(tree, d1, 2)
    _, on2, _ = O.bow_transform(tree, d2, 2)
    order = np.argsort(on2, kind="stable")
    if m2 is not None:
        order = order[m2[order] != 0]
    keys = on2[order]
    pi1, pi2, rows = [], [], []
    for i in range(len(d1)):
        if m1 is None or m1[i]:
            lo, hi = np.searchsorted(keys, on1[i], "left"), np.searchsorted(keys, on1[i], "right")
            c = order[lo:hi]
            pi1 += [i] * len(c); pi2 += c.tolist()
        rows.append(len(pi2))
    ok = O.epipolar_check(k1, k2, pi1, pi2, Fm).astype(bool)
    pi2 = np.array(pi2, np.int64)
    off, idx, prev = [0], [], 0
    for i in range(len(d1)):
        seg = slice(prev, rows[i])
        idx += pi2[seg][ok[seg]].tolist()
        off.append(len(idx)); prev = rows[i]
    obi, ob, osd, oacc = O.match_csr(d1, d2, np.array(off, np.int32), np.array(idx if idx else [0], np.int32), mode=1, thr=50.0, ratio=1.0)
    assert np.array_equal(n1, on1) and np.array_equal(n2, on2)
    for g, w_, name in zip((bi, b, s, acc), (obi, ob, osd, oacc), ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, w_), name
    assert ok.sum() > 500 and (~ok).sum() > 500, "the gate must both accept and reject"
    # the same gate in front of explicit candidate lists (the unfiltered node lists as CSR)
    uoff = np.array([0] + rows, np.int32)
    cbi, cb, cs, cacc = m.match_csr_epipolar(k1, d1, k2, d2, uoff, pi2.astype(np.int32) if len(pi2) else np.zeros(1, np.int32), Fm, rule=1, thr=50.0, ratio=1.0)
    for g, w_, name in zip((cbi, cb, cs, cacc), (obi, ob, osd, oacc), ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, w_), "csr " + name
    assert acc.sum() > 100


@pytest.mark.gpu
def test_epipolar_gate_edge_cases():
    """den == 0 (a degenerate F12 maps every keypoint to the zero line) rejects every candidate (MatchCriteria.cpp:670-671); empty
    candidate lists and an empty query set are handled; the gate needs candidate lists."""
    import hyslam_b200 as hb
    from hyslam_b200 import _ffi as F
    tree = Vocabulary.random_tree(6, 3, 5)
    k1, d1, k2, d2, Fm = _epipolar_scene(tree, 300, 7, False)
    m = hb.FeatureMatcher()
    n = len(d1)
    off = np.arange(0, n + 1, dtype=np.int32) * 4
    idx = np.random.default_rng(0).integers(0, n, 4 * n).astype(np.int32)
    bi, b, s, acc = m.match_csr_epipolar(k1, d1, k2, d2, off, idx, np.zeros((3, 3), np.float32), thr=256.0)
    assert (bi == -1).all() and not acc.any()
    ok = O.epipolar_check(k1, k2, np.repeat(np.arange(n, dtype=np.int32), 4), idx, np.zeros(9, np.float32))
    assert not ok.any()
    # empty lists for every query
    bi, b, s, acc = m.match_csr_epipolar(k1, d1, k2, d2, np.zeros(n + 1, np.int32), np.zeros(1, np.int32), Fm, thr=256.0)
    assert (bi == -1).all() and not acc.any()
    # no queries at all
    bi, b, s, acc = m.match_csr_epipolar(k1[:0], d1[:0], k2, d2, np.zeros(1, np.int32), np.zeros(1, np.int32), Fm)
    assert len(bi) == 0
    v = Vocabulary(tree)
    out = m.SearchForTriangulation(v, k1[:0], d1[:0], k2, d2, Fm)
    assert len(out[0]) == 0
    with pytest.raises(hb.HyorbError):
        F.check(F.lib().hyorb_match_csr_epipolar_host(m._h, F.ptr(k1), F.ptr(d1), n, F.ptr(k2), F.ptr(d2), n, None, None, F.ptr(Fm.reshape(9)),
                                                      1.0, 31.0, 1, 50.0, 1.0, F.ptr(bi), F.ptr(b), F.ptr(s), F.ptr(acc)))
