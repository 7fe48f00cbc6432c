"""Pins the level-synchronous ("closed form") quadtree restatement -- the formulation the CUDA
kernel uses -- against the literal list-based restatement of DistributeOctTree
(reference src/features/ORBExtractor.cpp:179-403), including output ORDER.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O


def _points(rng, kind, W, H, n):
    if kind == "uniform":
        x = rng.integers(3, W - 3, n); y = rng.integers(3, H - 3, n)
    elif kind == "cluster":
        k = max(1, n // 40)
        cx = rng.integers(10, W - 10, k); cy = rng.integers(10, H - 10, k)
        sel = rng.integers(0, k, n)
        x = np.clip(cx[sel] + rng.integers(-6, 7, n), 3, W - 4); y = np.clip(cy[sel] + rng.integers(-6, 7, n), 3, H - 4)
    elif kind == "line":
        x = rng.integers(3, W - 3, n); y = np.full(n, H // 3) + rng.integers(0, 2, n)
    else:  # right edge: exercises the root-assignment overflow column
        x = np.clip(W - 4 - rng.integers(0, 5, n), 3, W - 4); y = rng.integers(3, H - 3, n)
    pts = np.unique(np.stack([x, y], 1), axis=0)
    rng.shuffle(pts)
    resp = rng.integers(20, 60, len(pts)).astype(np.float32)     # few distinct values: many response ties
    return pts[:, 0].astype(np.float32), pts[:, 1].astype(np.float32), resp


@pytest.mark.parametrize("kind", ["uniform", "cluster", "line", "edge"])
def test_closed_form_equals_literal(kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    dims = [(720, 448), (1209, 344), (178, 102), (400, 400), (300, 100), (150, 213), (3808, 2128)]
    for trial in range(120):
        W, H = dims[trial % len(dims)]
        n = int(rng.choice([1, 2, 3, 7, 40, 200, 1000, 4000]))
        N = int(rng.choice([1, 5, 60, 217, 434, 1000, 1737]))
        x, y, r = _points(rng, kind, W, H, n)
        a = O.distribute_octtree(x, y, r, 16, 16 + W, 16, 16 + H, N)
        b = O.distribute_octtree(x, y, r, 16, 16 + W, 16, 16 + H, N, closed_form=True)
        assert np.array_equal(a, b), (kind, trial, W, H, len(x), N)


def test_closed_form_on_real_candidates():
    from hyslam_b200 import synth
    for fn, seed in [(synth.noise_image, 0), (synth.blocks_image, 3), (synth.noise_image, 7)]:
        img = fn(376, 1241, seed)
        p = O.default_params(2000)
        _, _, info = O.extract(img, p, debug=True)
        quota = O.scale_tables(p)[4]
        for l, (cx, cy, cr) in enumerate(info["cand"]):
            w, h = info["sizes"][l]
            a = O.distribute_octtree(cx, cy, cr, 16, w - 16, 16, h - 16, int(quota[l]))
            b = O.distribute_octtree(cx, cy, cr, 16, w - 16, 16, h - 16, int(quota[l]), closed_form=True)
            assert np.array_equal(a, b), (seed, l)
            assert len(a) == info["level_count"][l]
