"""Randomised shapes and extractor settings (seeded): odd widths, tiny upper levels, other scale factors / level counts / cell
sizes, strided views -- the whole extraction path against the oracle, bit for bit.  Exercises the TMA tile edges, the
blur / resize edge handling and the quadtree limits away from the benchmark shapes."""
import os

import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        w, h = int(rng.integers(150, 900)), int(rng.integers(120, 520))
        nlev = int(rng.integers(2, 9)); scale = float(rng.choice([1.1, 1.2, 1.25, 1.33, 1.5])); cell = int(rng.choice([20, 30, 40]))
        nf = int(rng.integers(100, 2500))
        top = min(w, h) / scale ** (nlev - 1)
        if top < 32 + cell + 8 or max(w, h) / min(w, h) > 8:
            continue                                   # the reference itself breaks on levels smaller than one cell
        out.append((w, h, nlev, scale, cell, nf, int(rng.integers(0, 1 << 30)), str(rng.choice(["noise", "blocks"]))))
    return out


@pytest.mark.parametrize("w,h,nlev,scale,cell,nf,seed,kind", _cases(int(os.environ.get("HYORB_FUZZ_CASES", "16")), int(os.environ.get("HYORB_FUZZ_SEED", "2024"))))
def test_random_shape_matches_oracle(w, h, nlev, scale, cell, nf, seed, kind):
    img = (synth.noise_image if kind == "noise" else synth.blocks_image)(h, w, seed % 1000)
    s = hb.FeatureExtractorSettings(nFeatures=nf, fScaleFactor=scale, nLevels=nlev, N_CELLS=cell)
    p = O.default_params(nf, scale, nlev, cell)
    try:
        ok, od = O.extract(img, p, cap=8 * nf + 4096)
    except RuntimeError:
        pytest.skip("oracle refuses this configuration (level smaller than a cell)")
    ex = hb.ORBExtractor(s)
    try:
        k, d = ex(img, None, capacity=8 * nf + 4096)
    except hb.HyorbError as e:
        assert e.rc == hb._ffi.EUNSUPPORTED, e
        pytest.skip(f"library reports the configuration as unsupported: {e}")
    assert len(k) == len(ok)
    assert k.tobytes() == ok.tobytes()
    assert np.array_equal(d, od)
