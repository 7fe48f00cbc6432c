"""Landmark representative descriptor (MapPointDBEntry::_computeDistinctiveDescriptor_, src/core/MapPointDB.cpp:127-171):
the C oracle against a literal numpy restatement (CPU), the CUDA kernel against the oracle through the C ABI (GPU)."""
import numpy as np
import pytest

from hyslam_b200 import synth
from oracle import oracle as O


def _literal(desc):
    """the reference loop, literally: float distance matrix, int row copy, sort, vDists[0.5*(N-1)], strict <"""
    n = len(desc)
    bits = np.unpackbits(desc, axis=1).astype(np.int32)
    D = (bits[:, None, :] != bits[None, :, :]).sum(2).astype(np.float32)
    best_median, best_idx = np.finfo(np.float32).max, 0
    for i in range(n):
        v = np.sort(D[i].astype(np.int32))
        median = int(v[int(0.5 * (n - 1))])
        if median < best_median:
            best_median, best_idx = median, i
    return best_idx, int(best_median)


def _landmarks(seed, n_lm, max_obs):
    """observation lists: a base descriptor with a few flipped bits per observation (many equal medians -> tie cases),
    plus empty and single-observation landmarks"""
    rng = np.random.default_rng(seed)
    rows, off = [], [0]
    for l in range(n_lm):
        n = int(rng.integers(0, max_obs + 1)) if l % 7 else (l // 7) % 3        # 0, 1, 2 observations appear regularly
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        for _ in range(n):
            d = np.unpackbits(base)
            flips = rng.choice(256, size=int(rng.integers(0, 12)), replace=False)
            d[flips] ^= 1
            rows.append(np.packbits(d))
        off.append(off[-1] + n)
    desc = np.stack(rows) if rows else np.zeros((0, 32), np.uint8)
    return desc, np.array(off, np.int32)


def test_oracle_matches_the_literal_loop():
    desc, off = _landmarks(1, 60, 40)
    bi, bm = O.distinctive_descriptor(desc, off)
    for l in range(len(off) - 1):
        d = desc[off[l]:off[l + 1]]
        if len(d) == 0:
            assert bi[l] == -1 and bm[l] == -1
        else:
            assert (int(bi[l]), int(bm[l])) == _literal(d), l


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_lm,max_obs", [(2, 300, 30), (3, 50, 200), (4, 1, 1), (5, 2000, 12)])
def test_gpu_matches_oracle(seed, n_lm, max_obs):
    import hyslam_b200 as hb
    desc, off = _landmarks(seed, n_lm, max_obs)
    m = hb.FeatureMatcher()
    bi, bm = m.ComputeDistinctiveDescriptors(desc, off)
    obi, obm = O.distinctive_descriptor(desc, off)
    assert np.array_equal(bi, obi) and np.array_equal(bm, obm)


@pytest.mark.gpu
def test_gpu_real_descriptors_and_errors():
    import hyslam_b200 as hb
    from hyslam_b200 import _ffi as F
    a = synth.random_descriptors(500, 9)
    off = np.arange(0, 501, 50, dtype=np.int32)
    m = hb.FeatureMatcher()
    bi, bm = m.ComputeDistinctiveDescriptors(a, off)
    obi, obm = O.distinctive_descriptor(a, off)
    assert np.array_equal(bi, obi) and np.array_equal(bm, obm)
    with pytest.raises(hb.HyorbError) as e:
        m.ComputeDistinctiveDescriptors(a, np.array([0, 10, 5], np.int32))
    assert e.value.rc == F.EINVAL
