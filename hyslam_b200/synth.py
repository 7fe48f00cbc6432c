"""Seeded synthetic frames, integer arithmetic only so that every platform
generates identical bytes from the same seed (numpy's PCG64 integer streams are portable).

  noise_image   "G-noise":  dense corners (uniform noise, binomial blur, min-max stretch)
  blocks_image  "G-blocks": sparse corners (flat canvas + random rectangles + +-2 noise);
                            exercises empty cells and the quadtree's early exits
  stereo_right  rectified right view: per-row-band integer disparity + independent +-1 noise
"""
import numpy as np


def _binomial_blur(a, passes=2):
    a = a.astype(np.int32)
    for _ in range(passes):
        p = np.pad(a, ((0, 0), (2, 2)), mode="reflect")
        a = (p[:, :-4] + 4 * p[:, 1:-3] + 6 * p[:, 2:-2] + 4 * p[:, 3:-1] + p[:, 4:] + 8) >> 4
        p = np.pad(a, ((2, 2), (0, 0)), mode="reflect")
        a = (p[:-4] + 4 * p[1:-3] + 6 * p[2:-2] + 4 * p[3:-1] + p[4:] + 8) >> 4
    return a


def noise_image(h, w, seed):
    rng = np.random.default_rng(seed)
    a = _binomial_blur(rng.integers(0, 256, (h, w), dtype=np.uint8), passes=1)
    mn, mx = int(a.min()), int(a.max())
    return (((a - mn) * 255 + (mx - mn) // 2) // max(mx - mn, 1)).astype(np.uint8)


def blocks_image(h, w, seed):
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128, np.int32)
    n = max(4, w * h // 2400)
    for _ in range(n):
        bw, bh = int(rng.integers(10, 121)), int(rng.integers(10, 121))
        x0, y0 = int(rng.integers(-20, w)), int(rng.integers(-20, h))
        img[max(y0, 0):max(y0 + bh, 0), max(x0, 0):max(x0 + bw, 0)] = int(rng.integers(0, 256))
    img += rng.integers(-2, 3, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def stereo_right(left, seed, band=32, dmin=4, dmax=64):
    rng = np.random.default_rng(seed ^ 0x5EED)
    h, w = left.shape
    right = np.empty_like(left)
    for y0 in range(0, h, band):
        d = int(rng.integers(dmin, dmax + 1))
        rows = left[y0:y0 + band]
        # right(x) = left(x + d): a point at uL appears at uR = uL - d
        shifted = np.concatenate([rows[:, d:], rows[:, ::-1][:, :d]], axis=1)
        right[y0:y0 + band] = shifted
    noise = rng.integers(-1, 2, (h, w))
    return np.clip(right.astype(np.int32) + noise, 0, 255).astype(np.uint8)


def stereo_pair(h, w, seed, kind="noise"):
    left = noise_image(h, w, seed) if kind == "noise" else blocks_image(h, w, seed)
    return left, stereo_right(left, seed)


def random_descriptors(n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def perturbed_descriptors(a, seed, max_flips=80, frac_related=0.75):
    """C4 workload: B = permutation of A, k~U{0..max_flips} random bit flips for frac_related of
    the rows, fresh random rows for the rest."""
    rng = np.random.default_rng(seed)
    n = len(a)
    perm = rng.permutation(n)
    b = a[perm].copy()
    bits = np.unpackbits(b, axis=1)
    for i in range(n):
        if rng.random() < frac_related:
            k = int(rng.integers(0, max_flips + 1))
            pos = rng.choice(256, size=k, replace=False)
            bits[i, pos] ^= 1
        else:
            bits[i] = rng.integers(0, 2, 256, dtype=np.uint8)
    return np.packbits(bits, axis=1), perm
