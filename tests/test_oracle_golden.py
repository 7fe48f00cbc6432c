"""Pins the C oracle against the committed golden fixtures (tests/golden/*.npz), which were
produced by an independent cv2 + Python restatement of the reference (tests/golden/make_golden.py).
CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from hyslam_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


EXT = ["ext_noise_320x240", "ext_blocks_376x280", "ext_blocks_640x360", "ext_c1_noise_752x480"]


def _image(g):
    fn = synth.noise_image if str(g["kind"]) == "noise" else synth.blocks_image
    img = fn(int(g["h"]), int(g["w"]), int(g["seed"]))
    assert sha(img) == str(g["image_sha"]), "synthetic generator drifted from the golden input"
    return img


@pytest.mark.parametrize("name", EXT)
def test_extract_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    img = _image(g)
    p = O.default_params(int(g["nfeatures"]), float(g["scale_factor"]), int(g["nlevels"]), int(g["cell_px"]))
    k, d, info = O.extract(img, p, debug=True)
    assert [sha(x) for x in info["pyramid"]] == [str(s) for s in g["pyr_sha"]]
    assert [len(c[0]) for c in info["cand"]] == g["cand_count"].tolist()
    assert info["level_count"].tolist() == g["level_count"].tolist()
    gk = g["kps"]
    assert len(k) == len(gk)
    for f in gk.dtype.names:       # exact order, bit-exact fields (float compared as bits)
        assert np.array_equal(k[f].view(np.uint32) if k[f].dtype.kind == "f" else k[f],
                              gk[f].view(np.uint32) if gk[f].dtype.kind == "f" else gk[f]), f
    assert np.array_equal(d, g["desc"])


def test_stereo_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "stereo_noise_480x240.npz"))
    L, R = synth.stereo_pair(int(g["h"]), int(g["w"]), int(g["seed"]))
    assert sha(L) == str(g["left_sha"]) and sha(R) == str(g["right_sha"])
    p = O.default_params(int(g["nfeatures"]), float(g["scale_factor"]), int(g["nlevels"]), int(g["cell_px"]))
    kl, dl = O.extract(L, p)
    kr, dr = O.extract(R, p)
    assert np.array_equal(kl, g["kl"]) and np.array_equal(kr, g["kr"])
    assert np.array_equal(dl, g["dl"]) and np.array_equal(dr, g["dr"])
    sp = O.StereoParams(float(g["mbf"]), float(g["fx"]), int(g["h"]), 100.0, 50.0, 31.0)
    uR, depth, _, _ = O.stereo_match(sp, kl, dl, kr, dr)
    assert np.array_equal(uR.view(np.uint32), g["uR"].view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), g["depth"].view(np.uint32))
    assert (uR >= 0).sum() > 50


def test_match_rules_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "match_300.npz"))
    a, b = g["a"], g["b"]
    for key in g.files:
        if not key.startswith("m"):
            continue
        mode, thr, ratio = key[1:].split("_")
        bi, bd, bs, ac = O.match_csr(a, b, mode=int(mode), thr=float(thr), ratio=float(ratio))
        r = g[key]
        assert np.array_equal(bi, r[0]) and np.array_equal(bd, r[1]) and np.array_equal(bs, r[2])
        assert np.array_equal(ac, r[3])
    bi, bd, bs, ac = O.match_csr(a, b, g["csr_off"], g["csr_idx"], mode=1, thr=50.0, ratio=0.6)
    r = g["csr_res"]
    assert np.array_equal(bi, r[0]) and np.array_equal(bd, r[1]) and np.array_equal(bs, r[2]) and np.array_equal(ac, r[3])


def test_grid_and_rotation_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_rotation.npz"))
    k = g["kps"]
    b = O.Bounds(*[float(v) for v in g["bounds"]])
    off, idx = O.grid_build(k, b)
    assert np.array_equal(off, g["cell_off"]) and np.array_equal(idx, g["cell_idx"])
    for i, (x, y, r) in enumerate(g["queries"]):
        got = O.grid_query(k, b, off, idx, float(x), float(y), float(r))
        assert np.array_equal(got, g["q_idx"][g["q_off"][i]:g["q_off"][i + 1]])
    keep = O.rotation_consistency(g["angle_prev"], g["angle_curr"])
    assert np.array_equal(keep, g["keep"])


def test_hamming_known_answers():
    z = np.zeros(32, np.uint8)
    f = np.full(32, 255, np.uint8)
    assert O.hamming(z, z) == 0 and O.hamming(z, f) == 256
    one = z.copy(); one[17] = 0x10
    assert O.hamming(z, one) == 1
    rng = np.random.default_rng(5)
    for _ in range(200):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert O.hamming(a, b) == int(np.unpackbits(a ^ b).sum())


def test_scale_tables_match_survey_values():
    # SURVEY.md section 8 derived sizes (computed with the reference's float arithmetic)
    p = O.default_params(1000)
    assert O.scale_tables(p)[4].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert O.level_sizes(p, 752, 480) == [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231), (302, 193), (252, 161), (210, 134)]
    p2 = O.default_params(2000)
    assert O.scale_tables(p2)[4].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert O.level_sizes(p2, 1241, 376)[-1] == (346, 105)
    p3 = O.default_params(8000)
    assert O.scale_tables(p3)[4].tolist() == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    assert O.level_sizes(p3, 3840, 2160)[-1] == (1072, 603)
