"""The C oracle against fixtures produced by the REFERENCE'S OWN CODE (tests/golden/ref_*.npz, written by
tests/golden/make_ref_golden.py from oracle/_ref).  Needs neither /root/reference nor oracle/_ref at run time.  CPU only."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from hyslam_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXT = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "ref_c[12]_*.npz")))
STEREO = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "ref_stereo_*.npz")))


def test_fixtures_present():
    assert len(EXT) >= 8 and len(STEREO) >= 2


@pytest.mark.parametrize("name", EXT)
def test_oracle_extract_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    img = (synth.noise_image if str(g["kind"]) == "noise" else synth.blocks_image)(int(g["h"]), int(g["w"]), int(g["seed"]))
    assert sha(img) == str(g["image_sha"]), "synthetic generator drifted from the golden input"
    k, d, info = O.extract(img, O.default_params(int(g["nfeatures"])), debug=True)
    assert [sha(x) for x in info["pyramid"]] == [str(s) for s in g["pyr_sha"]]
    assert info["level_count"].tolist() == g["level_count"].tolist()
    assert len(k) == int(g["n"]) and sha(k) == str(g["kps_sha"]) and sha(d) == str(g["desc_sha"])
    if "kps" in g.files:
        assert k.tobytes() == g["kps"].tobytes() and np.array_equal(d, g["desc"])


@pytest.mark.parametrize("name", STEREO)
def test_oracle_stereo_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    h, w = int(g["h"]), int(g["w"])
    L, Rt = synth.stereo_pair(h, w, int(g["seed"]), str(g["kind"]))
    assert sha(L) == str(g["left_sha"]) and sha(Rt) == str(g["right_sha"])
    p = O.default_params(int(g["nfeatures"]))
    kl, dl = O.extract(L, p)
    kr, dr = O.extract(Rt, p)
    assert (sha(kl), sha(dl), sha(kr), sha(dr)) == (str(g["kl_sha"]), str(g["dl_sha"]), str(g["kr_sha"]), str(g["dr_sha"]))
    uR, depth, _, _ = O.stereo_match(O.StereoParams(float(g["mbf"]), float(g["fx"]), h, 100.0, 50.0, 31.0), kl, dl, kr, dr)
    assert uR.tobytes() == g["uR"].tobytes() and depth.tobytes() == g["depth"].tobytes()
