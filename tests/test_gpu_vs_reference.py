"""GPU parity against the REFERENCE ITSELF: the CUDA path through the C ABI versus (a) fixtures written by the reference's own
translation units (tests/golden/ref_*.npz) at C1 / C2 / C3 full size and (b), when oracle/_ref travelled to this box, the
reference's code executed live on fresh seeds.  Bit-exact, exact order."""
import glob
import hashlib
import os

import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "ref_c*.npz"))))
def test_gpu_extract_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    img = (synth.noise_image if str(g["kind"]) == "noise" else synth.blocks_image)(int(g["h"]), int(g["w"]), int(g["seed"]))
    assert sha(img) == str(g["image_sha"])
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=int(g["nfeatures"])))
    k, d = ex(img, None)
    ex.close()
    assert len(k) == int(g["n"]) and sha(k) == str(g["kps_sha"]) and sha(d) == str(g["desc_sha"])


@pytest.mark.parametrize("name", sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "ref_stereo_*.npz"))))
def test_gpu_stereo_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    h, w = int(g["h"]), int(g["w"])
    L, Rt = synth.stereo_pair(h, w, int(g["seed"]), str(g["kind"]))
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=int(g["nfeatures"])))
    kl, dl = ex(L, None)
    kr, dr = ex(Rt, None)
    ex.close()
    assert (sha(kl), sha(dl), sha(kr), sha(dr)) == (str(g["kl_sha"]), str(g["dl_sha"]), str(g["kr_sha"]), str(g["dr_sha"]))
    sm = hb.Stereomatcher((kl, dl, kr, dr), hb.StereoCamera(float(g["mbf"]), float(g["fx"]), float(h)))
    sm.computeStereoMatches()
    uR, depth = sm.getData()
    assert uR.tobytes() == g["uR"].tobytes() and depth.tobytes() == g["depth"].tobytes()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref did not travel to this box")
@pytest.mark.parametrize("kind,h,w,nf,seeds", [("noise", 480, 752, 1000, range(20, 26)), ("blocks", 480, 752, 1000, range(20, 24)),
                                               ("noise", 376, 1241, 2000, range(20, 24)), ("blocks", 376, 1241, 2000, range(20, 23))])
def test_gpu_matches_reference_code_live(kind, h, w, nf, seeds):
    p = O.default_params(nf)
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=nf))
    for seed in seeds:
        L, Rt = synth.stereo_pair(h, w, seed, kind)
        kl, dl = ex(L, None)
        kr, dr = ex(Rt, None)
        rkl, rdl = R.extract(L, p)
        rkr, rdr = R.extract(Rt, p)
        assert kl.tobytes() == rkl.tobytes() and dl.tobytes() == rdl.tobytes(), (kind, seed)
        assert kr.tobytes() == rkr.tobytes() and dr.tobytes() == rdr.tobytes(), (kind, seed)
        sm = hb.Stereomatcher((kl, dl, kr, dr), hb.StereoCamera(386.1448, 718.856, float(h)))
        sm.computeStereoMatches()
        uR, depth = sm.getData()
        ruR, rdepth = R.stereo_match(O.StereoParams(386.1448, 718.856, h, 100.0, 50.0, 31.0), rkl, rdl, rkr, rdr)
        assert uR.tobytes() == ruR.tobytes() and depth.tobytes() == rdepth.tobytes(), (kind, seed)
    ex.close()
