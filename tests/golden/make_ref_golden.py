"""Generates tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN CODE (oracle/_ref: hySLAM's ORBExtractor / ORBFinder /
Stereomatcher translation units compiled unmodified from /root/reference, monotonic-allocator tie policy).  Run in the build
container (needs /root/reference):  python tests/golden/make_ref_golden.py
The fixtures travel to the GPU box, where /root/reference does not exist."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (parameter structs only)
from oracle import ref as R  # noqa: E402
from hyslam_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def extract_case(name, kind, h, w, seed, nf, full):
    img = (synth.noise_image if kind == "noise" else synth.blocks_image)(h, w, seed)
    p = O.default_params(nf)
    k, d, lv = R.extract(img, p, arena=True, levels=True)
    data = dict(kind=kind, h=h, w=w, seed=seed, nfeatures=nf, image_sha=sha(img), n=len(k), kps_sha=sha(k), desc_sha=sha(d),
                pyr_sha=np.array([sha(x) for x in lv]), level_count=np.bincount(k["octave"], minlength=p.nlevels).astype(np.int32))
    if full:
        data.update(kps=k, desc=d)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, len(k))


def stereo_case(name, kind, h, w, seed, nf):
    L, Rt = synth.stereo_pair(h, w, seed, kind)
    p = O.default_params(nf)
    kl, dl = R.extract(L, p)
    kr, dr = R.extract(Rt, p)
    mbf, fx = 386.1448, 718.856
    uR, depth = R.stereo_match(O.StereoParams(mbf, fx, h, 100.0, 50.0, 31.0), kl, dl, kr, dr)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, h=h, w=w, seed=seed, nfeatures=nf, mbf=mbf, fx=fx, left_sha=sha(L), right_sha=sha(Rt),
                        nl=len(kl), nr=len(kr), kl_sha=sha(kl), dl_sha=sha(dl), kr_sha=sha(kr), dr_sha=sha(dr), uR=uR, depth=depth)
    print(name, len(kl), len(kr), int((uR >= 0).sum()))


if __name__ == "__main__":
    extract_case("ref_c1_noise_752x480_s0", "noise", 480, 752, 0, 1000, True)
    extract_case("ref_c1_blocks_752x480_s1", "blocks", 480, 752, 1, 1000, True)
    for s in range(2, 6):
        extract_case(f"ref_c1_noise_752x480_s{s}", "noise", 480, 752, s, 1000, False)
    extract_case("ref_c2_noise_1241x376_s2", "noise", 376, 1241, 2, 2000, False)
    extract_case("ref_c2_blocks_1241x376_s4", "blocks", 376, 1241, 4, 2000, False)
    extract_case("ref_c3_noise_3840x2160_s0", "noise", 2160, 3840, 0, 8000, False)
    stereo_case("ref_stereo_c2_noise_s0", "noise", 376, 1241, 0, 2000)
    stereo_case("ref_stereo_c2_blocks_s3", "blocks", 376, 1241, 3, 2000)
