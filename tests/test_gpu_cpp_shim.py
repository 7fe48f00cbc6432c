"""The C++ host side (include/hyorb_hyslam.hpp) run through ImageProcessing::ProcessStereoImage's call sequence
(two extractor objects, the left one on its own thread; FeatureViews; Stereomatcher; getData) by tests/cpp/shim_driver,
compared bit for bit with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from hyslam_b200 import _ffi as F, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "_build", "shim_driver")


def _driver():
    if not os.path.exists(DRIVER):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return DRIVER


@pytest.mark.parametrize("kind,h,w,seed,nf", [("noise", 240, 480, 3, 1000), ("blocks", 376, 1241, 6, 2000)])
def test_process_stereo_image_through_the_cpp_shim(tmp_path, kind, h, w, seed, nf):
    L, R = synth.stereo_pair(h, w, seed, kind=kind)
    lp, rp, op = (str(tmp_path / n) for n in ("left.raw", "right.raw", "out.bin"))
    L.tofile(lp); R.tofile(rp)
    mbf, fx = 386.1448, 718.856
    r = subprocess.run([_driver(), lp, rp, str(w), str(h), str(nf), repr(mbf), repr(fx), op], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    print(r.stdout)
    assert "fused stereo ok" in r.stdout          # CudaStereoFrontEnd (one device call) == extractor x 2 + CudaStereomatcher
    assert "camera types ok" in r.stdout          # getExtractor("Imaging") / ("SLAM") read their own settings block
    assert "camera frame ok" in r.stdout          # PreProcessImg + extraction through the shim == the gray path
    buf = open(op, "rb").read()
    nl, nr, nlev, kpsz = (int(v) for v in np.frombuffer(buf, np.int32, 4))
    assert kpsz == 28 and nlev == 8
    o = 16

    def take(dtype, n):
        nonlocal o
        a = np.frombuffer(buf, dtype, n, o)
        o += a.nbytes
        return a
    kl = take(F.KP_DTYPE, nl); dl = take(np.uint8, nl * 32).reshape(nl, 32)
    kr = take(F.KP_DTYPE, nr); dr = take(np.uint8, nr * 32).reshape(nr, 32)
    uR = take(np.float32, nl); depth = take(np.float32, nl)
    bi = take(np.int32, nl); b = take(np.uint16, nl); s = take(np.uint16, nl); acc = take(np.uint8, nl)
    scales = take(np.float32, nlev)
    nd = int(take(np.int32, 1)[0])
    distinctive = take(np.int32, nd)
    ntri = int(take(np.int32, 1)[0])
    tbi = take(np.int32, nl); tb = take(np.uint16, nl); ts = take(np.uint16, nl); tacc = take(np.uint8, nl)
    assert o == len(buf)

    p = O.default_params(nf)
    okl, odl = O.extract(L, p)
    okr, odr = O.extract(R, p)
    assert kl.tobytes() == okl.tobytes() and np.array_equal(dl, odl), "left keypoints / descriptors"
    assert kr.tobytes() == okr.tobytes() and np.array_equal(dr, odr), "right keypoints / descriptors"
    ouR, odepth, _, _ = O.stereo_match(O.StereoParams(mbf, fx, h, 100.0, 50.0, 31.0), okl, odl, okr, odr)
    assert np.array_equal(uR, ouR) and np.array_equal(depth, odepth), "stereo association"
    want = O.match_csr(odl, odr, mode=1, thr=50.0, ratio=0.6)
    for g, wv, name in zip((bi, b, s, acc), want, ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, wv), name
    assert np.array_equal(scales, O.scale_tables(p)[0])
    # CudaDescriptorScan::distinctiveDescriptors: landmark l = left descriptors 3l..3l+2 + right descriptor l
    obs = np.concatenate([np.concatenate([odl[3 * l:3 * l + 3], odr[l:l + 1]]) for l in range(nd)]) if nd else np.zeros((0, 32), np.uint8)
    want_idx, _ = O.distinctive_descriptor(obs, np.arange(0, 4 * nd + 1, 4, dtype=np.int32))
    assert nd > 0 and np.array_equal(distinctive, want_idx)
    # CudaDescriptorScan::searchForTriangulation: first ntri left keypoints x all right keypoints behind the rectified-pair epipolar gate
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    i1 = np.repeat(np.arange(ntri, dtype=np.int32), nr); i2 = np.tile(np.arange(nr, dtype=np.int32), ntri)
    ok = O.epipolar_check(okl, okr, i1, i2, F12).astype(bool).reshape(ntri, nr)
    off, idx = [0], []
    for i in range(nl):
        if i < ntri:
            idx += np.nonzero(ok[i])[0].tolist()
        off.append(len(idx))
    want = O.match_csr(odl, odr, np.array(off, np.int32), np.array(idx if idx else [0], np.int32), mode=1, thr=50.0, ratio=1.0)
    for g, wv, name in zip((tbi, tb, ts, tacc), want, ("best_idx", "best", "second", "accepted")):
        assert np.array_equal(g, wv), "triangulation " + name
    assert ntri > 0 and ok.any() and not ok.all()
    assert "shim ok" in r.stdout
