"""GPU parity of the camera-image preparation (ImageProcessing::PreProcessImg, ImageProcessing.cpp:118-138: scale 1.0 / 0.5
then RGB|BGR[A] -> gray) in front of the extractor, through the C ABI against the CPU oracle (itself pinned to cv2 by
tests/test_oracle_vs_cv2.py).  Bit-exact."""
import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import _ffi as F, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _color(h, w, cn, seed):
    """A colour frame whose gray version has corners: a synthetic gray scene modulated per channel + noise."""
    rng = np.random.default_rng(seed)
    base = synth.noise_image(h, w, seed).astype(np.int32)
    if cn == 1:
        return base.astype(np.uint8)
    ch = [np.clip(base * g // 16 + rng.integers(-6, 7, (h, w)), 0, 255) for g in (15, 16, 13, 0)[:cn]]
    if cn == 4:
        ch[3] = rng.integers(0, 256, (h, w))
    return np.stack(ch, 2).astype(np.uint8)


@pytest.mark.parametrize("h,w", [(376, 1241), (480, 752), (250, 333)])
@pytest.mark.parametrize("cn", [1, 3, 4])
@pytest.mark.parametrize("half", [False, True])
def test_extract_color_matches_oracle(h, w, cn, half):
    if half:
        h, w = 2 * h, 2 * w + (1 if w % 2 == 0 else 0)   # 4k+1 source widths: cvRound(w/2) = 2k (ties to even), the box path still applies
    img = _color(h, w, cn, 11 * cn + h)
    s = hb.FeatureExtractorSettings(nFeatures=1000)
    ex = hb.ORBExtractor(s)
    for rgb in (True, False):
        want_gray = O.preprocess(img, rgb, half)
        ok, od = O.extract(want_gray, O.default_params(s.nFeatures, s.fScaleFactor, s.nLevels, s.N_CELLS))
        gray, k, d = ex.extract_color(img, rgb=rgb, half_scale=half)
        assert np.array_equal(gray, want_gray), f"gray differs in {np.count_nonzero(gray != want_gray)} pixels"
        assert len(k) == len(ok) and k.tobytes() == ok.tobytes()
        assert np.array_equal(d, od)
        assert len(k) > 100
    ex.close()


def test_extract_color_strided_rows_and_errors():
    img = _color(240, 322, 3, 5)
    wide = np.zeros((240, 340, 3), np.uint8)
    wide[:, :322] = img
    view = wide[:, :322]                      # row stride 1020 bytes, 966 used
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=500))
    g1, k1, d1 = ex.extract_color(img)
    g2, k2, d2 = ex.extract_color(view)
    assert np.array_equal(g1, g2) and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    with pytest.raises(hb.HyorbError):
        ex.extract_color(np.zeros((35, 51, 3), np.uint8), half_scale=True)     # outside the 2x2 box path
    with pytest.raises(hb.HyorbError):
        ex.extract_color(np.zeros((64, 64, 2), np.uint8))
    g, k, d = ex.extract_color(np.zeros((0, 0, 3), np.uint8))
    assert len(k) == 0
    ex.close()


@pytest.mark.parametrize("aligned", [True, False])
def test_preprocess_device_batch_feeds_the_device_extractor(aligned):
    torch = pytest.importorskip("torch")
    B, H, W, cn = 4, 376, 1241, 3
    imgs = np.stack([_color(H, W, cn, 40 + i) for i in range(B)])
    s = hb.FeatureExtractorSettings(nFeatures=2000)
    ex = hb.ORBExtractor(s)
    want = [ex.extract_color(imgs[i], rgb=False) for i in range(B)]
    spitch = W * cn + (0 if aligned else 5)
    dsrc = torch.zeros((B, H, spitch), dtype=torch.uint8, device="cuda")
    dsrc[:, :, :W * cn] = torch.from_numpy(imgs.reshape(B, H, W * cn)).cuda()
    gpitch = (W + 15) & ~15 if aligned else W
    dgray = torch.zeros((B, H, gpitch), dtype=torch.uint8, device="cuda")
    cap = ex.default_capacity()
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    F.check(F.lib().hyorb_preprocess_device(ex._h, dsrc.data_ptr(), B, W, H, spitch, spitch * H, cn, 0, 0, dgray.data_ptr(), gpitch, gpitch * H))
    ex.extract_batch_device(dgray.data_ptr(), B, W, H, gpitch, gpitch * H, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_counts.data_ptr())
    ex.sync()
    torch.cuda.synchronize()
    g = dgray.cpu().numpy()
    k2 = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28)
    for i in range(B):
        gray, k, d = want[i]
        assert np.array_equal(g[i, :, :W], gray)
        n = int(d_counts[i])
        assert n == len(k) and k2[i, :n].tobytes() == k.tobytes()
        assert np.array_equal(d_desc[i, :n].cpu().numpy(), d)
    ex.close()
