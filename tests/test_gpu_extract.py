"""GPU parity of the extraction path (pyramid, FAST + cell-local NMS, quadtree distribution, blur, orientation,
rBRIEF) through the C ABI against the CPU oracle and the committed golden fixtures.  Bit-exact, exact ORDER."""
import os

import numpy as np
import pytest

import hyslam_b200 as hb
from hyslam_b200 import _ffi as F, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _settings(nfeatures=1000, scale=1.2, nlevels=8, cell=30):
    return hb.FeatureExtractorSettings(nFeatures=nfeatures, fScaleFactor=scale, nLevels=nlevels, N_CELLS=cell)


def _oparams(s):
    return O.default_params(s.nFeatures, s.fScaleFactor, s.nLevels, s.N_CELLS)


def _bits(a):
    return a.view(np.uint32) if a.dtype.kind == "f" else a


def assert_kps_equal(k, gk, what=""):
    assert len(k) == len(gk), f"{what}: {len(k)} keypoints vs {len(gk)} expected"
    for f in gk.dtype.names:
        bad = np.nonzero(_bits(np.ascontiguousarray(k[f])) != _bits(np.ascontiguousarray(gk[f])))[0]
        assert len(bad) == 0, f"{what}: field {f} differs at {bad[:8]} ({len(bad)} of {len(k)}): got {k[f][bad[:4]]} want {gk[f][bad[:4]]}"


CASES = [
    ("noise", 240, 320, 3, 500),
    ("blocks", 280, 376, 5, 700),
    ("noise", 480, 752, 0, 1000),      # C1
    ("blocks", 480, 752, 1, 1000),     # C1, sparse corners: empty cells, quadtree early exits
    ("noise", 376, 1241, 2, 2000),     # C2
    ("blocks", 376, 1241, 4, 2000),
]


@pytest.mark.parametrize("fused", [0, 1])      # pyramid.cu + blur.cu / level.cu (one pass per level produces both)
@pytest.mark.parametrize("kind,h,w,seed,nf", CASES)
def test_extract_stages_match_oracle(kind, h, w, seed, nf, fused, monkeypatch):
    monkeypatch.setenv("HYORB_FUSED_LEVELS", str(fused))
    img = (synth.noise_image if kind == "noise" else synth.blocks_image)(h, w, seed)
    s = _settings(nf)
    ok, od, info = O.extract(img, _oparams(s), debug=True)
    ex = hb.ORBExtractor(s)
    k, d = ex(img, None)
    errs = []
    for l in range(s.nLevels):
        if not np.array_equal(ex.debug_level(w, h, l, F.DBG_PYRAMID), info["pyramid"][l]):
            errs.append(f"pyramid level {l}")
    for l in range(s.nLevels):
        cx, cy, cr = info["cand"][l]
        want = np.stack([cx, cy, cr], 1).astype(np.int32)
        got = ex.debug_candidates(w, h, l)
        if got.shape != want.shape or not np.array_equal(got, want):
            errs.append(f"FAST candidates level {l}: got {len(got)} want {len(want)}")
    for l in range(s.nLevels):
        if ex.debug_level_count(l) != info["level_count"][l]:
            errs.append(f"quadtree count level {l}: got {ex.debug_level_count(l)} want {info['level_count'][l]}")
    for l in range(s.nLevels):
        if info["level_count"][l] and not np.array_equal(ex.debug_level(w, h, l, F.DBG_BLURRED), info["blurred"][l]):
            errs.append(f"blur level {l}")
    assert not errs, errs
    assert_kps_equal(k, ok, "keypoints")
    assert np.array_equal(d, od), f"descriptors differ in {np.count_nonzero((d != od).any(1))} rows"
    ex.close()


@pytest.mark.parametrize("name", ["ext_noise_320x240", "ext_blocks_376x280", "ext_blocks_640x360", "ext_c1_noise_752x480"])
def test_extract_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    fn = synth.noise_image if str(g["kind"]) == "noise" else synth.blocks_image
    img = fn(int(g["h"]), int(g["w"]), int(g["seed"]))
    s = _settings(int(g["nfeatures"]), float(g["scale_factor"]), int(g["nlevels"]), int(g["cell_px"]))
    ex = hb.ORBExtractor(s)
    k, d = ex(img, None)
    assert [ex.debug_level_count(l) for l in range(s.nLevels)] == g["level_count"].tolist()
    assert_kps_equal(k, g["kps"], name)
    assert np.array_equal(d, g["desc"])


def test_c3_full_size_matches_oracle():
    """C3: 3840x2160, 8000 features (oracle takes a few seconds)."""
    img = synth.blocks_image(2160, 3840, 11)
    img[400:1400, 600:2600] = synth.noise_image(1000, 2000, 12)      # a dense region on a sparse canvas
    s = _settings(8000)
    ok, od = O.extract(img, _oparams(s), cap=40000)
    ex = hb.ORBExtractor(s)
    k, d = ex(img, None, capacity=40000)
    assert_kps_equal(k, ok, "C3")
    assert np.array_equal(d, od)


def test_batch_equals_single_and_is_deterministic():
    imgs = np.stack([synth.noise_image(376, 1241, 20 + i) if i % 2 == 0 else synth.blocks_image(376, 1241, 20 + i) for i in range(6)])
    s = _settings(2000)
    ex = hb.ORBExtractor(s)
    kps, desc, counts = ex.extract_batch(imgs)
    kps2, desc2, counts2 = ex.extract_batch(imgs)
    assert np.array_equal(counts, counts2)
    for i in range(len(imgs)):
        k1, d1 = ex(imgs[i], None)
        n = counts[i]
        assert n == len(k1)
        assert_kps_equal(kps[i, :n], k1, f"image {i}")
        assert np.array_equal(desc[i, :n], d1)
        assert np.array_equal(kps[i, :n], kps2[i, :n]) and np.array_equal(desc[i, :n], desc2[i, :n])
    ok, od = O.extract(imgs[3], _oparams(s))
    assert_kps_equal(kps[3, :counts[3]], ok, "batch image 3 vs oracle")
    assert np.array_equal(desc[3, :counts[3]], od)


def test_strided_input_and_other_params():
    big = synth.noise_image(300, 500, 9)
    view = big[10:250, 17:417]                        # non-contiguous rows, odd offset
    s = _settings(600, scale=1.3, nlevels=5, cell=24)
    ok, od = O.extract(np.ascontiguousarray(view), _oparams(s))
    ex = hb.ORBExtractor(s)
    k, d = ex(view, None)
    assert_kps_equal(k, ok, "strided")
    assert np.array_equal(d, od)
    assert np.array_equal(ex.GetScaleFactors(), O.scale_tables(_oparams(s))[0])
    assert ex.features_per_level().tolist() == O.scale_tables(_oparams(s))[4].tolist()


def test_flat_image_gives_no_keypoints_and_empty_image_returns_silently():
    ex = hb.ORBExtractor(_settings(500))
    k, d = ex(np.full((240, 320), 77, np.uint8), None)
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.zeros((0, 0), np.uint8), None)
    assert len(k) == 0


def test_error_behaviour():
    ex = hb.ORBExtractor(_settings(1000))
    img = synth.noise_image(480, 752, 0)
    with pytest.raises(hb.HyorbError) as e:
        ex(img, None, capacity=100)                   # output capacity overflow is reported, not truncated
    assert e.value.rc == F.ECAPACITY
    k, _ = ex(img, None)                              # the handle stays usable
    assert len(k) > 900
    with pytest.raises(hb.HyorbError) as e:
        ex(synth.noise_image(100, 120, 1), None)      # level 7 smaller than one FAST cell: the reference divides by zero
    assert e.value.rc == F.EUNSUPPORTED
    with pytest.raises(ValueError):
        ex(img.astype(np.float32), None)


@pytest.mark.parametrize("pitched", [False, True])
def test_device_pointer_entry_matches_host_entry(pitched):
    """hyorb_extract_batch_device on a dense odd-pitch device batch (repacked on the device to a TMA-addressable layout) and on
    a pitched one (read in place through TMA) gives what the host entry gives."""
    torch = pytest.importorskip("torch")
    imgs = np.stack([synth.noise_image(376, 1241, 70 + i) if i % 2 else synth.blocks_image(376, 1241, 70 + i) for i in range(4)])
    s = _settings(2000)
    ex = hb.ORBExtractor(s)
    kps, desc, counts = ex.extract_batch(imgs)
    B, H, W = imgs.shape
    cap = ex.default_capacity()
    pitch = (W + 15) & ~15 if pitched else W
    dev = torch.zeros((B, H, pitch), dtype=torch.uint8, device="cuda")
    dev[:, :, :W] = torch.from_numpy(imgs).cuda()
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_batch_device(dev.data_ptr(), B, W, H, pitch, pitch * H, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_counts.data_ptr())
    ex.sync()
    torch.cuda.synchronize()
    c2 = d_counts.cpu().numpy()
    assert np.array_equal(c2, counts)
    k2 = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28)
    for i in range(B):
        n = counts[i]
        assert k2[i, :n].tobytes() == kps[i, :n].tobytes()
        assert np.array_equal(d_desc[i, :n].cpu().numpy(), desc[i, :n])


def test_large_host_batch_fetches_what_exceeds_the_download_bound(monkeypatch):
    """Host batches of more than 8 images download min(capacity, keypoint bound) entries per image while the batch runs and fetch the
    rest after the counts have arrived (api.cu ex_fetch_overflow).  With the bound forced far below what the images produce, the
    results must still be complete and identical to the single-image calls."""
    imgs = np.stack([synth.noise_image(240, 320, 100 + i) for i in range(12)])
    s = _settings(600)
    ref = hb.ORBExtractor(s)
    want = [ref(im, None) for im in imgs]
    ref.close()
    monkeypatch.setenv("HYORB_KP_BOUND", "100")
    ex = hb.ORBExtractor(s)
    assert ex.keypoint_bound(320, 240) == 100
    kps, desc, counts = ex.extract_batch(imgs, capacity=1024)
    for i, (k, d) in enumerate(want):
        assert counts[i] == len(k) and len(k) > 300
        assert_kps_equal(kps[i, :counts[i]], k, f"image {i}")
        assert np.array_equal(desc[i, :counts[i]], d)
    ex.close()


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("nl,sf,h,w,nf", [(1, 1.2, 240, 320, 300),      # a single level: nothing to resize
                                          (3, 2.0, 480, 640, 500),      # exact 2x: cv::resize's INTER_AREA shortcut (k_resize) between fused levels
                                          (4, 1.05, 300, 400, 400),     # more than 32 destination groups per 128-pixel tile column
                                          (2, 1.999, 400, 600, 300)])   # the widest tap spread the resize supports
def test_level_kernel_edge_pyramids(nl, sf, h, w, nf, fused, monkeypatch):
    monkeypatch.setenv("HYORB_FUSED_LEVELS", str(fused))
    img = synth.noise_image(h, w, 5)
    s = hb.FeatureExtractorSettings(nFeatures=nf, fScaleFactor=sf, nLevels=nl)
    ok, od = O.extract(img, O.default_params(nf, sf, nl, 30))
    ex = hb.ORBExtractor(s)
    k, d = ex(img, None)
    assert_kps_equal(k, ok, "keypoints")
    assert np.array_equal(d, od)
    ex.close()


@pytest.mark.parametrize("name", ["binary_noise", "checker2", "dots4", "sparse_dots"])
def test_adversarial_textures_match_oracle(name):
    """High-frequency synthetic textures at C2 size: up to 26k NMS survivors on one level (5.5 % of its pixels), levels with no corner
    at all, dense isolated peaks -- the candidate capacity, the per-warp survivor staging of k_fast and the quadtree's early exits."""
    h, w = 376, 1241
    rng = np.random.default_rng(3)
    yy, xx = np.arange(h)[:, None], np.arange(w)[None, :]
    img = {"binary_noise": lambda: (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8),
           "checker2": lambda: (((yy // 2 + xx // 2) % 2) * 255).astype(np.uint8),
           "dots4": lambda: np.where((yy % 4 == 0) & (xx % 4 == 0), 255, 0).astype(np.uint8),
           "sparse_dots": lambda: np.where(rng.random((h, w)) < 0.03, 255, 30).astype(np.uint8)}[name]()
    ok, od = O.extract(img, _oparams(_settings(2000)))
    ex = hb.ORBExtractor(_settings(2000))
    k, d = ex(img, None)
    assert_kps_equal(k, ok, name)
    assert np.array_equal(d, od)
    ex.close()


def test_pitched_host_batch_equals_dense_host_batch():
    """Host frames with a 16-byte-aligned row pitch are read in place through TMA (no device-side repack); dense odd-width rows are
    repacked on the device.  Both layouts of the same frames must give the same bytes (large-batch path, > 8 images)."""
    imgs = np.stack([synth.noise_image(200, 333, 60 + i) for i in range(10)])
    pitch = (333 + 15) & ~15
    store = np.zeros((10, 200, pitch), np.uint8)
    store[:, :, :333] = imgs
    ex = hb.ORBExtractor(_settings(500, nlevels=5))
    kd, dd, cd = ex.extract_batch(imgs, capacity=1024)
    kp, dp, cp = ex.extract_batch(store[:, :, :333], capacity=1024)
    assert np.array_equal(cd, cp) and cd.min() > 100
    for i in range(10):
        assert_kps_equal(kp[i, :cp[i]], kd[i, :cd[i]], f"image {i}")
        assert np.array_equal(dp[i, :cp[i]], dd[i, :cd[i]])
    ex.close()
