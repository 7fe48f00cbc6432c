"""Host-side mirror of HYSLAM::ORBExtractor (src/features/ORBExtractor.h:63-116) over the C ABI."""
import ctypes as C

import numpy as np

from . import _ffi as F
from .settings import FeatureExtractorSettings


class ORBExtractor:
    """Drop-in for the reference's extractor object: ``kps, desc = extractor(image, mask)``.

    One object == one mutable workspace + CUDA stream, like the reference (its pyramid is a member,
    ORBExtractor.h:102); use one object per thread.  The mask is ignored, as in the reference (ORBExtractor.h:75).
    """

    def __init__(self, settings=None, device=0, stream=None):
        s = settings or FeatureExtractorSettings()
        self.settings = s
        self._p = F.ExtractorParams(int(s.nFeatures), float(s.fScaleFactor), int(s.nLevels), int(s.N_CELLS),
                                    int(s.init_threshold), int(s.min_threshold), 0)
        self._h = C.c_void_p()
        F.check(F.lib().hyorb_extractor_create(C.byref(self._p), int(device), stream, C.byref(self._h)))
        n = s.nLevels
        self._scale, self._inv, self._s2, self._is2 = (np.zeros(n, np.float32) for _ in range(4))
        self._quota = np.zeros(n, np.int32)
        F.check(F.lib().hyorb_extractor_get_scales(self._h, F.ptr(self._scale), F.ptr(self._inv), F.ptr(self._s2),
                                                   F.ptr(self._is2), F.ptr(self._quota)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            F.lib().hyorb_extractor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter teardown: module globals may already be gone
            pass

    # ---- FeatureExtractor interface (src/features/FeatureExtractor.h:25-37)
    def GetLevels(self):
        return F.lib().hyorb_extractor_get_levels(self._h)

    def GetScaleFactor(self):
        return float(np.float32(self.settings.fScaleFactor))

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv.copy()

    def GetScaleSigmaSquares(self):
        return self._s2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._is2.copy()

    def features_per_level(self):
        return self._quota.copy()

    def default_capacity(self):
        return int(4 * self.settings.nFeatures + 1024)

    def __call__(self, image, mask=None, capacity=None):
        """ORBExtractor::operator() (ORBExtractor.cpp:496-562).  image: HxW uint8.  Returns (kps[KP_DTYPE], desc[n,32])."""
        if image is None or image.size == 0:
            return np.zeros(0, F.KP_DTYPE), np.zeros((0, 32), np.uint8)     # :499-500
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("image must be 8-bit single channel (assert at ORBExtractor.cpp:503)")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = capacity or self.default_capacity()
        kps = np.empty(cap, F.KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        H, W = image.shape
        F.check(F.lib().hyorb_extract_host(self._h, F.ptr(image), W, H, image.strides[0], F.ptr(kps), F.ptr(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_color(self, image, rgb=True, half_scale=False, capacity=None):
        """ImageProcessing::PreProcessImg (ImageProcessing.cpp:118-138: scale 1.0 / 0.5, RGB|BGR[A] -> gray) + operator() on one camera
        frame (HxW or HxWxC uint8).  Returns (gray, kps, desc); gray is the frame the reference keeps as track_data.image."""
        if image is None or image.size == 0:
            return np.zeros((0, 0), np.uint8), np.zeros(0, F.KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim not in (2, 3):
            raise ValueError("image must be 8-bit with 1, 3 or 4 interleaved channels")
        cn = 1 if image.ndim == 2 else image.shape[2]
        if image.strides[-1] != 1 or (image.ndim == 3 and image.strides[1] != cn):
            image = np.ascontiguousarray(image)
        H, W = image.shape[:2]
        gw, gh = C.c_int(), C.c_int()
        F.check(F.lib().hyorb_preprocess_size(W, H, int(half_scale), C.byref(gw), C.byref(gh)))
        gray = np.empty((gh.value, gw.value), np.uint8)
        cap = capacity or self.default_capacity()
        kps = np.empty(cap, F.KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        F.check(F.lib().hyorb_extract_color_host(self._h, F.ptr(image), W, H, image.strides[0], cn, int(rgb), int(half_scale), F.ptr(gray),
                                                 gray.strides[0], F.ptr(kps), F.ptr(desc), cap, C.byref(n)))
        return gray, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, capacity=None):
        """Throughput form over a [B,H,W] uint8 host array.  Returns (kps[B,cap], desc[B,cap,32], counts[B])."""
        if images.dtype != np.uint8 or images.ndim != 3 or images.strides[2] != 1 or images.strides[1] < images.shape[2]:
            images = np.ascontiguousarray(images, np.uint8)      # pitched views (rows contiguous, any row / image stride) go through as they are
        B, H, W = images.shape
        cap = capacity or self.default_capacity()
        kps = np.empty((B, cap), F.KP_DTYPE)
        desc = np.empty((B, cap, 32), np.uint8)
        counts = np.zeros(B, np.int32)
        F.check(F.lib().hyorb_extract_batch_host(self._h, F.ptr(images), B, W, H, images.strides[1], images.strides[0], F.ptr(kps), F.ptr(desc), cap,
                                                 F.ptr(counts)))
        return kps, desc, counts

    def extract_batch_device(self, d_images, B, W, H, stride, image_stride, d_kps, d_desc, capacity, d_counts):
        """Device-pointer form (ints); asynchronous on the handle's stream -- call sync() to collect errors."""
        F.check(F.lib().hyorb_extract_batch_device(self._h, d_images, B, W, H, stride, image_stride, d_kps, d_desc, capacity, d_counts))

    def sync(self):
        F.check(F.lib().hyorb_extractor_sync(self._h))

    # ---- ImageProcessing::ProcessStereoImage (src/main/ImageProcessing.cpp:69-116) over a batch of pairs
    @staticmethod
    def stereo_params(camera, matcher_settings=None, size_ref=31.0):
        from .settings import FeatureMatcherSettings
        ms = matcher_settings or FeatureMatcherSettings()
        return F.StereoParams(float(camera.mbf), float(camera.fx), int(camera.mnMaxY), float(ms.TH_HIGH), float(ms.TH_LOW), float(size_ref))

    def process_stereo_batch(self, images, camera, capacity=None, out=None):
        """images: [2P,H,W] uint8 host array, (left, right) interleaved.  Returns (kps[2P,cap], desc[2P,cap,32], counts[2P],
        uR[P,cap], depth[P,cap]).  `out` may hold preallocated (e.g. pinned) arrays of those shapes."""
        B, H, W = images.shape
        P = B // 2
        cap = capacity or self.default_capacity()
        if out is None:
            out = (np.empty((B, cap), F.KP_DTYPE), np.empty((B, cap, 32), np.uint8), np.zeros(B, np.int32),
                   np.empty((P, cap), np.float32), np.empty((P, cap), np.float32))
        kps, desc, counts, uR, depth = out
        sp = self.stereo_params(camera)
        F.check(F.lib().hyorb_process_stereo_batch_host(self._h, C.byref(sp), F.ptr(images), P, W, H, images.strides[1], images.strides[0],
                                                        F.ptr(kps), F.ptr(desc), cap, F.ptr(counts), F.ptr(uR), F.ptr(depth)))
        return out

    def process_stereo_batch_device(self, sp, d_images, P, W, H, stride, image_stride, d_kps, d_desc, capacity, d_counts, d_uR, d_depth):
        F.check(F.lib().hyorb_process_stereo_batch_device(self._h, C.byref(sp), d_images, P, W, H, stride, image_stride, d_kps, d_desc,
                                                          capacity, d_counts, d_uR, d_depth))

    def set_profiling(self, on=True):
        F.check(F.lib().hyorb_extractor_set_profiling(self._h, int(on)))

    def set_pipelining(self, device_lanes=-1, host_lanes=-1, side_blur=-1):
        """Sub-batch lanes / side-stream blur of the batch entry points (-1 keeps a setting); lanes=1, side_blur=0 serialises the kernels."""
        F.check(F.lib().hyorb_extractor_set_pipelining(self._h, int(device_lanes), int(host_lanes), int(side_blur)))

    def stage_times(self, reset=True):
        """{stage: accumulated ms} and the number of profiled calls since the last reset (synchronises)."""
        ms = (C.c_double * F.N_STAGES)()
        calls = C.c_long(0)
        F.check(F.lib().hyorb_extractor_stage_times(self._h, ms, C.byref(calls), int(reset)))
        return dict(zip(F.STAGE_NAMES, list(ms))), calls.value

    def launch_count(self):
        return int(F.lib().hyorb_extractor_launch_count(self._h))

    # ---- stage outputs of the last call (parity tests)
    def level_size(self, W, H, level):
        w, h = C.c_int(), C.c_int()
        F.check(F.lib().hyorb_extractor_level_size(self._h, W, H, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def keypoint_bound(self, W, H):
        """most keypoints one WxH image can produce with these quotas (= entries per image a large host batch downloads)"""
        n = F.lib().hyorb_extractor_keypoint_bound(self._h, W, H)
        if n < 0:
            F.check(n)
        return n

    def debug_level(self, W, H, level, what=F.DBG_PYRAMID, image_index=0):
        w, h = self.level_size(W, H, level)
        out = np.empty((h, w), np.uint8)
        n = F.lib().hyorb_extractor_debug_read(self._h, image_index, what, level, F.ptr(out), out.nbytes)
        if n < 0:
            F.check(int(n))
        return out

    def debug_candidates(self, W, H, level, image_index=0):
        w, h = self.level_size(W, H, level)
        out = np.empty((w * h // 8 + 1024, 3), np.int32)
        n = F.lib().hyorb_extractor_debug_read(self._h, image_index, F.DBG_CANDIDATES, level, F.ptr(out), out.nbytes)
        if n < 0:
            F.check(int(n))
        return out[: n // 12].copy()

    def debug_level_count(self, level, image_index=0):
        out = np.zeros(1, np.int32)
        n = F.lib().hyorb_extractor_debug_read(self._h, image_index, F.DBG_LEVEL_COUNT, level, F.ptr(out), 4)
        if n < 0:
            F.check(int(n))
        return int(out[0])
