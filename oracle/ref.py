"""ctypes wrapper over oracle/_ref/libhyslam_ref.so -- the REFERENCE's own translation units (test infrastructure).

The library holds hySLAM's ORBExtractor / ORBFinder / ORBDistance / FeatureDescriptor / Stereomatcher / FeatureViews
compiled unmodified from /root/reference against oracle/cvshim (see oracle/Makefile, oracle/ref_glue.cpp).  It is the
parity pin: tests compare the C oracle and the CUDA path against what the reference's code itself returns.
Only tests/ and bench.py's CPU legs may import this module.  /root/reference exists only in the build container; on the
GPU box the prebuilt library travels with the snapshot (oracle/_ref is git-ignored, not gpurun-ignored).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import KP_DTYPE, Params, StereoParams, default_params, level_sizes

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libhyslam_ref.so")
REFERENCE_ROOT = os.environ.get("HYSLAM_REFERENCE", "/root/reference")


def available():
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "features"))


def build(force=False):
    """Compile oracle/_ref from the reference tree when it is present (idempotent); otherwise use the prebuilt file."""
    have_src = os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "features"))
    if not have_src:
        if os.path.exists(_SO):
            return _SO
        raise RuntimeError("oracle/_ref is not built and the reference tree is absent")
    deps = [os.path.join(_HERE, f) for f in ("ref_glue.cpp", "ref_glue_match.cpp", "ref_scene.hpp", "Makefile", "cvshim/cvshim.cpp", "cvshim/opencv2/core/core.hpp")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in deps):
        return _SO
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", _HERE, "ref", "CXX=g++", f"REF={REFERENCE_ROOT}"], check=True, capture_output=True, env=env)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ref_hamming.restype = C.c_float
        _lib.cvshim_fast_atan2.restype = C.c_float
        _lib.cvshim_norm.restype = C.c_double
        _lib.ref_process_stereo_pairs.restype = C.c_long
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def extract(img, p=None, cap=None, arena=True, levels=False):
    """HYSLAM::ORBExtractor::operator() on one 8-bit frame.  arena=True: monotonic allocator (canonical quadtree tie
    policy); arena=False: glibc malloc, like a stock hySLAM build."""
    p = p or default_params()
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    cap = cap or max(4 * p.nfeatures + 1024, 4096)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int32()
    lv = None
    ptrs = None
    if levels:
        sizes = level_sizes(p, W, H)
        lv = [np.zeros((h, w), np.uint8) for (w, h) in sizes]
        ptrs = (C.c_void_p * p.nlevels)(*[a.ctypes.data for a in lv])
    rc = lib().ref_extract(C.byref(p), _p(img), W, H, img.strides[0], int(bool(arena)), _p(kps), _p(desc), cap, C.byref(n), ptrs)
    if rc != 0:
        raise RuntimeError(f"ref_extract rc={rc} n={n.value}")
    k, d = kps[:n.value].copy(), desc[:n.value].copy()
    return (k, d, lv) if levels else (k, d)


def scale_tables(p):
    n = p.nlevels
    s, i, s2, i2 = (np.zeros(n, np.float32) for _ in range(4))
    lib().ref_scale_tables(C.byref(p), _p(s), _p(i), _p(s2), _p(i2))
    return s, i, s2, i2


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return float(lib().ref_hamming(_p(a), _p(b)))


def bf_scan(q, t, lean=False):
    """brute-force best / second-best scan on the CPU: the reference's own FeatureDescriptor::distance per pair (lean=False) or contiguous
    descriptors + hardware popcount (lean=True).  Returns (best_idx, best, second)."""
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    bi = np.empty(len(q), np.int32); b = np.empty(len(q), np.int32); s2 = np.empty(len(q), np.int32)
    L = lib()
    L.ref_bf_scan.restype = C.c_long
    L.ref_bf_scan(_p(q), len(q), _p(t), len(t), int(bool(lean)), _p(bi), _p(b), _p(s2))
    return bi, b, s2


def stereo_match(sp, kl, dl, kr, dr):
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    uR = np.empty(len(kl), np.float32); depth = np.empty(len(kl), np.float32)
    rc = lib().ref_stereo_match(C.byref(sp), _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(uR), _p(depth))
    if rc < 0:
        raise RuntimeError(f"ref_stereo_match rc={rc}")
    return uR, depth


def process_stereo_pairs(p, sp, left, right, threads_per_pair=2):
    """n pairs through extract L + R + Stereomatcher, in the reference's threading shape.  Returns (keypoints, matches)."""
    left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
    n, H, W = left.shape
    m = C.c_long()
    tot = lib().ref_process_stereo_pairs(C.byref(p), C.byref(sp) if sp is not None else None, _p(left), _p(right), n, W, H,
                                         int(threads_per_pair), C.byref(m))
    return int(tot), int(m.value)


# ---- cvshim primitives (pinned against cv2 in tests/test_cvshim_vs_cv2.py) ----
def shim_fast(img, threshold=20, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 2 + 16
    x, y, r = (np.empty(cap, np.float32) for _ in range(3))
    n = lib().cvshim_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(x), _p(y), _p(r), cap)
    assert n >= 0, n
    return x[:n].copy(), y[:n].copy(), r[:n].copy()


def shim_resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().cvshim_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def shim_blur(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().cvshim_blur(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def shim_border(src, b, border_type=4):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    lib().cvshim_border(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0], b, border_type)
    return dst


def shim_fast_atan2(y, x):
    return float(lib().cvshim_fast_atan2(C.c_float(y), C.c_float(x)))


def shim_gemm(A, B, Cm=None, ta=False):
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    m = A.shape[1] if ta else A.shape[0]
    k = A.shape[0] if ta else A.shape[1]
    n = B.shape[1]
    D = np.empty((m, n), np.float32)
    Cc = None if Cm is None else np.ascontiguousarray(Cm, np.float32)
    lib().cvshim_gemm(_p(A), m, k, int(ta), _p(B), n, _p(Cc), _p(D))
    return D


def shim_norm(a):
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    return float(lib().cvshim_norm(_p(a), len(a)))


# ---------------------------------------------------------------------------------------------------------------------
# matcher side: the reference's FeatureMatcher / MatchCriteria / Frame / KeyFrame / MapPoint (oracle/ref_glue_match.cpp)
# ---------------------------------------------------------------------------------------------------------------------
class FrameDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("kps", C.c_void_p), ("desc", C.c_void_p), ("uR", C.c_void_p), ("depth", C.c_void_p),
                ("K", C.c_float * 9), ("mbf", C.c_float), ("sensor", C.c_int32),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("Tcw", C.c_float * 16), ("size_ref", C.c_float), ("sigma_ref", C.c_float)]


class MatcherSettings(C.Structure):
    _fields_ = [("nnratio", C.c_float), ("th_high", C.c_float), ("th_low", C.c_float), ("check_ori", C.c_int32)]


def settings(nnratio=0.6, th_high=100.0, th_low=50.0, check_ori=True):
    return MatcherSettings(nnratio, th_high, th_low, int(check_ori))


class Scene:
    """MapPoints + Frames / KeyFrames built from flat arrays inside the reference's own classes."""

    def __init__(self, max_mappoints, matcher_lib=None, prefix="refm_"):
        """matcher_lib / prefix: which library's matcher entry points drive this scene -- oracle/_ref's refm_* (the reference's
        FeatureMatcher, default) or tests/cpp/_build/libmatcher_shim_test.so's shimm_* (the C++ drop-in CudaFeatureMatcher)."""
        L = lib()
        L.refm_scene_create.restype = C.c_void_p
        self.h = C.c_void_p(L.refm_scene_create(int(max_mappoints)))
        self._keep = []
        self._mlib = matcher_lib or L
        self._prefix = prefix

    def _m(self, name):
        return getattr(self._mlib, self._prefix + name)

    def close(self):
        if self.h:
            lib().refm_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add_mappoints(self, Pw, desc, normal=None, size=None, min_dist=None, max_dist=None, bad=None, n_protected=None):
        Pw = np.ascontiguousarray(Pw, np.float32).reshape(-1, 3)
        n = len(Pw)
        f = lambda a, dt=np.float32: None if a is None else np.ascontiguousarray(a, dt)
        normal, size, min_dist, max_dist = f(normal), f(size), f(min_dist), f(max_dist)
        desc = np.ascontiguousarray(desc, np.uint8)
        first = lib().refm_add_mappoints(self.h, n, _p(Pw), _p(normal), _p(size), _p(min_dist), _p(max_dist), _p(desc), _p(f(bad, np.uint8)),
                                         _p(f(n_protected, np.int32)))
        if first < 0:
            raise RuntimeError("scene capacity exceeded")
        return first

    def add_frame(self, kps, desc, K, Tcw, bounds, mbf=0.0, stereo=False, uR=None, depth=None, assoc=None, keyframe=False, size_ref=31.0, sigma_ref=1.0):
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        d = FrameDesc()
        d.n = len(kps); d.kps = kps.ctypes.data; d.desc = desc.ctypes.data
        if uR is not None:
            uR = np.ascontiguousarray(uR, np.float32)
            depth = np.ascontiguousarray(depth if depth is not None else np.where(uR >= 0, 1.0, -1.0), np.float32)
            d.uR = uR.ctypes.data; d.depth = depth.ctypes.data
        d.K[:] = [float(v) for v in np.asarray(K, np.float32).reshape(9)]
        d.Tcw[:] = [float(v) for v in np.asarray(Tcw, np.float32).reshape(16)]
        d.mbf = float(mbf); d.sensor = 1 if stereo else 0
        d.min_x, d.max_x, d.min_y, d.max_y = [float(v) for v in bounds]
        d.size_ref = float(size_ref); d.sigma_ref = float(sigma_ref)
        a = None if assoc is None else np.ascontiguousarray(assoc, np.int32)
        self._keep += [kps, desc, uR, depth, a]
        fid = lib().refm_add_frame(self.h, C.byref(d), _p(a), int(keyframe))
        if fid < 0:
            raise RuntimeError("refm_add_frame failed")
        return fid

    def set_feature_nodes(self, frame, node_of):
        a = np.ascontiguousarray(node_of, np.int32)
        lib().refm_set_feature_nodes(self.h, frame, _p(a), len(a))

    def set_observation_count(self, mp, n):
        lib().refm_set_observation_count(self.h, int(mp), int(n))

    def assoc(self, frame, n):
        out = np.empty(n, np.int32)
        lib().refm_get_assoc(self.h, frame, _p(out), n)
        return out

    def features_in_area(self, frame, x, y, r, cap=20000):
        out = np.empty(cap, np.int32)
        n = lib().refm_features_in_area(self.h, frame, C.c_float(x), C.c_float(y), C.c_float(r), _p(out), cap)
        assert n >= 0
        return out[:n].copy()

    def camera_center(self, frame):
        o = np.empty(3, np.float32)
        lib().refm_camera_center(self.h, frame, _p(o))
        return o

    def project(self, frame, mp):
        uv = np.empty(3, np.float32); sz = C.c_float()
        ok = lib().refm_project(self.h, frame, int(mp), _p(uv), C.byref(sz))
        return bool(ok), uv, np.float32(sz.value)

    def search_by_projection(self, frame, lm_ids, th, st):
        a = np.ascontiguousarray(lm_ids, np.int32)
        return self._m("search_by_projection")(self.h, frame, _p(a), len(a), C.c_float(th), C.byref(st))

    def search_by_projection_motion(self, cur, last, th, st, mono=False):
        return self._m("search_by_projection_motion")(self.h, cur, last, C.c_float(th), int(mono), C.byref(st))

    def search_by_projection_reloc(self, cur, kf, found, th, orb_dist, st):
        a = np.ascontiguousarray(found, np.int32)
        return self._m("search_by_projection_reloc")(self.h, cur, kf, _p(a), len(a), C.c_float(th), int(orb_dist), C.byref(st))

    def fuse(self, kf, lm_ids, th, reproj_err, st):
        a = np.ascontiguousarray(lm_ids, np.int32)
        oi = np.empty(len(a) + 1, np.int32); ol = np.empty(len(a) + 1, np.int32)
        n = self._m("fuse")(self.h, kf, _p(a), len(a), C.c_float(th), C.c_float(reproj_err), C.byref(st), _p(oi), _p(ol), len(oi))
        assert n >= 0
        return oi[:n].copy(), ol[:n].copy()

    def search_for_initialization(self, f1, f2, prev_matched, window, st):
        pm = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.empty(len(pm), np.int32)
        n = self._m("search_for_initialization")(self.h, f1, f2, _p(pm), _p(m12), int(window), C.byref(st))
        return n, m12, pm

    def search_by_sim3(self, kf1, kf2, matches12, s12, R12, t12, th, st):
        m = np.ascontiguousarray(matches12, np.int32).copy()
        R = np.ascontiguousarray(R12, np.float32); t = np.ascontiguousarray(t12, np.float32)
        n = self._m("search_by_sim3")(self.h, kf1, kf2, _p(m), len(m), C.c_float(s12), _p(R), _p(t), C.c_float(th), C.byref(st))
        return n, m

    def search_for_triangulation(self, kf1, kf2, F12, only_stereo, st, cap=20000):
        Fm = np.ascontiguousarray(F12, np.float32)
        a = np.empty(cap, np.int32); b = np.empty(cap, np.int32)
        n = self._m("search_for_triangulation")(self.h, kf1, kf2, _p(Fm), int(only_stereo), C.byref(st), _p(a), _p(b), cap)
        assert n >= 0
        return a[:n].copy(), b[:n].copy()

    def search_by_bow(self, kf, frame, st, cap=20000):
        a = np.empty(cap, np.int32); b = np.empty(cap, np.int32)
        n = self._m("search_by_bow")(self.h, kf, frame, C.byref(st), _p(a), _p(b), cap)
        assert n >= 0
        return a[:n].copy(), b[:n].copy()
