"""ctypes binding of libhyorb.so (include/hyorb.h).  No torch types cross the boundary: numpy arrays for the
``_host`` entry points, raw device addresses (ints) for the ``_device`` ones.

The library is built in-tree (hyslam_b200/lib/libhyorb.so) by ``build()`` / ``__graft_entry__.build()``.
There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HYORB_LIB") or os.path.join(_HERE, "lib", "libhyorb.so")     # HYORB_LIB: A/B-test another build
CSRC = os.path.join(_HERE, "csrc")

OK, EINVAL, ECAPACITY, EUNSUPPORTED, ENOMEM, ECUDA = 0, -1, -2, -3, -4, -5
MAX_LEVELS = 16
GRID_COLS, GRID_ROWS = 64, 48
RULE_LANDMARK, RULE_BOW, RULE_MONOINIT = 0, 1, 2
SBP_DISTANCE, SBP_STEREO, SBP_ROTATION = 1, 2, 4
DBG_PYRAMID, DBG_BLURRED, DBG_CANDIDATES, DBG_LEVEL_COUNT = 0, 1, 2, 3

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
WQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("size_lo", "<f4"), ("size_hi", "<f4"),
                     ("ur", "<f4"), ("ur_radius", "<f4")])
assert KP_DTYPE.itemsize == 28 and WQ_DTYPE.itemsize == 28


class ExtractorParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32), ("cell_px", C.c_int32),
                ("ini_th", C.c_int32), ("min_th", C.c_int32), ("flags", C.c_uint32)]


class StereoParams(C.Structure):
    _fields_ = [("mbf", C.c_float), ("fx", C.c_float), ("n_rows", C.c_int32), ("th_high", C.c_float),
                ("th_low", C.c_float), ("size_ref", C.c_float)]


class Bounds(C.Structure):
    _fields_ = [("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


LM_DTYPE = np.dtype([("Pw", "<f4", 3), ("size", "<f4"), ("min_dist", "<f4"), ("max_dist", "<f4"), ("assoc_idx", "<i4")])
assert LM_DTYPE.itemsize == 28


class Projection(C.Structure):
    """hyorb_projection: pose + camera of the frame landmarks are projected into"""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("K", C.c_float * 9),
                ("mbf", C.c_float), ("stereo", C.c_int32), ("bounds", Bounds)]


class HyorbError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"hyorb rc={rc}: {msg}")
        self.rc = rc


# every symbol include/hyorb.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = [
    "hyorb_extractor_create", "hyorb_extractor_destroy", "hyorb_extractor_get_levels", "hyorb_extractor_get_scales",
    "hyorb_extract_host", "hyorb_extract_batch_host", "hyorb_extract_batch_device", "hyorb_extractor_sync",
    "hyorb_extractor_level_size", "hyorb_extractor_keypoint_bound", "hyorb_extractor_debug_read", "hyorb_extractor_launch_count",
    "hyorb_matcher_create", "hyorb_matcher_destroy", "hyorb_matcher_sync", "hyorb_matcher_launch_count",
    "hyorb_stereo_match_host", "hyorb_stereo_match_batch_device", "hyorb_match_csr_host",
    "hyorb_match_bruteforce_device", "hyorb_grid_build_host", "hyorb_match_window_host",
    "hyorb_rotation_consistency_host", "hyorb_last_error", "hyorb_version", "hyorb_device_count",
    "hyorb_process_stereo_batch_host", "hyorb_process_stereo_batch_device", "hyorb_extractor_set_profiling",
    "hyorb_extractor_stage_times", "hyorb_extractor_set_pipelining", "hyorb_distinctive_descriptor_host", "hyorb_project_landmarks_host", "hyorb_search_by_projection_host", "hyorb_search_by_projection_ex_host",
    "hyorb_vocabulary_create", "hyorb_vocabulary_destroy", "hyorb_bow_transform_host", "hyorb_search_by_bow_host",
    "hyorb_preprocess_size", "hyorb_preprocess_device", "hyorb_extract_color_host", "hyorb_search_for_triangulation_host", "hyorb_match_csr_epipolar_host",
    "hyorb_fuse_host", "hyorb_search_by_sim3_host", "hyorb_search_for_initialization_host",
]
N_STAGES = 6
STAGE_NAMES = ("pyramid", "fast", "quadtree", "blur", "describe", "stereo")


def build(force=False, verbose=False):
    """Compile libhyorb.so for sm_100a with nvcc (cross-compiles without a GPU).  Idempotent."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("hyorb.h", "hyorb_brief_pattern.inc")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libhyorb.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load the library (never builds implicitly on a GPU box: the .so ships with the tree)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.hyorb_last_error.restype = C.c_char_p
        L.hyorb_version.restype = C.c_char_p
        L.hyorb_extractor_debug_read.restype = C.c_long
        L.hyorb_extractor_launch_count.restype = C.c_long
        L.hyorb_matcher_launch_count.restype = C.c_long
        L.hyorb_extractor_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.hyorb_extract_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hyorb_preprocess_size.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.hyorb_preprocess_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_int, C.c_size_t]
        L.hyorb_extract_color_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hyorb_extract_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hyorb_extract_batch_device.argtypes = L.hyorb_extract_batch_host.argtypes
        L.hyorb_process_stereo_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_process_stereo_batch_device.argtypes = L.hyorb_process_stereo_batch_host.argtypes
        L.hyorb_extractor_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.hyorb_extractor_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hyorb_extractor_set_pipelining.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.hyorb_extractor_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hyorb_extractor_destroy.argtypes = [C.c_void_p]
        L.hyorb_extractor_sync.argtypes = [C.c_void_p]
        L.hyorb_extractor_get_levels.argtypes = [C.c_void_p]
        L.hyorb_extractor_launch_count.argtypes = [C.c_void_p]
        L.hyorb_extractor_get_scales.argtypes = [C.c_void_p] * 6
        L.hyorb_extractor_level_size.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.hyorb_extractor_keypoint_bound.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hyorb_matcher_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.hyorb_matcher_destroy.argtypes = [C.c_void_p]
        L.hyorb_matcher_sync.argtypes = [C.c_void_p]
        L.hyorb_matcher_launch_count.argtypes = [C.c_void_p]
        L.hyorb_stereo_match_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_stereo_match_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_match_csr_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_match_bruteforce_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_grid_build_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_match_window_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_rotation_consistency_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hyorb_distinctive_descriptor_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hyorb_search_by_projection_ex_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint,
                                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_vocabulary_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_vocabulary_destroy.argtypes = [C.c_void_p]
        L.hyorb_match_csr_epipolar_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p]
        L.hyorb_search_for_triangulation_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                          C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_bow_transform_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_search_by_bow_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                               C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_project_landmarks_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                                   C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.hyorb_search_by_projection_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_fuse_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hyorb_search_by_sim3_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p]
        L.hyorb_search_for_initialization_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, Bounds,
                                                           C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise HyorbError(rc, lib().hyorb_last_error().decode())
    return rc


def ptr(a):
    """address of a numpy array (or pass through None / int device pointers)"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def device_count():
    return lib().hyorb_device_count()
