"""CPU checks of the drop-in boundary: libhyorb.so builds for sm_100a, loads, exports exactly the symbols
include/hyorb.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from hyslam_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    F.build()
    return F.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "hyorb.h")).read()
    return sorted(set(re.findall(r"HYORB_API\s+[\w\s\*]+?\b(hyorb_\w+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    declared = _declared()
    assert declared == sorted(F.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", F.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\b(hyorb_\w+)$", out, re.M)))
    assert exported == declared              # nothing missing, nothing extra (visibility=hidden for the rest)
    for s in declared:
        assert getattr(lib, s) is not None


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", F.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+\w?)", out))
    assert archs == {"100a"}, archs


def test_keypoint_layout_matches_cv_keypoint():
    assert F.KP_DTYPE.itemsize == 28
    assert [F.KP_DTYPE.fields[n][1] for n in F.KP_DTYPE.names] == [0, 4, 8, 12, 16, 20, 24]


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.hyorb_version()
    assert isinstance(lib.hyorb_last_error(), bytes)


def test_bad_params_are_rejected_before_touching_cuda(lib):
    h = C.c_void_p()
    p = F.ExtractorParams(1000, 1.2, 99, 30, 20, 4, 0)      # nlevels out of range
    assert lib.hyorb_extractor_create(C.byref(p), 0, None, C.byref(h)) == F.EINVAL
    p = F.ExtractorParams(1000, 1.2, 8, 30, 20, 4, 7)       # reserved flags
    assert lib.hyorb_extractor_create(C.byref(p), 0, None, C.byref(h)) == F.EINVAL
    assert lib.hyorb_extractor_create(None, 0, None, C.byref(h)) == F.EINVAL


def test_no_cpu_fallback(lib):
    if lib.hyorb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    p = F.ExtractorParams(1000, 1.2, 8, 30, 20, 4, 0)
    assert lib.hyorb_extractor_create(C.byref(p), 0, None, C.byref(h)) == F.ECUDA
    assert b"no CPU fallback" in lib.hyorb_last_error()
    m = C.c_void_p()
    assert lib.hyorb_matcher_create(0, None, C.byref(m)) == F.ECUDA
    import hyslam_b200 as hb
    with pytest.raises(hb.HyorbError):
        hb.ORBExtractor()


def test_product_never_imports_the_oracle():
    """the product path may not import, include, link or dlopen anything under oracle/"""
    pkg = os.path.join(ROOT, "hyslam_b200")
    bad = re.compile(r"^\s*(from|import)\s+[\w\.]*oracle|#\s*include[^\n]*oracle|liborb_oracle|dlopen|orc_\w+\s*\(", re.M)
    for dp, _, fs in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), f


def test_header_is_plain_c99(tmp_path):
    """the ABI header is consumable from C (cgo / JNI / N-API style bindings bind exactly these declarations)"""
    src = tmp_path / "c.c"
    src.write_text('#include "hyorb.h"\nint main(void){ hyorb_projection p; hyorb_landmark l; hyorb_window_query q; (void)p; (void)l; (void)q; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
