"""The C++ drop-in (include/hyorb_hyslam.hpp: hySLAM's FeatureExtractor / ORBFactory / Stereomatcher surfaces over the C ABI).
With the reference tree present it must compile -- every `override` checked by the compiler -- against the REAL hySLAM headers
(/root/reference/src/{features,core}) and link with the reference's own FeatureDescriptor / FeatureViews translation units;
the test doubles used on a box without the reference tree must keep compiling too.  Running it needs a GPU:
tests/test_gpu_cpp_shim.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HYSLAM_REFERENCE", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "src", "features", "FeatureExtractor.h"))


def _make(*targets):
    from hyslam_b200 import _ffi
    _ffi.build()
    return subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), f"REF={REF}", *targets], capture_output=True, text=True)


def test_shim_compiles_and_links():
    r = _make()
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "_build", "shim_driver"))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree absent")
def test_shim_is_built_against_the_real_hyslam_headers():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "clean"], capture_output=True)
    r = _make()
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "real hySLAM headers" in open(os.path.join(ROOT, "tests", "cpp", "_build", "flavour.txt")).read()
    # the compiler, not a grep, is the judge of the signatures: every member the shim marks `override` exists as a virtual of the
    # reference's FeatureExtractor / ORBFactory with exactly that signature, or the line above would not have compiled.
    src = open(os.path.join(ROOT, "include", "hyorb_hyslam.hpp")).read()
    assert src.count(" override") >= 8 and "~CudaORBExtractor() override" not in src


def test_test_doubles_still_compile():
    r = _make("doubles")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(not HAVE_REF, reason="reference tree absent")
def test_doubles_declare_what_the_reference_declares():
    """the doubles are only trustworthy if they do not promise more than hySLAM's headers: no virtual destructor on
    FeatureExtractor (FeatureExtractor.h:25-37), extractor settings private to ORBFactory (ORBFactory.h:29-34)"""
    real = open(os.path.join(REF, "src", "features", "FeatureExtractor.h")).read()
    dbl = open(os.path.join(ROOT, "tests", "cpp", "mock_hyslam", "hyslam_test_doubles.hpp")).read()
    assert "~FeatureExtractor" not in real and "~FeatureExtractor" not in dbl
    for name in ("GetLevels", "GetScaleFactor", "GetScaleFactors", "GetInverseScaleFactors", "GetScaleSigmaSquares", "GetInverseScaleSigmaSquares"):
        assert f"{name}()" in real and f"{name}()" in dbl


@pytest.mark.skipif(not HAVE_REF, reason="reference tree absent")
def test_matcher_shim_links_against_the_reference_classes():
    """include/hyorb_hyslam_matcher.hpp (CudaFeatureMatcher: FeatureMatcher.h:105-176's public signatures) compiles against the real hySLAM
    headers and links (-z defs) against the reference's own Frame / KeyFrame / MapPoint objects in oracle/_ref"""
    from oracle import ref as R
    R.build()
    r = _make()
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "_build", "libmatcher_shim_test.so"))
    hdr = open(os.path.join(ROOT, "include", "hyorb_hyslam_matcher.hpp")).read()
    ref = open(os.path.join(REF, "src", "features", "FeatureMatcher.h")).read()
    for sig in ("int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th = 3)",
                "int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)",
                "int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize = 10)"):
        assert sig.replace(" ", "") in ref.replace(" ", "")          # the reference declares it ...
    for name in ("SearchByProjection", "SearchByBoW", "SearchForTriangulation", "SearchForInitialization", "Fuse", "SearchBySim3"):
        assert f"int {name}(" in hdr                                   # ... and the drop-in defines the same entry point
