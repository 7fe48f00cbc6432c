"""The C++ drop-in (include/hyorb_hyslam.hpp: hySLAM's FeatureExtractor / ORBFactory / Stereomatcher surfaces over the C
ABI) compiles against test doubles of the hySLAM and OpenCV headers.  Running it needs a GPU: tests/test_gpu_cpp_shim.py."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_and_links():
    from hyslam_b200 import _ffi
    _ffi.build()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "_build", "shim_driver"))


def test_shim_mirrors_the_reference_signatures():
    """the names the reference's call sites use (ImageProcessing.cpp:82-103, System.cc:77-85) exist in the shim"""
    src = open(os.path.join(ROOT, "include", "hyorb_hyslam.hpp")).read()
    for needle in ["class CudaORBExtractor : public FeatureExtractor", "class CudaORBFactory : public ORBFactory",
                   "void operator()(cv::InputArray image, cv::InputArray", "std::vector<cv::KeyPoint> &keypoints",
                   "std::vector<FeatureDescriptor> &descriptors", "CudaStereomatcher(FeatureViews views, Camera cam_data, FeatureMatcherSettings settings",
                   "void computeStereoMatches()", "void getData(std::vector<float> &mvuRight_, std::vector<float> &mvDepth_)",
                   "void getData(FeatureViews &views)", "std::shared_ptr<FeatureExtractor> getExtractor(FeatureExtractorSettings s) override"]:
        assert needle in src, needle
