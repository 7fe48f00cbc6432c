// matcher_shim_test.cpp -- the C++ drop-in matcher (include/hyorb_hyslam_matcher.hpp: HYSLAM::CudaFeatureMatcher, the public surface of
// FeatureMatcher.h:105-176 over the C ABI) driven on the reference's REAL Frame / KeyFrame / MapPoint objects.
//
// This file is oracle/ref_glue_match.cpp's matcher entry points compiled a second time with the matcher class swapped: the same scene
// handle (built by oracle/_ref's refm_scene_* / refm_add_* functions) is handed to shimm_search_by_projection(...) etc., which call
// CudaFeatureMatcher where the refm_* twins call HYSLAM::FeatureMatcher.  tests/test_gpu_cpp_matcher_shim.py runs both on twin scenes and
// compares what a hySLAM caller would observe.  Needs the reference headers: built only where /root/reference exists (tests/cpp/Makefile).
#include <hyorb_hyslam_matcher.hpp>

#define REFM_ENTRY_POINTS_ONLY
#define REFM_MATCHER CudaFeatureMatcher
#define REFM_NAME(x) shimm_##x
#include "../../oracle/ref_glue_match.cpp"
