// forwards to the test doubles (tests/cpp/mock_hyslam/hyslam_test_doubles.hpp); not a hySLAM source file
#include "hyslam_test_doubles.hpp"
