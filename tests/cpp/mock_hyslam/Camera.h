// TEST DOUBLE: the members of hySLAM src/core/Camera.h:26-45 that the stereo matcher reads.
#pragma once
#include <opencv2/core/core.hpp>
namespace HYSLAM {
class Camera {
public:
    cv::Mat K;              // 3x3 CV_32F calibration matrix
    float mbf = 0;          // stereo baseline times fx
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    float fx() const { return K.at<float>(0, 0); }
    float mb() const { return mbf / K.at<float>(0, 0); }
};
}
