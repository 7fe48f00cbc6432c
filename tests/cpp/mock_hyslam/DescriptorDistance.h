// TEST DOUBLE mirroring the interface of hySLAM src/features/low_level/DescriptorDistance.h:22-35 (declarations only).
#pragma once
#include <opencv2/core/core.hpp>
namespace HYSLAM {
class DescriptorDistance {
public:
    virtual ~DescriptorDistance() {}
    virtual float distance(const cv::Mat &D1, const cv::Mat &D2) = 0;
};
class ORBDistance : public DescriptorDistance {
public:
    float distance(const cv::Mat &D1, const cv::Mat &D2) override {      // 256-bit Hamming distance
        int d = 0;
        for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(D1.data[i] ^ D2.data[i]));
        return (float)d;
    }
};
}
