// TEST DOUBLE mirroring hySLAM src/features/FeatureExtractor.h:25-37.
#pragma once
#include <FeatureExtractorSettings.h>
#include <FeatureDescriptor.h>
#include <DescriptorDistance.h>
#include <opencv/cv.h>
#include <memory>
#include <vector>
namespace HYSLAM {
class FeatureExtractor {
public:
    virtual ~FeatureExtractor() {}
    virtual void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints, std::vector<FeatureDescriptor> &descriptors) = 0;
    virtual int GetLevels() = 0;
    virtual float GetScaleFactor() = 0;
    virtual std::vector<float> GetScaleFactors() = 0;
    virtual std::vector<float> GetInverseScaleFactors() = 0;
    virtual std::vector<float> GetScaleSigmaSquares() = 0;
    virtual std::vector<float> GetInverseScaleSigmaSquares() = 0;
};
}
