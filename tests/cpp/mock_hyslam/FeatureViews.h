// TEST DOUBLE mirroring the parts of hySLAM src/core/FeatureViews.h:20-81 the stereo path uses.
#pragma once
#include <FeatureExtractorSettings.h>
#include <FeatureDescriptor.h>
#include <opencv2/opencv.hpp>
#include <vector>
namespace HYSLAM {
class FeatureViews {
public:
    FeatureViews() {}
    FeatureViews(std::vector<cv::KeyPoint> k, std::vector<cv::KeyPoint> kr, std::vector<FeatureDescriptor> d, std::vector<FeatureDescriptor> dr, FeatureExtractorSettings p)
        : is_stereo(true), is_empty(false), N((int)k.size()), mvKeys(k), mvKeysRight(kr), mDescriptors(d), mDescriptorsRight(dr), orb_params(p) {}
    bool empty() const { return is_empty; }
    bool isStereo() const { return is_stereo; }
    int numViews() const { return N; }
    FeatureExtractorSettings getOrbParams() const { return orb_params; }
    std::vector<cv::KeyPoint> getKeys() const { return mvKeys; }
    std::vector<cv::KeyPoint> getKeysR() const { return mvKeysRight; }
    std::vector<float> getuRs() const { return mvuRight; }
    std::vector<float> getDepths() const { return mvDepth; }
    std::vector<FeatureDescriptor> getDescriptors() const { return mDescriptors; }
    std::vector<FeatureDescriptor> getDescriptorsR() const { return mDescriptorsRight; }
    void setuRs(std::vector<float> uRs) { mvuRight = uRs; }
    void setDepths(std::vector<float> depths) { mvDepth = depths; }
private:
    bool is_stereo = false, is_empty = true; int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight; std::vector<float> mvuRight, mvDepth;
    std::vector<FeatureDescriptor> mDescriptors, mDescriptorsRight; FeatureExtractorSettings orb_params;
};
}
