// hyslam_test_doubles.hpp -- TEST DOUBLES of the hySLAM declarations that include/hyorb_hyslam.hpp builds on.
// This image has neither OpenCV's C++ headers nor a buildable hySLAM, so the C++ drop-in is compiled and run here against
// these stand-ins: each block states which hySLAM header it stands in for (file:line of the reference) and carries only
// the members the drop-in and its test driver touch.  The files named after hySLAM's headers next to this one are
// one-line forwards so that `#include <FeatureExtractor.h>` etc. resolve exactly as they do inside a hySLAM build.
#pragma once
#include <opencv2/core/core.hpp>
#include <memory>
#include <string>
#include <vector>

// ---- stands in for DescriptorDistance.h
// TEST DOUBLE mirroring the interface of hySLAM src/features/low_level/DescriptorDistance.h:22-35 (declarations only).
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

// ---- stands in for FeatureDescriptor.h
// TEST DOUBLE mirroring hySLAM src/features/low_level/FeatureDescriptor.h:26-38.
namespace HYSLAM {
class FeatureDescriptor {
public:
    FeatureDescriptor() {}
    FeatureDescriptor(cv::Mat desc, std::shared_ptr<DescriptorDistance> distfunc_) : descriptor(desc.clone()), distfunc(distfunc_), is_empty(false) {}
    float distance(const FeatureDescriptor &d2) const { return distfunc->distance(descriptor, d2.rawDescriptor()); }
    cv::Mat rawDescriptor() const { return descriptor.clone(); }
    bool isEmpty() { return is_empty; }
private:
    cv::Mat descriptor; std::shared_ptr<DescriptorDistance> distfunc; bool is_empty = true;
};
}

// ---- stands in for FeatureExtractorSettings.h
// TEST DOUBLE mirroring hySLAM src/core/FeatureExtractorSettings.h:19-33.
namespace HYSLAM {
class FeatureExtractorSettings {
public:
    int nFeatures; float fScaleFactor; int nLevels; int init_threshold; int min_threshold;
    float size_ref = 31; float sigma_ref = 1.0; int N_CELLS;
};
}

// ---- stands in for FeatureExtractor.h
// TEST DOUBLE mirroring hySLAM src/features/FeatureExtractor.h:25-37.
namespace HYSLAM {
class FeatureExtractor {           // like the reference: NO virtual destructor
public:
    virtual void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints, std::vector<FeatureDescriptor> &descriptors) = 0;
    virtual int GetLevels() = 0;
    virtual float GetScaleFactor() = 0;
    virtual std::vector<float> GetScaleFactors() = 0;
    virtual std::vector<float> GetInverseScaleFactors() = 0;
    virtual std::vector<float> GetScaleSigmaSquares() = 0;
    virtual std::vector<float> GetInverseScaleSigmaSquares() = 0;
};
}

// ---- stands in for FeatureMatcher.h
// TEST DOUBLE: only the settings struct of hySLAM src/features/FeatureMatcher.h:98-103 (the matcher class itself needs
// MapPoint / KeyFrame / Frame, which are outside the hot path).
namespace HYSLAM {
struct FeatureMatcherSettings { float nnratio = 0.6; float TH_HIGH = 100.0; float TH_LOW = 50.0; bool checkOri = true; };
}

// ---- stands in for FeatureVocabulary.h
// TEST DOUBLE: opaque stand-in for hySLAM's FeatureVocabulary (DBoW2-backed, out of scope).
namespace HYSLAM { class FeatureVocabulary {}; }

// ---- stands in for FeatureFactory.h
// TEST DOUBLE mirroring hySLAM src/features/FeatureFactory.h:21-33.
namespace HYSLAM {
class FeatureFactory {
public:
    virtual ~FeatureFactory() {}
    virtual std::shared_ptr<FeatureExtractor> getExtractor(std::string type) = 0;
    virtual std::shared_ptr<FeatureExtractor> getExtractor(FeatureExtractorSettings settings) = 0;
    virtual FeatureVocabulary *getVocabulary(std::string type) = 0;
    virtual std::shared_ptr<DescriptorDistance> getDistanceFunc() = 0;
    virtual FeatureExtractorSettings getFeatureExtractorSettings() = 0;
    FeatureMatcherSettings getFeatureMatcherSettings() const { return matcher_settings; }
    void setFeatureMatcherSettings(FeatureMatcherSettings fm_settings) { matcher_settings = fm_settings; }
protected:
    FeatureMatcherSettings matcher_settings;
};
}

// ---- stands in for ORBFactory.h
// TEST DOUBLE mirroring hySLAM src/features/ORBFactory.h:19-37 with the defaults of ORBFactory.cpp:15-23 (no YAML here).
namespace HYSLAM {
class ORBFactory : public FeatureFactory {
public:
    ORBFactory() { load("SLAM"); extractor_settings.min_threshold = 4; }
    ORBFactory(std::string settings_path_) : settings_path(settings_path_) { load("SLAM"); }
    // ORBFactory.cpp:32-35: re-read the block of this camera type, then dispatch VIRTUALLY to the settings overload
    std::shared_ptr<FeatureExtractor> getExtractor(std::string type) override { load(type); return getExtractor(extractor_settings); }
    std::shared_ptr<FeatureExtractor> getExtractor(FeatureExtractorSettings) override { return nullptr; }      // the CPU extractor is not built here
    FeatureVocabulary *getVocabulary(std::string) override { return nullptr; }
    std::shared_ptr<DescriptorDistance> getDistanceFunc() override { return std::make_shared<ORBDistance>(); }
    FeatureExtractorSettings getFeatureExtractorSettings() override { return extractor_settings; }
private:                                           // private in the reference too (ORBFactory.h:29-34)
    FeatureExtractorSettings extractor_settings;
    std::string vocab_path, settings_path;
    void load(const std::string &type)             // stands in for LoadSettings: the values of config/slam_feature_config.yaml
    {
        const bool imaging = type == "Imaging";
        extractor_settings.nFeatures = imaging ? 3000 : 1000; extractor_settings.fScaleFactor = imaging ? 1.4f : 1.2f; extractor_settings.nLevels = 8;
        extractor_settings.init_threshold = 20; extractor_settings.min_threshold = 4; extractor_settings.N_CELLS = 30;
        matcher_settings.TH_HIGH = 100.0; matcher_settings.TH_LOW = 50.0;
    }
};
}

// ---- stands in for Camera.h
// TEST DOUBLE: the members of hySLAM src/core/Camera.h:26-45 that the stereo matcher reads.
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

// ---- stands in for FeatureViews.h
// TEST DOUBLE mirroring the parts of hySLAM src/core/FeatureViews.h:20-81 the stereo path uses.
namespace HYSLAM {
class FeatureViews {
public:
    FeatureViews() {}
    FeatureViews(std::vector<cv::KeyPoint> k, std::vector<FeatureDescriptor> d, FeatureExtractorSettings p)      // mono constructor (FeatureViews.h:24)
        : is_stereo(false), is_empty(false), N((int)k.size()), mvKeys(k), mDescriptors(d), orb_params(p) {}
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
