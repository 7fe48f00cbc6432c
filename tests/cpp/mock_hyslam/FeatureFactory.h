// TEST DOUBLE mirroring hySLAM src/features/FeatureFactory.h:21-33.
#pragma once
#include <FeatureExtractor.h>
#include <FeatureVocabulary.h>
#include <FeatureMatcher.h>
#include <memory>
#include <string>
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
