// TEST DOUBLE mirroring hySLAM src/features/ORBFactory.h:19-37 with the defaults of ORBFactory.cpp:15-23 (no YAML here).
#pragma once
#include <FeatureFactory.h>
namespace HYSLAM {
class ORBFactory : public FeatureFactory {
public:
    ORBFactory() { extractor_settings.nFeatures = 1000; extractor_settings.fScaleFactor = 1.2f; extractor_settings.nLevels = 8;
                   extractor_settings.init_threshold = 20; extractor_settings.min_threshold = 7; extractor_settings.N_CELLS = 30; }
    ORBFactory(std::string) : ORBFactory() {}
    std::shared_ptr<FeatureExtractor> getExtractor(std::string) override { return nullptr; }              // the CPU extractor is not built here
    std::shared_ptr<FeatureExtractor> getExtractor(FeatureExtractorSettings) override { return nullptr; }
    FeatureVocabulary *getVocabulary(std::string) override { return nullptr; }
    std::shared_ptr<DescriptorDistance> getDistanceFunc() override { return std::make_shared<ORBDistance>(); }
    FeatureExtractorSettings getFeatureExtractorSettings() override { return extractor_settings; }
protected:
    FeatureExtractorSettings extractor_settings;
};
}
