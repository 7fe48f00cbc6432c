// Test scaffolding for the REAL-HEADER build of tests/cpp/shim_driver: member definitions for HYSLAM::ORBFactory as
// declared in the reference's src/features/ORBFactory.h.  The reference's ORBFactory.cpp cannot be compiled here (it pulls
// ORBVocabulary.h -> DBoW2's templated vocabulary, an un-vendored dependency), so this file supplies the same control flow
// with the YAML read replaced by a table holding the values of config/slam_feature_config.yaml:
//   ORBFactory()                    defaults of ORBFactory.cpp:13-25
//   getExtractor(std::string type)  LoadSettings(type) then the VIRTUAL getExtractor(settings)   (ORBFactory.cpp:32-35)
// The CPU extractor itself is not built in this driver; getExtractor(settings) of the base returns null.
#include <ORBFactory.h>

namespace HYSLAM {

ORBFactory::ORBFactory()
{
    extractor_settings.nFeatures = 1000;
    extractor_settings.fScaleFactor = 1.2;
    extractor_settings.nLevels = 8;
    extractor_settings.N_CELLS = 30;
    extractor_settings.init_threshold = 20;
    extractor_settings.min_threshold = 4;
    matcher_settings.TH_HIGH = 100.0;
    matcher_settings.TH_LOW = 50.0;
}

ORBFactory::ORBFactory(std::string settings_path_) : settings_path(settings_path_) { LoadSettings(settings_path, "SLAM"); }

std::shared_ptr<FeatureExtractor> ORBFactory::getExtractor(std::string type)
{
    LoadSettings(settings_path, type);
    return getExtractor(extractor_settings);
}

std::shared_ptr<FeatureExtractor> ORBFactory::getExtractor(FeatureExtractorSettings) { return nullptr; }
FeatureVocabulary *ORBFactory::getVocabulary(std::string) { return nullptr; }
std::shared_ptr<DescriptorDistance> ORBFactory::getDistanceFunc() { return std::make_shared<ORBDistance>(); }
FeatureExtractorSettings ORBFactory::getFeatureExtractorSettings() { return extractor_settings; }

void ORBFactory::LoadSettings(std::string, std::string type)
{
    // config/slam_feature_config.yaml: ORB.SLAM.Extractor and ORB.Imaging.Extractor
    const bool imaging = type == "Imaging";
    extractor_settings.nFeatures = imaging ? 3000 : 1000;
    extractor_settings.fScaleFactor = imaging ? 1.4f : 1.2f;
    extractor_settings.nLevels = 8;
    extractor_settings.N_CELLS = 30;
    extractor_settings.init_threshold = 20;
    extractor_settings.min_threshold = 4;
    matcher_settings.TH_HIGH = 100.0;
    matcher_settings.TH_LOW = 50.0;
}

}  // namespace HYSLAM
