// TEST DOUBLE mirroring hySLAM src/core/FeatureExtractorSettings.h:19-33.
#pragma once
namespace HYSLAM {
class FeatureExtractorSettings {
public:
    int nFeatures; float fScaleFactor; int nLevels; int init_threshold; int min_threshold;
    float size_ref = 31; float sigma_ref = 1.0; int N_CELLS;
};
}
