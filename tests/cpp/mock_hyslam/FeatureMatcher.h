// TEST DOUBLE: only the settings struct of hySLAM src/features/FeatureMatcher.h:98-103 (the matcher class itself needs
// MapPoint / KeyFrame / Frame, which are outside the hot path).
#pragma once
namespace HYSLAM {
struct FeatureMatcherSettings { float nnratio = 0.6; float TH_HIGH = 100.0; float TH_LOW = 50.0; bool checkOri = true; };
}
