// TEST DOUBLE: opaque stand-in for hySLAM's FeatureVocabulary (DBoW2-backed, out of scope).
#pragma once
namespace HYSLAM { class FeatureVocabulary {}; }
