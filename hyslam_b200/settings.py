"""Plain-data mirrors of the reference's settings structs."""
from dataclasses import dataclass


@dataclass
class FeatureExtractorSettings:
    """HYSLAM::FeatureExtractorSettings (src/core/FeatureExtractorSettings.h:19-33); defaults of ORBFactory.cpp:15-23 /
    config/slam_feature_config.yaml:8-14."""
    nFeatures: int = 1000
    fScaleFactor: float = 1.2
    nLevels: int = 8
    N_CELLS: int = 30          # really the FAST cell edge in px (ORBExtractor.cpp:409)
    init_threshold: int = 20   # dead in the reference (ORBFinder.cpp:58-60)
    min_threshold: int = 4     # dead in the reference
    size_ref: float = 31.0


@dataclass
class FeatureMatcherSettings:
    """HYSLAM::FeatureMatcherSettings (src/features/FeatureMatcher.h:98-103)."""
    nnratio: float = 0.6
    TH_HIGH: float = 100.0
    TH_LOW: float = 50.0
    checkOri: bool = True


@dataclass
class StereoCamera:
    """The Camera fields Stereomatcher reads (src/features/Stereomatcher.cpp:14-34): mbf, fx, mnMaxY."""
    mbf: float
    fx: float
    mnMaxY: float
