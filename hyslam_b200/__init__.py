"""hyslam_b200 -- B200-native ORB front end (extract + stereo + match) behind hySLAM's extractor / matcher interfaces.

The compute lives in libhyorb.so (hand-written sm_100a CUDA behind the C ABI of include/hyorb.h); this package is the
thin host-side mirror of the reference's interfaces used by the tests and the benchmark.
"""
import os as _os

# more hardware queues than the default 8, before anything creates the CUDA context (see api.cu, hyorb_default_connections)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._ffi import KP_DTYPE, WQ_DTYPE, HyorbError, build, device_count  # noqa: F401
from .settings import FeatureExtractorSettings, FeatureMatcherSettings, StereoCamera  # noqa: F401
from .extractor import ORBExtractor  # noqa: F401
from .matcher import FeatureMatcher, Stereomatcher  # noqa: F401
