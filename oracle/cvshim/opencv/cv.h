// cvshim forward: see opencv2/core/core.hpp (test infrastructure; stands in for the OpenCV 3.x header of this name)
#pragma once
#include <opencv2/core/core.hpp>
