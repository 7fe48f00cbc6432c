// cvshim: stand-in for the g2o header of this name (hySLAM vendors g2o under Thirdparty/, but it needs Eigen).  Converter.h
// only names these types in declarations; nothing on the feature path uses them.
#pragma once
#include <Eigen/Core>
namespace g2o { class SE3Quat {}; }
