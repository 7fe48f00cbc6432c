// cvshim: see g2o/types/sba/types_six_dof_expmap.h
#pragma once
#include <Eigen/Core>
namespace g2o { class Sim3 {}; }
