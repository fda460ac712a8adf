#pragma once
#include <Eigen/Dense>
#include <tf/tf.h>
namespace tf {
inline void transformTFToEigen(const Transform &, Eigen::Affine3d &) {}
inline void transformEigenToTF(const Eigen::Affine3d &, Transform &) {}
}
