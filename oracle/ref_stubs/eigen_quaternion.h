// ORACLE - TEST INFRASTRUCTURE ONLY.  Force-included (-include) when app/player/ply_player.cpp is compiled: its trajectory writer
// (pose_to_tum_string, ply_player.cpp:676-694) names Eigen::Quaternionf, which oracle/eigen_compat does not carry because nothing on the
// hot path uses it.  Standard rotation-matrix -> quaternion conversion; not called by the tests (they only use the PLY reader).
#pragma once
#include <cmath>
#include <Eigen/Dense>
namespace Eigen {
class Quaternionf {
 public:
  explicit Quaternionf(const Matrix3f& R) {
    const float t = R(0, 0) + R(1, 1) + R(2, 2);
    if (t > 0.0f) { float s = std::sqrt(t + 1.0f) * 2.0f; w_ = 0.25f * s; x_ = (R(2, 1) - R(1, 2)) / s; y_ = (R(0, 2) - R(2, 0)) / s; z_ = (R(1, 0) - R(0, 1)) / s; }
    else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) { float s = std::sqrt(1.0f + R(0, 0) - R(1, 1) - R(2, 2)) * 2.0f; w_ = (R(2, 1) - R(1, 2)) / s; x_ = 0.25f * s; y_ = (R(0, 1) + R(1, 0)) / s; z_ = (R(0, 2) + R(2, 0)) / s; }
    else if (R(1, 1) > R(2, 2)) { float s = std::sqrt(1.0f + R(1, 1) - R(0, 0) - R(2, 2)) * 2.0f; w_ = (R(0, 2) - R(2, 0)) / s; x_ = (R(0, 1) + R(1, 0)) / s; y_ = 0.25f * s; z_ = (R(1, 2) + R(2, 1)) / s; }
    else { float s = std::sqrt(1.0f + R(2, 2) - R(0, 0) - R(1, 1)) * 2.0f; w_ = (R(1, 0) - R(0, 1)) / s; x_ = (R(0, 2) + R(2, 0)) / s; y_ = (R(1, 2) + R(2, 1)) / s; z_ = 0.25f * s; }
  }
  float x() const { return x_; } float y() const { return y_; } float z() const { return z_; } float w() const { return w_; }
 private:
  float x_ = 0, y_ = 0, z_ = 0, w_ = 1;
};
}  // namespace Eigen
