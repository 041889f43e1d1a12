// ref_stubs.h — the minimum of the reference's util/ + database/ types that b2lo_dropin.h touches, for the
// stand-alone compile / run check of this repository (Eigen and the reference tree are absent on the GPU box).
// Same names, members and layouts as /root/reference/src/util/PointCloudUtils.h:34-297, MathUtils.h:101-168,
// database/LidarFrame.h:102-213, optimization/AdaptiveMEstimator.h:28-85.  NOT used when building inside the reference.
#pragma once
#include <cmath>
#include <memory>
#include <string>
#include <vector>

namespace Eigen {
struct Vector3f { float v[3]; Vector3f() : v{0, 0, 0} {} Vector3f(float a, float b, float c) : v{a, b, c} {} float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; } };
struct Vector3d { double v[3]; Vector3d() : v{0, 0, 0} {} Vector3d(double a, double b, double c) : v{a, b, c} {} double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; } };
struct Matrix4f {
  float m[16];
  Matrix4f() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }
  float& operator()(int r, int c) { return m[c * 4 + r]; }          // column-major like Eigen
  float operator()(int r, int c) const { return m[c * 4 + r]; }
};
}  // namespace Eigen

namespace lidar_slam {
namespace util {
struct Point3D { float x, y, z; Point3D() : x(0), y(0), z(0) {} Point3D(float a, float b, float c) : x(a), y(b), z(c) {} };
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud>;
  using ConstPtr = std::shared_ptr<const PointCloud>;
  void push_back(const Point3D& p) { points.push_back(p); }
  void push_back(float x, float y, float z) { points.emplace_back(x, y, z); }
  const Point3D& operator[](size_t i) const { return points[i]; }
  Point3D& operator[](size_t i) { return points[i]; }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void reserve(size_t n) { points.reserve(n); }
  std::vector<Point3D> points;
};
using PointCloudPtr = PointCloud::Ptr;
using PointCloudConstPtr = PointCloud::ConstPtr;
class KdTree {};
}  // namespace util

class SE3f {
 public:
  SE3f() {}
  explicit SE3f(const Eigen::Matrix4f& M) : m_(M) {}
  Eigen::Matrix4f Matrix() const { return m_; }
 private:
  Eigen::Matrix4f m_;
};

namespace database {
class LidarFrame {
 public:
  explicit LidarFrame(util::PointCloudPtr c) : m_feature_cloud(std::move(c)) {}
  util::PointCloudConstPtr get_feature_cloud() const { return m_feature_cloud; }
  util::PointCloudConstPtr get_processed_cloud() const { return m_feature_cloud; }
  void set_pose(const SE3f& p) { m_pose = p; }
  SE3f get_pose() const { return m_pose; }
 private:
  util::PointCloudPtr m_feature_cloud;
  SE3f m_pose;
};
}  // namespace database

namespace optimization {
struct AdaptiveMEstimatorConfig {
  bool use_adaptive_m_estimator = true;
  std::string loss_type = "huber";
  double min_scale_factor = 0.1, max_scale_factor = 10.0;
  int num_alpha_segments = 100;
  double truncated_threshold = 10.0;
  int gmm_components = 3, gmm_sample_size = 100;
  std::string pko_kernel_type = "huber";
};
class AdaptiveMEstimator {
 public:
  const AdaptiveMEstimatorConfig& get_config() const { return m_config; }
  void reset() {}
 private:
  AdaptiveMEstimatorConfig m_config;
};
}  // namespace optimization
}  // namespace lidar_slam
