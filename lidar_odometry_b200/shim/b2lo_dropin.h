// b2lo_dropin.h — C++ drop-in for the reference's hot-path classes over the C ABI of include/b2lo.h.
//
// Same namespaces, class names, member signatures and error behaviour as
//   lidar_slam::map::FastVoxelFilter / FastVoxelGrid          /root/reference/src/database/VoxelMap.h:53-143
//   lidar_slam::map::VoxelMap                                 src/database/VoxelMap.h:188-332, VoxelMap.cpp
//   lidar_slam::optimization::IterativeClosestPointOptimizer  src/optimization/IterativeClosestPointOptimizer.h:158-225
// so that src/processing/Estimator.cpp (and PangolinViewer.cpp:966-973) compile against it unchanged.  Everything
// the reference computes on the CPU inside these classes runs on the GPU behind libb2lo.so; the classes themselves
// only marshal host buffers.  optimize_loop (loop-closure ICP) runs on the GPU as well (b2lo_icp_optimize_loop, on its own context
// because the reference calls it from the loop/PGO thread next to optimize()); define B2LO_KEEP_REFERENCE_LOOP_ICP to keep the
// reference's CPU body instead.  util::VoxelGrid (final-map export) and the scan-ingest helpers are at the end.
//
// Build inside the reference tree: replace the two headers' class bodies with `#include "b2lo_dropin.h"` as
// INTEGRATION.md shows; the reference's own util/ headers (PointCloud, SE3f, KdTree, LidarFrame, ICPConfig,
// AdaptiveMEstimator) are used as they are.  For the stand-alone compile check of this repository the same
// names come from shim/test/ref_stubs.h (define B2LO_SHIM_STUBS).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "b2lo.h"

#ifdef B2LO_SHIM_STUBS
#include "test/ref_stubs.h"
#else
#include "database/LidarFrame.h"
#include "optimization/AdaptiveMEstimator.h"
#include "util/MathUtils.h"
#include "util/PointCloudUtils.h"
#endif

namespace lidar_slam {

namespace b2lo_detail {
// one context (device 0 unless B2LO_DEVICE says otherwise) shared by all objects of the process
inline b2lo_ctx* context() {
  static b2lo_ctx* ctx = [] {
    b2lo_ctx* c = nullptr;
    int dev = 0;
    if (const char* e = std::getenv("B2LO_DEVICE")) dev = std::atoi(e);
    if (b2lo_ctx_create(dev, &c) != B2LO_OK) throw std::runtime_error(std::string("b2lo: ") + b2lo_last_error());
    return c;
  }();
  return ctx;
}
inline size_t stride_floats() { return sizeof(util::Point3D) / sizeof(float); }
inline const float* data(const util::PointCloud& c) { return c.empty() ? nullptr : &c[0].x; }
inline void to_row_major(const Eigen::Matrix4f& M, float T[16]) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[r * 4 + c] = M(r, c); }
inline Eigen::Matrix4f from_row_major(const float T[16]) { Eigen::Matrix4f M; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M(r, c) = T[r * 4 + c]; return M; }
}  // namespace b2lo_detail

namespace map {

class FastVoxelFilter {  // VoxelMap.h:53-143
 public:
  explicit FastVoxelFilter(float voxel_size = 0.5f) : m_voxel_size(voxel_size) {}
  void setVoxelSize(float voxel_size) { m_voxel_size = voxel_size; }
  float getVoxelSize() const { return m_voxel_size; }
  void filter(const util::PointCloud& input, util::PointCloud& output, int stride = 1) {
    output.clear();
    m_count = 0;
    if (input.empty() || stride < 1) return;  // VoxelMap.h:76
    const size_t ns = (input.size() + (size_t)stride - 1) / (size_t)stride;
    m_buf.resize(ns * 3);
    size_t m = 0;
    int rc = b2lo_filter(b2lo_detail::context(), b2lo_detail::data(input), input.size(), b2lo_detail::stride_floats(), stride, m_voxel_size,
                         m_buf.data(), nullptr, &m);
    if (rc < 0) throw std::runtime_error(std::string("b2lo_filter: ") + b2lo_last_error());
    output.reserve(m);
    for (size_t i = 0; i < m; ++i) output.push_back(util::Point3D(m_buf[i * 3], m_buf[i * 3 + 1], m_buf[i * 3 + 2]));
    m_count = m;
  }
  size_t getVoxelCount() const { return m_count; }

 private:
  float m_voxel_size;
  size_t m_count = 0;
  std::vector<float> m_buf;
};
using FastVoxelGrid = FastVoxelFilter;  // VoxelMap.h:143

class VoxelMap {  // VoxelMap.h:188-332
 public:
  using PointCloud = lidar_slam::util::PointCloud;
  using PointCloudPtr = lidar_slam::util::PointCloudPtr;
  using PointCloudConstPtr = lidar_slam::util::PointCloudConstPtr;
  using Point3D = lidar_slam::util::Point3D;

  explicit VoxelMap(float voxel_size = 0.5f) : m_voxel_size(voxel_size) { create(); }
  ~VoxelMap() { if (m_map) b2lo_map_destroy(m_map); }
  VoxelMap(const VoxelMap&) = delete;
  VoxelMap& operator=(const VoxelMap&) = delete;

  void SetVoxelSize(float size) {  // VoxelMap.cpp:27-36
    if (size <= 0) throw std::invalid_argument("Voxel size must be positive");
    if (std::abs(m_voxel_size - size) > 1e-6f) { m_voxel_size = size; create(); }
  }
  void SetMaxHitCount(int max_count) { m_max_hit_count = max_count; }   // dead fields of the reference, kept for the setters
  void SetInitHitCount(int count) { m_init_hit_count = count; }
  void SetHierarchyFactor(int factor) {  // VoxelMap.cpp:38-48: rejects even / non-positive factors
    if (factor <= 0 || factor % 2 == 0) return;
    if (factor != 3) throw std::invalid_argument("b2lo: this build specialises the 3x3x3 hierarchy the reference uses (Estimator.cpp:79)");
  }
  void SetPlanarityThreshold(float threshold) { m_planarity_threshold = threshold; b2lo_map_set_planarity_threshold(m_map, threshold); }
  void SetComputeSurfels(bool compute) { m_compute_surfels = compute; b2lo_map_set_compute_surfels(m_map, compute ? 1 : 0); }

  float GetVoxelSize() const { return m_voxel_size; }
  int GetHierarchyFactor() const { return 3; }
  size_t GetVoxelCount() const { size_t a = 0; b2lo_map_counts(m_map, &a, nullptr, nullptr); return a; }
  size_t GetL1VoxelCount() const { size_t b = 0; b2lo_map_counts(m_map, nullptr, &b, nullptr); return b; }
  bool empty() const { return GetVoxelCount() == 0; }
  bool GetComputeSurfels() const { return m_compute_surfels; }
  size_t GetSurfelCount() const { size_t c = 0; b2lo_map_counts(m_map, nullptr, nullptr, &c); return c; }
  void Clear() { b2lo_map_clear(m_map); m_kdtree_ready = false; }

  void UpdateVoxelMap(const PointCloudConstPtr& new_cloud, const Eigen::Vector3d& sensor_position, double max_distance, bool is_keyframe) {
    if (!new_cloud || new_cloud->empty()) return;  // VoxelMap.cpp:134-136
    if (!is_keyframe) return;                      // :138-140
    const double s[3] = {sensor_position.x(), sensor_position.y(), sensor_position.z()};
    int rc = b2lo_map_update(m_map, b2lo_detail::data(*new_cloud), new_cloud->size(), b2lo_detail::stride_floats(), s, max_distance);
    if (rc < 0 && rc != B2LO_E_RANGE) throw std::runtime_error(std::string("b2lo_map_update: ") + b2lo_last_error());
  }
  void ApplyTransformAndRehash(const Eigen::Matrix4f& T_correction) {
    float T[16];
    b2lo_detail::to_row_major(T_correction, T);
    if (b2lo_map_transform_rehash(m_map, T) < 0) throw std::runtime_error(std::string("b2lo_map_transform_rehash: ") + b2lo_last_error());
  }
  bool GetSurfelAtPoint(const Eigen::Vector3f& point, Eigen::Vector3f& normal, Eigen::Vector3f& centroid) const {
    const float p[3] = {point.x(), point.y(), point.z()};
    float n[3], c[3];
    if (b2lo_map_lookup(m_map, p, n, c) != 1) return false;
    normal = Eigen::Vector3f(n[0], n[1], n[2]);
    centroid = Eigen::Vector3f(c[0], c[1], c[2]);
    return true;
  }
  PointCloudPtr GetPointCloud() const {  // VoxelMap.cpp:388-403: L0 centroids in the reference's dense order
    auto cloud = std::make_shared<PointCloud>();
    size_t n = GetVoxelCount();
    std::vector<float> xyz(n * 3 + 3);
    if (n && b2lo_map_export_l0(m_map, xyz.data(), nullptr, nullptr, n, &n) < 0) throw std::runtime_error(std::string("b2lo_map_export_l0: ") + b2lo_last_error());
    cloud->reserve(n);
    for (size_t i = 0; i < n; ++i) cloud->push_back(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]);
    return cloud;
  }
  // KDTree mode: the engine searches its own L0 hash (no host kd-tree is ever built); the accessor keeps the
  // reference's "non-null once built" contract for Estimator.cpp:460-462 by handing out an empty tree object.
  std::shared_ptr<lidar_slam::util::KdTree> GetKdTree() const { return m_kdtree_ready ? m_kdtree_token : nullptr; }
  void RebuildKdTree() {
    b2lo_map_rebuild_knn(m_map);
    m_kdtree_ready = b2lo_map_has_knn(m_map) != 0;
    if (m_kdtree_ready && !m_kdtree_token) m_kdtree_token = std::make_shared<lidar_slam::util::KdTree>();
  }
  bool HasKdTree() const { return m_kdtree_ready; }
  std::vector<std::tuple<Eigen::Vector3f, Eigen::Vector3f, float>> GetL1Surfels() const {  // VoxelMap.cpp:405-418 (viewer thread)
    size_t n1 = GetL1VoxelCount(), n = 0;
    std::vector<float> c(n1 * 3 + 3), nr(n1 * 3 + 3), pl(n1 + 1);
    std::vector<std::tuple<Eigen::Vector3f, Eigen::Vector3f, float>> out;
    if (n1 == 0) return out;
    b2lo_map_export_surfels(m_map, c.data(), nr.data(), pl.data(), nullptr, n1, &n);
    if (n > n1) n = n1;
    out.reserve(n);
    for (size_t i = 0; i < n; ++i)
      out.emplace_back(Eigen::Vector3f(c[i * 3], c[i * 3 + 1], c[i * 3 + 2]), Eigen::Vector3f(nr[i * 3], nr[i * 3 + 1], nr[i * 3 + 2]), pl[i]);
    return out;
  }

  b2lo_map* handle() const { return m_map; }  // used by IterativeClosestPointOptimizer below

 private:
  void create() {
    if (m_map) { b2lo_map_destroy(m_map); m_map = nullptr; }
    if (b2lo_map_create(b2lo_detail::context(), m_voxel_size, 3, m_planarity_threshold, m_compute_surfels ? 1 : 0, 0, &m_map) != B2LO_OK)
      throw std::runtime_error(std::string("b2lo_map_create: ") + b2lo_last_error());
    m_kdtree_ready = false;
  }
  b2lo_map* m_map = nullptr;
  float m_voxel_size;
  int m_max_hit_count = 10, m_init_hit_count = 1;
  float m_planarity_threshold = 0.1f;
  bool m_compute_surfels = true;
  bool m_kdtree_ready = false;
  std::shared_ptr<lidar_slam::util::KdTree> m_kdtree_token;
};

}  // namespace map

namespace optimization {

#ifndef B2LO_SHIM_HAVE_ICPCONFIG
struct ICPConfig {  // IterativeClosestPointOptimizer.h:55-76
  int max_iterations = 50;
  double translation_tolerance = 1e-6;
  double rotation_tolerance = 1e-6;
  double max_correspondence_distance = 1.0;
  int min_correspondence_points = 10;
  double outlier_rejection_ratio = 0.9;
  bool use_robust_loss = true;
  double robust_loss_delta = 0.1;
  bool use_kdtree = true;
  int max_kdtree_neighbors = 1;
  bool use_surfel_correspondence = true;
};
#endif

class IterativeClosestPointOptimizer {  // IterativeClosestPointOptimizer.h:158-225
 public:
  struct OptimizationStats {  // :203-210
    size_t num_correspondences = 0;
    size_t num_iterations = 0;
    double initial_cost = 0.0;
    double final_cost = 0.0;
    double optimization_time_ms = 0.0;
    bool converged = false;
  };
  explicit IterativeClosestPointOptimizer(const ICPConfig& config = ICPConfig()) : m_config(config) {}
  IterativeClosestPointOptimizer(const ICPConfig& config, std::shared_ptr<optimization::AdaptiveMEstimator> adaptive_estimator)
      : m_config(config), m_adaptive_estimator(std::move(adaptive_estimator)) {}

  // ICP.cpp:255-463.  initial_transform is a WORLD pose guess (Estimator.cpp:154); on failure (< min correspondences at
  // some iteration) returns false and leaves optimized_transform = initial_transform.
  bool optimize(map::VoxelMap* voxel_map, std::shared_ptr<database::LidarFrame> curr_frame, const SE3f& initial_transform, SE3f& optimized_transform) {
    m_last_stats = OptimizationStats();
    optimized_transform = initial_transform;
    if (!voxel_map || !curr_frame) return false;
    util::PointCloudConstPtr cloud = curr_frame->get_feature_cloud();   // get_frame_cloud (ICP.cpp:769-783)
    if (!cloud || cloud->empty()) cloud = curr_frame->get_processed_cloud();
    if (!cloud || cloud->empty()) return false;
    b2lo_icp_cfg cfg;
    fill_cfg(cfg);
    float T0[16], T1[16];
    b2lo_detail::to_row_major(initial_transform.Matrix(), T0);
    b2lo_icp_stats st;
    int rc = b2lo_icp_optimize(voxel_map->handle(), b2lo_detail::data(*cloud), cloud->size(), b2lo_detail::stride_floats(), T0, &cfg, T1, &st);
    if (rc < 0) throw std::runtime_error(std::string("b2lo_icp_optimize: ") + b2lo_last_error());
    m_last_stats.num_correspondences = (size_t)st.num_correspondences;
    m_last_stats.num_iterations = (size_t)st.num_iterations;
    m_last_stats.initial_cost = st.initial_cost;
    m_last_stats.final_cost = st.final_cost;
    m_last_stats.optimization_time_ms = st.device_ms;
    if (rc != B2LO_OK) return false;
    m_last_stats.converged = true;  // the reference reports true whenever it returns true (ICP.cpp:452-462)
    const Eigen::Matrix4f M = b2lo_detail::from_row_major(T1);
    optimized_transform = SE3f(M);
    curr_frame->set_pose(optimized_transform);  // the reference leaves the frame at the last iterate (ICP.cpp:284)
    return true;
  }

  // Loop-closure ICP (ICP.cpp:40-251): defined below on the GPU; INTEGRATION.md shows how to keep the reference's CPU body instead.
  bool optimize_loop(std::shared_ptr<database::LidarFrame> curr_keyframe, std::shared_ptr<database::LidarFrame> matched_keyframe,
                     SE3f& optimized_relative_transform, float& inlier_ratio);

  const OptimizationStats& get_last_stats() const { return m_last_stats; }
  void update_config(const ICPConfig& config) { m_config = config; }
  const ICPConfig& get_config() const { return m_config; }

 private:
  void fill_cfg(b2lo_icp_cfg& c) const {
    b2lo_default_icp_cfg(&c);
    c.max_iterations = m_config.max_iterations;   // any count >= 1 (ICPConfig's default is 50, ICP.h:57); the engine stops issuing work once converged
    c.translation_tolerance = m_config.translation_tolerance;
    c.rotation_tolerance = m_config.rotation_tolerance;
    c.max_correspondence_distance = m_config.max_correspondence_distance;
    c.min_correspondence_points = m_config.min_correspondence_points;
    c.use_robust_loss = m_config.use_robust_loss ? 1 : 0;
    c.robust_loss_delta = m_config.robust_loss_delta;
    c.use_surfel_correspondence = m_config.use_surfel_correspondence ? 1 : 0;
    if (!m_adaptive_estimator) { c.use_adaptive_m_estimator = 0; c.loss_type = 0; return; }  // ICP.cpp:319,394
    const AdaptiveMEstimatorConfig& a = m_adaptive_estimator->get_config();
    c.use_adaptive_m_estimator = a.use_adaptive_m_estimator ? 1 : 0;
    c.loss_type = (a.loss_type == "cauchy") ? 1 : 0;
    c.min_scale_factor = a.min_scale_factor;
    c.max_scale_factor = a.max_scale_factor;
    c.num_alpha_segments = a.num_alpha_segments;
    c.truncated_threshold = a.truncated_threshold;
    c.gmm_components = a.gmm_components;
    c.gmm_sample_size = a.gmm_sample_size;
    c.pko_kernel_type = (a.pko_kernel_type == "cauchy") ? 1 : 0;  // "huber" in both shipped configs (kitti.yaml:51)
  }
  ICPConfig m_config;
  std::shared_ptr<optimization::AdaptiveMEstimator> m_adaptive_estimator;
  OptimizationStats m_last_stats;
};

#ifndef B2LO_KEEP_REFERENCE_LOOP_ICP
namespace b2lo_detail_loop {
// the loop/PGO thread's own context: optimize_loop runs concurrently with optimize() (Estimator.cpp:1001)
inline b2lo_ctx* context() {
  static b2lo_ctx* ctx = [] {
    b2lo_ctx* c = nullptr;
    int dev = 0;
    if (const char* e = std::getenv("B2LO_DEVICE")) dev = std::atoi(e);
    if (b2lo_ctx_create(dev, &c) != B2LO_OK) throw std::runtime_error(std::string("b2lo: ") + b2lo_last_error());
    return c;
  }();
  return ctx;
}
}  // namespace b2lo_detail_loop
// ICP.cpp:40-251: relative = curr_pose^-1 * optimised curr pose; true only if the loop converged and >= half of the points are inliers.
// The keyframes are only read (the reference works on deep copies for the same effect).
inline bool IterativeClosestPointOptimizer::optimize_loop(std::shared_ptr<database::LidarFrame> curr_keyframe,
                                                          std::shared_ptr<database::LidarFrame> matched_keyframe,
                                                          SE3f& optimized_relative_transform, float& inlier_ratio) {
  inlier_ratio = 0.0f;
  if (!curr_keyframe || !matched_keyframe) return false;
  if (m_adaptive_estimator) m_adaptive_estimator->reset();   // :52-55
  auto cloud_of = [](const std::shared_ptr<database::LidarFrame>& f) {
    util::PointCloudConstPtr c = f->get_feature_cloud();     // get_frame_cloud (:769-783)
    if (!c || c->empty()) c = f->get_processed_cloud();
    return c;
  };
  util::PointCloudConstPtr cc = cloud_of(curr_keyframe), mc = cloud_of(matched_keyframe);
  if (!cc || cc->empty() || !mc || mc->empty()) return false;
  b2lo_icp_cfg cfg;
  fill_cfg(cfg);
  float Tc[16], Tm[16], Trel[16];
  b2lo_detail::to_row_major(curr_keyframe->get_pose().Matrix(), Tc);
  b2lo_detail::to_row_major(matched_keyframe->get_pose().Matrix(), Tm);
  b2lo_icp_stats st;
  const int rc = b2lo_icp_optimize_loop(b2lo_detail_loop::context(), b2lo_detail::data(*cc), cc->size(), b2lo_detail::stride_floats(), Tc,
                                        b2lo_detail::data(*mc), mc->size(), b2lo_detail::stride_floats(), Tm, &cfg, Trel, &inlier_ratio, &st);
  if (rc < 0) throw std::runtime_error(std::string("b2lo_icp_optimize_loop: ") + b2lo_last_error());
  optimized_relative_transform = SE3f(b2lo_detail::from_row_major(Trel));
  return rc == B2LO_OK;
}
#endif

}  // namespace optimization

// ---- final-map export and scan ingest -------------------------------------------------------------
namespace b2lo {

// util::VoxelGrid (src/util/PointCloudUtils.h:462-557) with the same three calls; Estimator::save_map_to_ply (Estimator.cpp:1290-1298)
// uses it as `lidar_slam::b2lo::VoxelGrid voxel_filter;` in place of `util::VoxelGrid voxel_filter;`
class VoxelGrid {
 public:
  VoxelGrid() : leaf_size_(0.01f) {}
  void setLeafSize(float size) { leaf_size_ = size; }
  void setInputCloud(const util::PointCloudConstPtr& cloud) { input_cloud_ = cloud; }
  void filter(util::PointCloud& output) {
    output.clear();
    if (!input_cloud_ || input_cloud_->empty() || leaf_size_ <= 0) return;
    std::vector<float> buf(input_cloud_->size() * 3);
    size_t m = 0;
    int rc = b2lo_voxel_grid_filter(b2lo_detail::context(), b2lo_detail::data(*input_cloud_), input_cloud_->size(), b2lo_detail::stride_floats(),
                                    leaf_size_, buf.data(), input_cloud_->size(), &m);
    if (rc < 0) throw std::runtime_error(std::string("b2lo_voxel_grid_filter: ") + b2lo_last_error());
    output.reserve(m);
    for (size_t i = 0; i < m; ++i) output.push_back(util::Point3D(buf[i * 3], buf[i * 3 + 1], buf[i * 3 + 2]));
  }

 private:
  float leaf_size_;
  util::PointCloudConstPtr input_cloud_;
};

// preprocess_frame (Estimator.cpp:561-589) straight from a scan FILE IMAGE: a KITTI .bin (util::load_kitti_binary) or a binary PLY
// (PLYPlayer::load_ply_point_cloud) is downsampled by K1 reading the file's own records; no util::PointCloud of the raw scan is built.
// Returns false for the files the reference's loaders reject; ASCII PLY bodies are parsed on the host first.
inline bool filter_scan_file_image(const void* image, size_t bytes, bool is_ply, float voxel_size, int stride, util::PointCloud& output) {
  output.clear();
  b2lo_record_fmt fmt;
  size_t n_records = 0, data_offset = 0;
  std::vector<float> ascii;
  const unsigned char* body = static_cast<const unsigned char*>(image);
  if (is_ply) {
    size_t vertices = 0;
    int is_binary = 0;
    if (b2lo_ply_parse_header(image, bytes, &fmt, &vertices, &data_offset, &is_binary, &n_records) != B2LO_OK) return false;
    if (!is_binary) {
      ascii.resize(vertices * 3);
      if (b2lo_ply_read_ascii(image, bytes, ascii.data(), vertices, &n_records) != B2LO_OK) return false;
      fmt.record_bytes = 12; fmt.off_x = 0; fmt.off_y = 4; fmt.off_z = 8;
      body = reinterpret_cast<const unsigned char*>(ascii.data());
      data_offset = 0;
    }
  } else {
    b2lo_kitti_record_fmt(&fmt);
    n_records = bytes / fmt.record_bytes;
  }
  if (n_records == 0 || stride < 1) return true;
  const size_t ns = (n_records + (size_t)stride - 1) / (size_t)stride;
  std::vector<float> buf(ns * 3);
  size_t m = 0;
  int rc = b2lo_filter_records(b2lo_detail::context(), body + data_offset, n_records, &fmt, stride, voxel_size, buf.data(), nullptr, &m);
  if (rc < 0) throw std::runtime_error(std::string("b2lo_filter_records: ") + b2lo_last_error());
  output.reserve(m);
  for (size_t i = 0; i < m; ++i) output.push_back(util::Point3D(buf[i * 3], buf[i * 3 + 1], buf[i * 3 + 2]));
  return true;
}

}  // namespace b2lo
}  // namespace lidar_slam
