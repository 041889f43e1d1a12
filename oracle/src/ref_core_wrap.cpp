// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the REAL reference classes, compiled by oracle/Makefile together
// with the UNMODIFIED reference translation units
//   /root/reference/src/database/VoxelMap.cpp, src/database/LidarFrame.cpp, src/util/MathUtils.cpp,
//   src/util/PointCloudUtils.cpp, src/optimization/IterativeClosestPointOptimizer.cpp, src/optimization/AdaptiveMEstimator.cpp
// against oracle/eigen_compat (Eigen3 is absent from the image) into oracle/_ref/libref_core.so.
// The entry points mirror oracle/include/orc_capi.h one for one (ref_* instead of orc_*), so tests/test_oracle_pins.py can run
// the restatement and the reference on the same inputs and compare bits.  This file is built with -fno-access-control: it
// reads the private containers of map::VoxelMap and calls the private find_correspondences* members; the reference sources
// themselves are compiled as they are.
#include <chrono>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>
#include "database/LidarFrame.h"
#include "database/VoxelMap.h"
#include "optimization/AdaptiveMEstimator.h"
#include "optimization/IterativeClosestPointOptimizer.h"
#include "util/LogUtils.h"
#include "util/MathUtils.h"
#include "util/PointCloudUtils.h"
#include "orc_capi.h"

using namespace lidar_slam;
using lidar_slam::util::Point3D;
using lidar_slam::util::PointCloud;
using lidar_slam::util::SE3f;
using lidar_slam::util::SO3f;
namespace rp = lidar_slam::optimization;

namespace {

std::shared_ptr<PointCloud> make_cloud(const float* xyz, size_t n) {
  auto c = std::make_shared<PointCloud>();
  c->reserve(n);
  for (size_t i = 0; i < n; ++i) { Point3D p; p.x = xyz[i * 3]; p.y = xyz[i * 3 + 1]; p.z = xyz[i * 3 + 2]; c->push_back(p); }
  return c;
}
// pose taken VERBATIM from a row-major 4x4 (no SVD re-projection), like orc_capi's raw_se3
SE3f raw_se3(const float* T) {
  SE3f p;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) p.Rotation().Matrix()(i, j) = T[i * 4 + j]; p.Translation()(i) = T[i * 4 + 3]; }
  return p;
}
void put_se3(const SE3f& p, float* T) {
  Eigen::Matrix4f M = p.Matrix();
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T[i * 4 + j] = M(i, j);
}
rp::ICPConfig to_icp(const orc_icp_cfg* c) {
  rp::ICPConfig o;
  o.max_iterations = c->max_iterations; o.translation_tolerance = c->translation_tolerance; o.rotation_tolerance = c->rotation_tolerance;
  o.max_correspondence_distance = c->max_correspondence_distance; o.min_correspondence_points = c->min_correspondence_points;
  o.use_robust_loss = c->use_robust_loss != 0; o.robust_loss_delta = c->robust_loss_delta;
  o.use_surfel_correspondence = c->use_surfel_correspondence != 0;
  return o;
}
std::shared_ptr<optimization::AdaptiveMEstimator> to_pko(const orc_icp_cfg* c) {
  // the constructor call of Estimator.cpp:49-59
  return std::make_shared<optimization::AdaptiveMEstimator>(c->use_adaptive_m_estimator != 0, c->loss_type == 1 ? "cauchy" : "huber",
                                                            c->min_scale_factor, c->max_scale_factor, c->num_alpha_segments, c->truncated_threshold,
                                                            c->gmm_components, c->gmm_sample_size, c->pko_kernel_type == 1 ? "cauchy" : "huber");
}
std::shared_ptr<database::LidarFrame> make_frame(int id, const float* local_xyz, size_t m) {
  auto cloud = make_cloud(local_xyz, m);
  auto f = std::make_shared<database::LidarFrame>(id, 0.0, cloud);
  f->set_processed_cloud(cloud);  // Estimator.cpp:581-582: both are the downsampled cloud
  f->set_feature_cloud(cloud);
  return f;
}
// the LDLT hook of oracle/eigen_compat (H, -g and the solution of every Gauss-Newton iteration) -> orc_iter_trace records
void begin_trace() { Eigen::compat::ldlt_trace().clear(); Eigen::compat::ldlt_trace_enabled() = true; }
void end_trace(orc_iter_trace* trace, int cap, int* n_trace) {
  Eigen::compat::ldlt_trace_enabled() = false;
  auto& t = Eigen::compat::ldlt_trace();
  int k = 0;
  for (; k < (int)t.size() && k < cap; ++k) {
    orc_iter_trace& d = trace[k];
    std::memset(&d, 0, sizeof d);
    for (int i = 0; i < 36; ++i) d.H[i] = (float)t[k].A[i];
    for (int i = 0; i < 6; ++i) { d.g[i] = -(float)t[k].b[i]; d.dx[i] = (float)t[k].x[i]; }
  }
  if (n_trace) *n_trace = k;
  t.clear();
}

}  // namespace

// the reference logs at INFO by default (LogUtils.h:55); keep test output readable: level 4 is above ERROR (silent)
static const bool g_quiet = [] { lidar_slam::Logger::level = static_cast<lidar_slam::LogLevel>(4); return true; }();

extern "C" {

void ref_set_log_level(int level) { lidar_slam::Logger::level = static_cast<lidar_slam::LogLevel>(level); }

// ---- keys ------------------------------------------------------------------------------------------------------------------
uint64_t ref_voxel_key_hash(int x, int y, int z) { return (uint64_t)map::VoxelKeyHash()(map::VoxelKey(x, y, z)); }
void ref_point_to_key(const float* p, float voxel, int factor, int level, int* key) {
  map::VoxelMap m(voxel); m.SetHierarchyFactor(factor);
  map::VoxelKey k = m.PointToVoxelKey(Eigen::Vector3f(p[0], p[1], p[2]), level);
  key[0] = k.x; key[1] = k.y; key[2] = k.z;
}
void ref_parent_key(const int* key, int factor, int* parent) {
  map::VoxelMap m(0.5f); m.SetHierarchyFactor(factor);
  map::VoxelKey k = m.GetParentKey(map::VoxelKey(key[0], key[1], key[2]));
  parent[0] = k.x; parent[1] = k.y; parent[2] = k.z;
}

// ---- FastVoxelFilter (VoxelMap.h:53-143) -------------------------------------------------------------------------------------
void ref_filter(const float* xyz, size_t n, int stride, float voxel, float* out_xyz, size_t* m) {
  map::FastVoxelFilter f(voxel);
  auto in = make_cloud(xyz, n);
  PointCloud out;
  f.filter(*in, out, stride);
  for (size_t i = 0; i < out.size(); ++i) { out_xyz[i * 3] = out[i].x; out_xyz[i * 3 + 1] = out[i].y; out_xyz[i * 3 + 2] = out[i].z; }
  *m = out.size();
}

// ---- util::VoxelGrid (PointCloudUtils.h:462-557), load_kitti_binary (PointCloudUtils.cpp:19-65) ---------------------------------
int ref_voxel_grid_filter(const float* xyz, size_t n, float leaf, float* out_xyz, size_t cap, size_t* m) {
  util::VoxelGrid g;
  g.setLeafSize(leaf);
  g.setInputCloud(make_cloud(xyz, n));
  PointCloud out;
  g.filter(out);
  *m = out.size();
  if (out.size() > cap) return -1;
  for (size_t i = 0; i < out.size(); ++i) { out_xyz[i * 3] = out[i].x; out_xyz[i * 3 + 1] = out[i].y; out_xyz[i * 3 + 2] = out[i].z; }
  return 0;
}
int ref_kitti_load_file(const char* path, float* out_xyz, size_t cap, size_t* n) {
  auto cloud = util::load_kitti_binary(path);
  if (!cloud) { *n = 0; return 1; }
  *n = cloud->size();
  if (cloud->size() > cap) return -1;
  for (size_t i = 0; i < cloud->size(); ++i) { out_xyz[i * 3] = (*cloud)[i].x; out_xyz[i * 3 + 1] = (*cloud)[i].y; out_xyz[i * 3 + 2] = (*cloud)[i].z; }
  return 0;
}
void ref_transform_point_cloud(const float* xyz, size_t n, const float* T16, float* out_xyz) {
  auto in = make_cloud(xyz, n);
  auto out = std::make_shared<PointCloud>();
  Eigen::Matrix4f T;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T16[i * 4 + j];
  util::transform_point_cloud(in, out, T);
  for (size_t i = 0; i < out->size(); ++i) { out_xyz[i * 3] = (*out)[i].x; out_xyz[i * 3 + 1] = (*out)[i].y; out_xyz[i * 3 + 2] = (*out)[i].z; }
}

// ---- map::VoxelMap ---------------------------------------------------------------------------------------------------------------
void* ref_map_create(float voxel, int factor, float planarity, int compute_surfels) {
  auto* m = new map::VoxelMap(voxel);
  m->SetHierarchyFactor(factor); m->SetPlanarityThreshold(planarity); m->SetComputeSurfels(compute_surfels != 0);
  return m;
}
void ref_map_destroy(void* h) { delete static_cast<map::VoxelMap*>(h); }
void ref_map_clear(void* h) { static_cast<map::VoxelMap*>(h)->Clear(); }
void ref_map_update(void* h, const float* xyz, size_t n, const double* sensor, double max_distance) {
  static_cast<map::VoxelMap*>(h)->UpdateVoxelMap(make_cloud(xyz, n), Eigen::Vector3d(sensor[0], sensor[1], sensor[2]), max_distance, true);
}
void ref_map_counts(void* h, size_t* l0, size_t* l1, size_t* surfels) {
  auto* m = static_cast<map::VoxelMap*>(h);
  if (l0) *l0 = m->GetVoxelCount();
  if (l1) *l1 = m->GetL1VoxelCount();
  if (surfels) *surfels = m->GetSurfelCount();
}
void ref_map_export_l0(void* h, int* keys, float* cent, int* counts) {
  auto* m = static_cast<map::VoxelMap*>(h);
  size_t i = 0;
  for (const auto& kv : m->m_voxels_L0) {
    if (keys) { keys[i * 3] = kv.first.x; keys[i * 3 + 1] = kv.first.y; keys[i * 3 + 2] = kv.first.z; }
    if (cent) for (int a = 0; a < 3; ++a) cent[i * 3 + a] = kv.second.centroid(a);
    if (counts) counts[i] = kv.second.point_count;
    ++i;
  }
}
// the public export the Estimator uses (VoxelMap.cpp:388-403); must equal the centroids of ref_map_export_l0
size_t ref_map_point_cloud(void* h, float* xyz, size_t cap) {
  auto c = static_cast<map::VoxelMap*>(h)->GetPointCloud();
  size_t k = std::min(cap, c->size());
  for (size_t i = 0; i < k; ++i) { xyz[i * 3] = (*c)[i].x; xyz[i * 3 + 1] = (*c)[i].y; xyz[i * 3 + 2] = (*c)[i].z; }
  return c->size();
}
void ref_map_export_l1(void* h, int* keys, int* nchild, int* children, int* has_surfel, float* normal, float* centroid, float* planarity,
                       int* last_child_count) {
  auto* m = static_cast<map::VoxelMap*>(h);
  size_t i = 0;
  for (const auto& kv : m->m_voxels_L1) {
    const auto& nd = kv.second;
    if (keys) { keys[i * 3] = kv.first.x; keys[i * 3 + 1] = kv.first.y; keys[i * 3 + 2] = kv.first.z; }
    if (nchild) nchild[i] = (int)nd.occupied_children.size();
    if (children) {
      size_t c = 0;
      for (const auto& ck : nd.occupied_children) { if (c >= 27) break; children[(i * 27 + c) * 3] = ck.x; children[(i * 27 + c) * 3 + 1] = ck.y; children[(i * 27 + c) * 3 + 2] = ck.z; ++c; }
    }
    if (has_surfel) has_surfel[i] = nd.has_surfel ? 1 : 0;
    if (normal) for (int a = 0; a < 3; ++a) normal[i * 3 + a] = nd.surfel_normal(a);
    if (centroid) for (int a = 0; a < 3; ++a) centroid[i * 3 + a] = nd.surfel_centroid(a);
    if (planarity) planarity[i] = nd.planarity_score;
    if (last_child_count) last_child_count[i] = nd.last_child_count;
    ++i;
  }
}
size_t ref_map_surfels(void* h, float* centroid, float* normal, float* planarity, size_t cap) {  // GetL1Surfels, VoxelMap.cpp:405-418
  auto s = static_cast<map::VoxelMap*>(h)->GetL1Surfels();
  size_t k = std::min(cap, s.size());
  for (size_t i = 0; i < k; ++i) {
    for (int a = 0; a < 3; ++a) { centroid[i * 3 + a] = std::get<0>(s[i])(a); normal[i * 3 + a] = std::get<1>(s[i])(a); }
    planarity[i] = std::get<2>(s[i]);
  }
  return s.size();
}
int ref_map_lookup(void* h, const float* p, float* n, float* c) {
  Eigen::Vector3f nn = Eigen::Vector3f::Zero(), cc = Eigen::Vector3f::Zero();
  bool ok = static_cast<map::VoxelMap*>(h)->GetSurfelAtPoint(Eigen::Vector3f(p[0], p[1], p[2]), nn, cc);
  for (int a = 0; a < 3; ++a) { n[a] = nn(a); c[a] = cc(a); }
  return ok ? 1 : 0;
}
void ref_map_transform_rehash(void* h, const float* T16) {
  Eigen::Matrix4f T;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T16[i * 4 + j];
  static_cast<map::VoxelMap*>(h)->ApplyTransformAndRehash(T);
}

// ---- IterativeClosestPointOptimizer ------------------------------------------------------------------------------------------------
// find_correspondences (ICP.cpp:587-645): the accepted list exactly as the reference stores it (80 B per correspondence)
size_t ref_icp_correspondence_list(void* map, const float* local_xyz, size_t m, const float* T16, double max_dist, int kdtree_mode,
                                   double* points_last, double* points_curr, double* normals_last, double* residuals, size_t cap) {
  rp::ICPConfig cfg; cfg.max_correspondence_distance = max_dist;
  rp::IterativeClosestPointOptimizer icp(cfg);
  auto frame = make_frame(0, local_xyz, m);
  frame->set_pose(raw_se3(T16));
  rp::DualFrameCorrespondences c;
  auto* vm = static_cast<map::VoxelMap*>(map);
  size_t n = kdtree_mode ? icp.find_correspondences_kdtree(vm, frame, c) : icp.find_correspondences(vm, frame, c);
  for (size_t i = 0; i < n && i < cap; ++i)
    for (int a = 0; a < 3; ++a) {
      points_last[i * 3 + a] = c.points_last[i](a); points_curr[i * 3 + a] = c.points_curr[i](a); normals_last[i * 3 + a] = c.normals_last[i](a);
    }
  for (size_t i = 0; i < n && i < cap; ++i) residuals[i] = c.residuals[i];
  return n;
}
void ref_map_rebuild_kdtree(void* h) { static_cast<map::VoxelMap*>(h)->RebuildKdTree(); }

// optimize (ICP.cpp:255-463).  stats6 = {num_correspondences, num_iterations, initial_cost, final_cost, converged, frame pose changed}
int ref_icp_optimize(void* map, const float* local_xyz, size_t m, const float* T_init16, const orc_icp_cfg* cfg, float* T_out16,
                     orc_iter_trace* trace, int trace_cap, int* n_trace, double* stats6, float* frame_pose16) {
  rp::IterativeClosestPointOptimizer icp(to_icp(cfg), to_pko(cfg));
  auto frame = make_frame(0, local_xyz, m);
  SE3f init = raw_se3(T_init16), out;
  if (trace) begin_trace();
  bool ok = icp.optimize(static_cast<map::VoxelMap*>(map), frame, init, out);
  if (trace) end_trace(trace, trace_cap, n_trace);
  put_se3(out, T_out16);
  if (stats6) {
    const auto& s = icp.get_last_stats();
    stats6[0] = (double)s.num_correspondences; stats6[1] = (double)s.num_iterations; stats6[2] = s.initial_cost; stats6[3] = s.final_cost;
    stats6[4] = s.converged ? 1.0 : 0.0; stats6[5] = 0.0;
  }
  if (frame_pose16) put_se3(frame->get_pose(), frame_pose16);
  return ok ? 1 : 0;
}

// optimize_loop (ICP.cpp:40-251)
int ref_icp_optimize_loop(const float* curr_xyz, size_t m_curr, const float* T_curr16, const float* matched_xyz, size_t m_matched,
                          const float* T_matched16, const orc_icp_cfg* cfg, float* T_rel16, float* inlier_ratio, int* iterations,
                          orc_iter_trace* trace, int trace_cap, int* n_trace) {
  rp::IterativeClosestPointOptimizer icp(to_icp(cfg), to_pko(cfg));
  auto curr = make_frame(1, curr_xyz, m_curr);
  auto matched = make_frame(0, matched_xyz, m_matched);
  curr->set_pose(raw_se3(T_curr16));
  matched->set_pose(raw_se3(T_matched16));
  SE3f rel;
  float ratio = 0.0f;
  int nt = 0;
  begin_trace();
  std::vector<orc_iter_trace> local(trace ? 0 : 128);
  bool ok = icp.optimize_loop(curr, matched, rel, ratio);
  end_trace(trace ? trace : local.data(), trace ? trace_cap : 128, &nt);
  if (T_rel16) put_se3(rel, T_rel16);
  if (inlier_ratio) *inlier_ratio = ratio;
  if (iterations) *iterations = nt;   // one LDLT solve per Gauss-Newton iteration
  if (n_trace) *n_trace = nt;
  return ok ? 1 : 0;
}

// ---- util::SO3 / SE3 (MathUtils.cpp:23-181) ------------------------------------------------------------------------------------------
void ref_so3_normalize(const float* R, float* out) {
  Eigen::Matrix3f M;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M(i, j) = R[i * 3 + j];
  SO3f r(M);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[i * 3 + j] = r.Matrix()(i, j);
}
void ref_so3_exp(const float* w, float* out) {
  SO3f r = SO3f::Exp(Eigen::Vector3f(w[0], w[1], w[2]));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[i * 3 + j] = r.Matrix()(i, j);
}
void ref_so3_log(const float* R, float* out) {
  SO3f r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.Matrix()(i, j) = R[i * 3 + j];
  Eigen::Vector3f w = r.Log();
  for (int i = 0; i < 3; ++i) out[i] = w(i);
}
void ref_se3_mul(const float* A16, const float* B16, float* C16) { put_se3(raw_se3(A16) * raw_se3(B16), C16); }
void ref_se3_inv(const float* A16, float* C16) { put_se3(raw_se3(A16).Inverse(), C16); }
void ref_svd3f(const float* A, float* U, float* S, float* V) {
  Eigen::Matrix3f M;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M(i, j) = A[i * 3 + j];
  Eigen::JacobiSVD<Eigen::Matrix3f> svd(M, Eigen::ComputeFullU | Eigen::ComputeFullV);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) { U[i * 3 + j] = svd.matrixU()(i, j); V[i * 3 + j] = svd.matrixV()(i, j); } S[i] = svd.singularValues()(i); }
}
// smallest right-singular vector of an n x 3 f64 matrix through JacobiSVD<MatrixXd>(A, ComputeFullV).matrixV().col(2)  (ICP.cpp:745-746)
void ref_plane_normal_nx3(const double* A, int n, double* normal) {
  Eigen::MatrixXd M(n, 3);
  for (int i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) M(i, j) = A[i * 3 + j];
  Eigen::JacobiSVD<Eigen::MatrixXd> svd(M, Eigen::ComputeFullV);
  Eigen::Vector3d v = svd.matrixV().col(2);
  for (int j = 0; j < 3; ++j) normal[j] = v(j);
}

}  // extern "C"

// ---- the per-scan driver over the REAL reference classes --------------------------------------------------------------------------------
// processing::Estimator itself does not compile here (it pulls in the loop detector, the pose graph and their third-party
// dependencies), so the driver below restates ONLY its control flow - which reference function is called with which argument, in
// Estimator.cpp's order (process_frame :116-233, initialize_first_frame :235-269, estimate_motion_dual_frame :271-320,
// should_create_keyframe :349-368, create_keyframe :449-470, preprocess_frame :561-589) - while every function on the hot path
// (FastVoxelFilter::filter, IterativeClosestPointOptimizer::optimize, transform_point_cloud, VoxelMap::UpdateVoxelMap /
// RebuildKdTree / GetPointCloud, SE3f algebra) is the reference's own compiled code.  Used (a) to pin oracle/include/orc_pipeline.hpp
// pose for pose and (b) as bench.py's CPU baseline of kind "reference".
namespace {
struct RefPipe {
  orc_pipe_cfg cfg;
  std::unique_ptr<map::FastVoxelFilter> filter;
  std::unique_ptr<map::VoxelMap> vmap;
  std::shared_ptr<optimization::AdaptiveMEstimator> pko;
  std::unique_ptr<rp::IterativeClosestPointOptimizer> icp;
  std::shared_ptr<database::LidarFrame> previous, last_keyframe;
  SE3f T_wl, velocity, last_keyframe_pose;
  bool initialized = false, last_kf = false, last_ok = false;
  int n_keyframes = 0, next_id = 0;
  double times[4] = {0, 0, 0, 0};
  size_t n_features = 0;

  explicit RefPipe(const orc_pipe_cfg& c) : cfg(c) {   // Estimator.cpp:48-81
    filter.reset(new map::FastVoxelFilter(c.voxel_size));
    vmap.reset(new map::VoxelMap(c.map_voxel_size));
    vmap->SetHierarchyFactor(3);
    vmap->SetPlanarityThreshold(c.surfel_planarity_threshold);
    vmap->SetComputeSurfels(c.icp.use_surfel_correspondence != 0);
    pko = to_pko(&c.icp);
    icp.reset(new rp::IterativeClosestPointOptimizer(to_icp(&c.icp), pko));
  }
  void create_keyframe(const std::shared_ptr<database::LidarFrame>& f) {   // :449-470
    Eigen::Vector3f pos = f->get_pose().Translation();
    Eigen::Vector3d sensor = pos.cast<double>();
    vmap->UpdateVoxelMap(f->get_feature_cloud_global(), sensor, cfg.max_range * 1.2, true);
    if (!cfg.icp.use_surfel_correspondence) vmap->RebuildKdTree();
    f->set_local_map(vmap->GetPointCloud());
    f->set_keyframe_id(n_keyframes);             // :376 - from now on get_pose() of this frame is its stored pose
    last_keyframe = f;
    last_keyframe_pose = f->get_pose();
    ++n_keyframes;
    last_kf = true;
  }
  bool process(const float* xyz, size_t n, size_t stride) {
    using clk = std::chrono::high_resolution_clock;
    auto t0 = clk::now();
    last_kf = last_ok = false;
    times[0] = times[1] = times[2] = times[3] = 0.0;
    auto raw = std::make_shared<PointCloud>();
    raw->reserve(n);
    for (size_t i = 0; i < n; ++i) { Point3D p; p.x = xyz[i * stride]; p.y = xyz[i * stride + 1]; p.z = xyz[i * stride + 2]; raw->push_back(p); }
    auto frame = std::make_shared<database::LidarFrame>(next_id++, 0.0, raw);
    auto ta = clk::now();                        // the cloud conversion above is the player's job, not the estimator's: outside the stage times
    auto down = std::make_shared<PointCloud>();  // preprocess_frame :561-589
    filter->filter(*raw, *down, cfg.point_stride);
    n_features = down->size();
    auto t1 = clk::now();
    times[0] = std::chrono::duration<double, std::milli>(t1 - ta).count();
    if (down->empty()) return false;
    frame->set_processed_cloud(down);
    frame->set_feature_cloud(down);
    if (!initialized) {                          // initialize_first_frame :235-269
      T_wl = frame->get_initial_pose();
      velocity = SE3f();
      frame->set_pose(T_wl);
      auto world = std::make_shared<PointCloud>();
      Eigen::Matrix4f M = T_wl.Matrix();
      util::transform_point_cloud(down, world, M);
      frame->set_feature_cloud_global(world);
      create_keyframe(frame);
      previous = frame;
      initialized = true;
      auto t2 = clk::now();
      times[2] = std::chrono::duration<double, std::milli>(t2 - t1).count();
      times[3] = std::chrono::duration<double, std::milli>(t2 - ta).count();
      (void)t0;
      return true;
    }
    SE3f guess = previous->get_pose() * velocity;                      // :154
    SE3f result = guess;
    auto local_map = last_keyframe ? last_keyframe->get_local_map() : nullptr;
    if (local_map && !local_map->empty()) {                            // estimate_motion_dual_frame :271-320
      SE3f init(guess.RotationMatrix(), guess.Translation()), opt;
      last_ok = icp->optimize(vmap.get(), frame, init, opt);
      if (last_ok) result = SE3f(opt.RotationMatrix(), opt.Translation());
    }
    auto t2 = clk::now();
    times[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    auto world = std::make_shared<PointCloud>();
    Eigen::Matrix4f M = result.Matrix();
    util::transform_point_cloud(down, world, M);                       // :163-166
    frame->set_feature_cloud_global(world);
    T_wl = result;
    velocity = previous->get_pose().Inverse() * T_wl;                  // :177
    frame->set_pose(T_wl);
    if (last_keyframe) {                                               // :186-191: a non-keyframe's get_pose() is keyframe pose * relative pose
      frame->set_previous_keyframe(last_keyframe);
      frame->set_relative_pose(last_keyframe->get_stored_pose().Inverse() * T_wl);
    }
    bool kf = n_keyframes == 0;                                        // should_create_keyframe :349-368
    if (!kf) {
      Eigen::Vector3f d = T_wl.Translation() - last_keyframe_pose.Translation();
      double distance = d.norm();
      SO3f rd = last_keyframe_pose.Rotation().Inverse() * T_wl.Rotation();
      double angle = rd.Log().norm();
      kf = distance > cfg.keyframe_distance_threshold || angle > cfg.keyframe_rotation_threshold;
    }
    if (kf) create_keyframe(frame);
    if (previous && previous != last_keyframe) previous->clear_non_keyframe_data();   // :214-219
    previous = frame;
    auto t3 = clk::now();
    times[2] = std::chrono::duration<double, std::milli>(t3 - t2).count();
    times[3] = std::chrono::duration<double, std::milli>(t3 - ta).count();
    return true;
  }
};
}  // namespace

extern "C" {
void* ref_pipe_create(const orc_pipe_cfg* cfg) { return new RefPipe(*cfg); }
void ref_pipe_destroy(void* h) { delete static_cast<RefPipe*>(h); }
int ref_pipe_process(void* h, const float* xyz, size_t n, size_t stride_floats, float* pose16, int* flags, double* times_ms, int* n_features,
                     int* n_corr, int* n_iters) {
  RefPipe* p = static_cast<RefPipe*>(h);
  bool ok = p->process(xyz, n, stride_floats);
  if (pose16) put_se3(p->T_wl, pose16);
  if (flags) *flags = (p->last_kf ? 1 : 0) | (p->last_ok ? 2 : 0);
  if (times_ms) for (int i = 0; i < 4; ++i) times_ms[i] = p->times[i];
  if (n_features) *n_features = (int)p->n_features;
  if (n_corr) *n_corr = (int)p->icp->get_last_stats().num_correspondences;
  if (n_iters) *n_iters = (int)p->icp->get_last_stats().num_iterations;
  return ok ? 1 : 0;
}
void* ref_pipe_map(void* h) { return static_cast<RefPipe*>(h)->vmap.get(); }
}
