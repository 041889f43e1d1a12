// ORACLE - TEST INFRASTRUCTURE ONLY.  The reference's OWN per-scan driver: the UNMODIFIED /root/reference/src/processing/Estimator.cpp
// compiled by oracle/Makefile into oracle/_ref/libref_estimator.so, so that oracle/include/orc_pipeline.hpp (the restated control flow
// of process_frame) can be compared with processing::Estimator::process_frame itself, scan after scan.
//
// Estimator.cpp needs, besides the classes already in libref_core.so, the loop detector (LidarIris -> OpenCV) and the pose graph
// (Eigen/Sparse).  Their HEADERS compile against the type-only stand-ins of oracle/ref_stubs (oracle/ref_stubs/thirdparty: opencv2/opencv.hpp, Eigen/Sparse); their
// translation units are not built: the eleven members Estimator.cpp references are defined below as no-ops, and the driver is run with
// loop detection and pose-graph optimisation switched off (SystemConfig::loop_enable_loop_detection / pgo_enable_pgo = false), which is
// also the configuration of the hot path this repository accelerates.  Nothing below touches the arithmetic of a scan.
//
// The same file is built a second time into oracle/_ref/libref_estimator_gpu.so: there Estimator.cpp is compiled through a header
// OVERLAY (oracle/_ref/overlay: symlinks to the reference tree, except database/VoxelMap.h and
// optimization/IterativeClosestPointOptimizer.h, which are the one-line `#include "b2lo_dropin.h"` of INTEGRATION.md) and linked with
// libb2lo.so - the reference's own, unmodified Estimator driving the CUDA engine through the drop-in shim.
#include <chrono>
#include <cstring>
#include <memory>
#include "database/LidarFrame.h"
#include "database/VoxelMap.h"
#include "optimization/PoseGraphOptimizer.h"
#include "processing/Estimator.h"
#include "processing/LoopClosureDetector.h"
#include "util/ConfigUtils.h"
#include "util/LogUtils.h"
#include "orc_capi.h"

using namespace lidar_slam;

// ---- no-op definitions of the loop detector / pose graph members Estimator.cpp links against ---------------------------------------
namespace lidar_slam {
namespace processing {
LoopClosureDetector::LoopClosureDetector(const LoopClosureConfig& config) : m_config(config) {}
LoopClosureDetector::~LoopClosureDetector() {}
bool LoopClosureDetector::add_keyframe(std::shared_ptr<database::LidarFrame>) { return false; }
std::vector<LoopCandidate> LoopClosureDetector::detect_loop_closures(std::shared_ptr<database::LidarFrame>) { return {}; }
void LoopClosureDetector::update_config(const LoopClosureConfig& config) { m_config = config; }
}  // namespace processing
namespace optimization {
PoseGraphOptimizer::PoseGraphOptimizer() {}
PoseGraphOptimizer::~PoseGraphOptimizer() {}
bool PoseGraphOptimizer::add_first_keyframe(int, const util::SE3f&) { return false; }
bool PoseGraphOptimizer::add_keyframe_with_odom(int, int, const util::SE3f&, const util::SE3f&, double, double) { return false; }
bool PoseGraphOptimizer::add_loop_and_optimize(int, int, const util::SE3f&, double, double) { return false; }
std::map<int, util::SE3f> PoseGraphOptimizer::get_all_optimized_poses() const { return {}; }
}  // namespace optimization
}  // namespace lidar_slam

namespace {
struct RefEstimator {
  std::unique_ptr<processing::Estimator> est;
  int next_id = 0;
  size_t n_features = 0;
  bool last_kf = false;
  double last_ms = 0.0;   // wall time of the last process_frame call alone (the cloud conversion around it is the player's job)
};
util::SystemConfig to_system_config(const orc_pipe_cfg* c) {   // the fields Estimator.cpp:30-105 reads, from the oracle's pipeline config
  util::SystemConfig s;
  s.voxel_size = c->voxel_size; s.point_stride = c->point_stride; s.map_voxel_size = c->map_voxel_size; s.max_range = (float)c->max_range;
  s.surfel_planarity_threshold = c->surfel_planarity_threshold;
  s.keyframe_distance_threshold = c->keyframe_distance_threshold; s.keyframe_rotation_threshold = c->keyframe_rotation_threshold;
  s.max_iterations = (size_t)c->icp.max_iterations; s.translation_threshold = c->icp.translation_tolerance; s.rotation_threshold = c->icp.rotation_tolerance;
  s.max_correspondence_distance = (float)c->icp.max_correspondence_distance;
  s.use_surfel_correspondence = c->icp.use_surfel_correspondence != 0;
  s.use_adaptive_m_estimator = c->icp.use_adaptive_m_estimator != 0;
  s.loss_type = c->icp.loss_type == 1 ? "cauchy" : "huber";
  s.min_scale_factor = c->icp.min_scale_factor; s.max_scale_factor = c->icp.max_scale_factor; s.num_alpha_segments = c->icp.num_alpha_segments;
  s.truncated_threshold = c->icp.truncated_threshold; s.gmm_components = c->icp.gmm_components; s.gmm_sample_size = c->icp.gmm_sample_size;
  s.pko_kernel_type = c->icp.pko_kernel_type == 1 ? "cauchy" : "huber";
  s.loop_enable_loop_detection = false;
  s.pgo_enable_pgo = false;
  s.output_save_map = false;
  return s;
}
}  // namespace

static const bool g_quiet = [] { lidar_slam::Logger::level = static_cast<lidar_slam::LogLevel>(4); return true; }();

extern "C" {
void* ref_est_create(const orc_pipe_cfg* cfg) {
  auto* r = new RefEstimator();
  r->est.reset(new processing::Estimator(to_system_config(cfg)));
  return r;
}
void ref_est_destroy(void* h) { delete static_cast<RefEstimator*>(h); }
// one scan through processing::Estimator::process_frame.  flags bit0 = a keyframe was created; n_features = size of the frame's feature cloud
int ref_est_process(void* h, const float* xyz, size_t n, size_t stride_floats, float* pose16, int* flags, int* n_features) {
  RefEstimator* r = static_cast<RefEstimator*>(h);
  auto raw = std::make_shared<util::PointCloud>();
  raw->reserve(n);
  for (size_t i = 0; i < n; ++i) { util::Point3D p; p.x = xyz[i * stride_floats]; p.y = xyz[i * stride_floats + 1]; p.z = xyz[i * stride_floats + 2]; raw->push_back(p); }
  auto frame = std::make_shared<database::LidarFrame>(r->next_id++, 0.1 * r->next_id, raw);
  const size_t kf_before = r->est->get_keyframe_count();
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = r->est->process_frame(frame);
  r->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  r->last_kf = r->est->get_keyframe_count() > kf_before;
  auto fc = frame->get_feature_cloud();
  r->n_features = fc ? fc->size() : 0;
  if (pose16) { Eigen::Matrix4f M = r->est->get_current_pose().Matrix(); for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) pose16[i * 4 + j] = M(i, j); }
  if (flags) *flags = r->last_kf ? 1 : 0;
  if (n_features) *n_features = (int)r->n_features;
  return ok ? 1 : 0;
}
// the estimator's voxel map (a lidar_slam::map::VoxelMap of libref_core.so: usable with the ref_map_* entry points; in the
// libref_estimator_gpu.so build it is the drop-in shim's class and only the counts below are meaningful across the boundary)
void* ref_est_map(void* h) { return static_cast<RefEstimator*>(h)->est->get_voxel_map(); }
double ref_est_last_process_ms(void* h) { return static_cast<RefEstimator*>(h)->last_ms; }
void ref_est_counts(void* h, size_t* l0, size_t* l1, size_t* surfels) {
  auto* m = static_cast<RefEstimator*>(h)->est->get_voxel_map();
  *l0 = m->GetVoxelCount(); *l1 = m->GetL1VoxelCount(); *surfels = m->GetSurfelCount();
}
}
