// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_pipeline.hpp — the per-scan driver around the hot path, restating the in-scope parts of
// processing::Estimator (/root/reference/src/processing/Estimator.cpp):
//   ctor wiring of ICPConfig / PKO / VoxelMap           :48-81
//   process_frame                                      :116-233
//   initialize_first_frame                             :235-269
//   estimate_motion_dual_frame                         :271-320
//   should_create_keyframe                             :349-368
//   create_keyframe (map update, kd-tree, L0 export)   :370-472
//   preprocess_frame                                   :561-589
// Loop closure, PGO, viewer and LidarFrame bookkeeping are out of scope.
#pragma once
#include <chrono>
#include <memory>
#include <vector>
#include "orc_icp.hpp"

namespace orc {

struct PipelineConfig {  // config/kitti.yaml defaults; mid360.yaml: voxel 0.4, stride 4, use_surfel=false
  float voxel_size = 0.5f;
  int point_stride = 8;
  float map_voxel_size = 0.5f;
  double max_range = 100.0;
  float surfel_planarity_threshold = 0.1f;
  double keyframe_distance_threshold = 1.0;
  double keyframe_rotation_threshold = 0.3;
  ICPConfig icp;
  PkoConfig pko;
};

struct StageTimes { double preprocess_ms = 0, icp_ms = 0, map_update_ms = 0, total_ms = 0; };

class Pipeline {
 public:
  PipelineConfig cfg;
  FastVoxelFilter filter;
  VoxelMap map;
  std::shared_ptr<AdaptiveMEstimator> ame;
  ICPOptimizer icp;
  KdTree kdtree;
  bool has_kdtree = false;
  std::vector<P3> kdtree_cloud;
  std::vector<P3> local_map;       // GetPointCloud() of the last keyframe
  std::vector<P3> feature_cloud;   // downsampled scan (sensor frame)
  std::vector<P3> feature_world;   // same, at the optimised pose
  SE3f pose, prev_pose, velocity, last_keyframe_pose;
  // m_previous_frame->get_pose() (LidarFrame.cpp:113-128): a keyframe returns its stored pose; any other frame returns
  // previous_keyframe.pose * relative_pose, the relative pose being last_keyframe.pose^-1 * pose as stored at Estimator.cpp:186-190 -
  // the same pose up to f32 rounding of the two products, and that rounding reaches the motion-model guess and the velocity
  SE3f prev_relative;
  bool prev_is_keyframe = true;
  SE3f prev_frame_pose() const { return prev_is_keyframe ? prev_pose : last_keyframe_pose * prev_relative; }
  bool initialized = false;
  int n_keyframes = 0;
  bool last_was_keyframe = false, last_icp_ok = false;
  StageTimes last_times;

  explicit Pipeline(const PipelineConfig& c)
      : cfg(c), filter(c.voxel_size), map(c.map_voxel_size), ame(std::make_shared<AdaptiveMEstimator>(c.pko)), icp(c.icp, ame) {
    map.SetHierarchyFactor(3);
    map.SetPlanarityThreshold(c.surfel_planarity_threshold);
    map.SetComputeSurfels(c.icp.use_surfel_correspondence);
  }

  void transform_cloud(const std::vector<P3>& in, const SE3f& T, std::vector<P3>& out) {  // PointCloudUtils.cpp:102-125
    float M[16]; T.Matrix(M);
    out.clear(); out.reserve(in.size());
    for (const auto& p : in) { float w[3]; transform_point_4x4(M, p.x, p.y, p.z, w); out.push_back(P3{w[0], w[1], w[2]}); }
  }

  void create_keyframe() {  // Estimator.cpp:449-470
    double sensor[3] = {(double)pose.t[0], (double)pose.t[1], (double)pose.t[2]};
    map.UpdateVoxelMap(feature_world.data(), feature_world.size(), sensor, cfg.max_range * 1.2, true);
    if (!cfg.icp.use_surfel_correspondence) {  // RebuildKdTree, VoxelMap.cpp:420-438
      map.GetPointCloud(kdtree_cloud);
      if (kdtree_cloud.empty()) has_kdtree = false; else { kdtree.setInputCloud(kdtree_cloud); has_kdtree = true; }
    }
    map.GetPointCloud(local_map);
    last_keyframe_pose = pose;
    n_keyframes++;
    last_was_keyframe = true;
  }

  bool should_create_keyframe(const SE3f& cur) {  // :349-368
    if (n_keyframes == 0) return true;
    float d[3] = {cur.t[0] - last_keyframe_pose.t[0], cur.t[1] - last_keyframe_pose.t[1], cur.t[2] - last_keyframe_pose.t[2]};
    double distance = norm3<float>(d);
    SO3f rd = last_keyframe_pose.R.Inverse() * cur.R;
    float lg[3]; rd.Log(lg);
    double angle = norm3<float>(lg);
    return distance > cfg.keyframe_distance_threshold || angle > cfg.keyframe_rotation_threshold;
  }

  // process_frame (:116-233).  raw: N points, sensor frame.  Returns false if preprocessing yields nothing.
  bool process_scan(const P3* raw, size_t n) {
    using clk = std::chrono::high_resolution_clock;
    auto t0 = clk::now();
    last_was_keyframe = false; last_icp_ok = false;
    filter.filter(raw, n, feature_cloud, cfg.point_stride);
    auto t1 = clk::now();
    last_times = StageTimes();
    last_times.preprocess_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (feature_cloud.empty()) return false;
    if (!initialized) {  // initialize_first_frame
      pose = SE3f(); velocity = SE3f();
      transform_cloud(feature_cloud, pose, feature_world);
      create_keyframe();
      prev_pose = pose; prev_is_keyframe = true;
      initialized = true;
      auto t2 = clk::now();
      last_times.map_update_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
      last_times.total_ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
      return true;
    }
    SE3f guess = prev_frame_pose() * velocity;  // :154
    SE3f result = guess;
    if (!local_map.empty()) {  // estimate_motion_dual_frame :271-320
      SE3f init = SE3f::FromRt(guess.R.m, guess.t);  // SE3f(initial_guess.RotationMatrix(), ...) re-projects
      SE3f opt;
      bool ok = icp.optimize(&map, feature_cloud.data(), feature_cloud.size(), init, opt, has_kdtree ? &kdtree : nullptr, &kdtree_cloud);
      last_icp_ok = ok;
      if (ok) result = SE3f::FromRt(opt.R.m, opt.t);
    }
    auto t2 = clk::now();
    last_times.icp_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    transform_cloud(feature_cloud, result, feature_world);
    pose = result;
    velocity = prev_frame_pose().Inverse() * pose;  // :177
    const SE3f rel = last_keyframe_pose.Inverse() * pose;   // :189 (relative to the keyframe that precedes this frame)
    const bool kf = should_create_keyframe(pose);
    if (kf) create_keyframe();
    prev_pose = pose; prev_is_keyframe = kf; prev_relative = rel;
    auto t3 = clk::now();
    last_times.map_update_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
    last_times.total_ms = std::chrono::duration<double, std::milli>(t3 - t0).count();
    return true;
  }
};

}  // namespace orc
