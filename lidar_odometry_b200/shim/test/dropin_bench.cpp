// dropin_bench.cpp — wall time of the CLASS-BY-CLASS drop-in: what an unmodified processing::Estimator pays per scan when
// lidar_slam::map::FastVoxelFilter, map::VoxelMap and optimization::IterativeClosestPointOptimizer are the classes of b2lo_dropin.h.
// The loop below is Estimator::process_frame's call sequence on the hot path (src/processing/Estimator.cpp:116-233, 271-320, 349-368,
// 449-470) with the reference's data flow: pageable std::vector clouds in, host clouds out, every hand-over through the host -
//   preprocess_frame      filter(raw, ds, stride)                 H2D of the sampled points, D2H of the feature cloud      (:570, :581-582)
//   estimate_motion       optimize(&map, frame, guess, out)       H2D of the feature cloud again, ICP, D2H of the state     (:297-302)
//   feature cloud -> world on the HOST (transform_point_cloud)                                                             (:165-169)
//   keyframe decision on the host                                                                                          (:349-368)
//   create_keyframe       UpdateVoxelMap(world cloud, ...)        H2D of the world cloud, K6                                (:457)
//                         GetPointCloud()                         D2H of EVERY L0 centroid, each keyframe                   (:469-470)
// Input: a file of concatenated scans written by bench.py (u32 count, then count x 4 f32 per scan).  Prints one JSON line.
#define B2LO_SHIM_STUBS
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "b2lo_dropin.h"

using namespace lidar_slam;

static void mul(const float* A, const float* B, float* C) { b2lo_se3_mul(A, B, C); }
static Eigen::Matrix4f to_eigen(const float* T) { Eigen::Matrix4f M; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M(i, j) = T[i * 4 + j]; return M; }
static void from_eigen(const Eigen::Matrix4f& M, float* T) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T[i * 4 + j] = M(i, j); }

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: dropin_bench <scans.bin> <warmup>\n"); return 2; }
  const int warmup = std::atoi(argv[2]);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  std::vector<util::PointCloudPtr> scans;
  for (;;) {
    unsigned n = 0;
    if (std::fread(&n, 4, 1, f) != 1) break;
    std::vector<float> buf((size_t)n * 4);
    if (std::fread(buf.data(), 16, n, f) != n) break;
    auto c = std::make_shared<util::PointCloud>();   // load_kitti_binary keeps x, y, z (PointCloudUtils.cpp:40-58)
    c->reserve(n);
    for (unsigned i = 0; i < n; ++i) c->push_back(buf[4 * i], buf[4 * i + 1], buf[4 * i + 2]);
    scans.push_back(c);
  }
  std::fclose(f);
  // Estimator's objects with config/kitti.yaml (Estimator.cpp:49-81)
  map::FastVoxelGrid grid(0.5f);
  map::VoxelMap vmap(0.5f);
  vmap.SetHierarchyFactor(3); vmap.SetPlanarityThreshold(0.1f); vmap.SetComputeSurfels(true);
  optimization::ICPConfig cfg;
  cfg.max_iterations = 4; cfg.translation_tolerance = 0.005; cfg.rotation_tolerance = 0.005; cfg.max_correspondence_distance = 1.0;
  optimization::IterativeClosestPointOptimizer icp(cfg, std::make_shared<optimization::AdaptiveMEstimator>());
  const int stride = 8;
  const double max_range = 100.0, kf_dist = 1.0, kf_rot = 0.3;
  float pose[16], prev[16], vel[16], last_kf[16], I[16];
  for (int i = 0; i < 16; ++i) I[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  std::memcpy(pose, I, sizeof I); std::memcpy(prev, I, sizeof I); std::memcpy(vel, I, sizeof I); std::memcpy(last_kf, I, sizeof I);
  bool initialized = false;
  int keyframes = 0;
  double t_total = 0, t_pre = 0, t_icp = 0, t_map = 0, t_export = 0, t_xf = 0;
  size_t exported = 0, feats = 0;
  int timed = 0;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  util::PointCloudPtr local_map;
  for (size_t k = 0; k < scans.size(); ++k) {
    const bool count = (int)k >= warmup;
    const double t0 = now();
    auto ds = std::make_shared<util::PointCloud>();
    grid.filter(*scans[k], *ds, stride);                                     // preprocess_frame
    const double t1 = now();
    bool make_kf = false;
    if (!initialized) { initialized = true; make_kf = true; }
    else {
      float guess[16], out[16];
      mul(prev, vel, guess);                                                 // :154
      auto frame = std::make_shared<database::LidarFrame>(ds);
      SE3f o;
      bool ok = local_map && !local_map->empty() && icp.optimize(&vmap, frame, SE3f(to_eigen(guess)), o);   // :279-302
      if (ok) from_eigen(o.Matrix(), out); else std::memcpy(out, guess, sizeof out);
      b2lo_se3_from_rt(out, pose);                                           // SE3f(R, t) re-projection (:300-302)
      float pinv[16];
      b2lo_se3_inv(prev, pinv); mul(pinv, pose, vel);                        // :177
      float kinv[16], rel[16], w[3];
      b2lo_se3_inv(last_kf, kinv); mul(kinv, pose, rel); b2lo_so3_log(rel, w);
      const double d = std::sqrt((double)(pose[3] - last_kf[3]) * (pose[3] - last_kf[3]) + (double)(pose[7] - last_kf[7]) * (pose[7] - last_kf[7]) +
                                 (double)(pose[11] - last_kf[11]) * (pose[11] - last_kf[11]));
      make_kf = d > kf_dist || std::sqrt((double)w[0] * w[0] + (double)w[1] * w[1] + (double)w[2] * w[2]) > kf_rot;   // :349-368
    }
    const double t2 = now();
    double t3 = t2, t4 = t2;
    if (make_kf) {
      auto world = std::make_shared<util::PointCloud>();                     // transform_point_cloud on the host (:165-169)
      world->reserve(ds->size());
      for (size_t i = 0; i < ds->size(); ++i) {
        const util::Point3D& p = (*ds)[i];
        world->push_back(((pose[0] * p.x + pose[1] * p.y) + pose[2] * p.z) + pose[3], ((pose[4] * p.x + pose[5] * p.y) + pose[6] * p.z) + pose[7],
                         ((pose[8] * p.x + pose[9] * p.y) + pose[10] * p.z) + pose[11]);
      }
      const double t2a = now();
      if (count) t_xf += t2a - t2;
      vmap.UpdateVoxelMap(world, Eigen::Vector3d(pose[3], pose[7], pose[11]), 1.2 * max_range, true);   // :455-457
      t3 = now();
      local_map = vmap.GetPointCloud();                                      // :469-470: every L0 centroid to the host
      t4 = now();
      std::memcpy(last_kf, pose, sizeof pose);
      ++keyframes;
      if (count) exported += local_map->size();
    }
    std::memcpy(prev, pose, sizeof pose);
    if (count) { t_total += t4 - t0; t_pre += t1 - t0; t_icp += t2 - t1; t_map += t3 - t2; t_export += t4 - t3; feats += ds->size(); ++timed; }
  }
  std::printf("{\"scans\": %d, \"ms_per_scan\": %.5f, \"scans_per_s\": %.2f, \"stage_ms_per_scan\": {\"filter\": %.5f, \"optimize\": %.5f, \"host_transform\": %.5f, \"update_voxel_map\": %.5f, "
              "\"get_point_cloud\": %.5f}, \"keyframes\": %d, \"features_per_scan\": %.1f, \"l0_exported_per_keyframe\": %.1f, \"final_pose_t\": [%.4f, %.4f, %.4f]}\n",
              timed, t_total / timed, 1e3 * timed / t_total, t_pre / timed, t_icp / timed, t_xf / timed, (t_map - t_xf) / timed, t_export / timed, keyframes, (double)feats / timed,
              keyframes ? (double)exported / keyframes : 0.0, pose[3], pose[7], pose[11]);
  return 0;
}
