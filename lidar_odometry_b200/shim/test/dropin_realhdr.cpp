// Compile-and-run check of b2lo_dropin.h against the reference's OWN headers (no stubs): util::PointCloud / Point3D, SE3f,
// database::LidarFrame and optimization::AdaptiveMEstimator are the reference's real types (database/LidarFrame.h, util/MathUtils.h,
// util/PointCloudUtils.h, optimization/AdaptiveMEstimator.h, compiled where they lie under /root/reference against oracle/eigen_compat),
// the three hot-path classes are the shim's.  Built by oracle/Makefile into oracle/_ref/dropin_realhdr (so that it travels to the GPU
// box, where the reference tree does not exist) and run by tests/test_shim.py: the Estimator's per-scan call sequence on a synthetic
// planar scene, filter -> UpdateVoxelMap -> optimize(frame) -> GetPointCloud -> optimize_loop.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "b2lo_dropin.h"

using namespace lidar_slam;
using lidar_slam::util::SE3f;

int main() {
  std::mt19937 gen(7);
  std::uniform_real_distribution<float> u(-20.0f, 20.0f);
  std::normal_distribution<float> nz(0.0f, 0.01f);
  auto scene = [&](float dx) {
    auto c = std::make_shared<util::PointCloud>();
    for (int i = 0; i < 30000; ++i) { float x = u(gen), y = u(gen); c->push_back(x - dx, y, -1.7f + nz(gen)); }
    for (int i = 0; i < 15000; ++i) { float x = u(gen), z = 0.15f * u(gen) + 1.0f; c->push_back(x - dx, 8.0f + nz(gen), z); }
    for (int i = 0; i < 15000; ++i) { float y = u(gen), z = 0.15f * u(gen) + 1.0f; c->push_back(15.0f - dx + nz(gen), y, z); }
    return c;
  };
  auto make_frame = [](int id, util::PointCloudPtr raw, util::PointCloudPtr ds) {   // Estimator::preprocess_frame :561-589
    auto f = std::make_shared<database::LidarFrame>(id, 0.1 * id, raw);
    f->set_processed_cloud(ds);
    f->set_feature_cloud(ds);
    return f;
  };
  map::FastVoxelFilter grid(0.5f);
  map::VoxelMap vmap(0.5f);
  vmap.SetHierarchyFactor(3);
  vmap.SetPlanarityThreshold(0.1f);
  vmap.SetComputeSurfels(true);
  auto raw0 = scene(0.0f);
  auto ds0 = std::make_shared<util::PointCloud>();
  grid.filter(*raw0, *ds0, 2);
  vmap.UpdateVoxelMap(ds0, Eigen::Vector3d(0, 0, 0), 120.0, true);
  std::printf("map: L0 %zu L1 %zu surfels %zu\n", vmap.GetVoxelCount(), vmap.GetL1VoxelCount(), vmap.GetSurfelCount());
  auto raw1 = scene(0.3f);  // the sensor moved 0.3 m along x
  auto ds1 = std::make_shared<util::PointCloud>();
  grid.filter(*raw1, *ds1, 2);
  optimization::ICPConfig cfg;   // the shim's copy of ICP.h:55-76 (the header that defines it is the one being replaced), filled as Estimator.cpp:61-76 does
  cfg.max_iterations = 10; cfg.translation_tolerance = 1e-4; cfg.rotation_tolerance = 1e-4;
  auto pko = std::make_shared<optimization::AdaptiveMEstimator>(true, "huber", 0.1, 10.0, 100, 10.0, 3, 100, "huber");   // Estimator.cpp:49-59 order
  optimization::IterativeClosestPointOptimizer icp(cfg, pko);
  auto frame = make_frame(1, raw1, ds1);
  SE3f init, out;
  bool ok = icp.optimize(&vmap, frame, init, out);
  Eigen::Matrix4f M = out.Matrix();
  std::printf("optimize: ok=%d iters=%zu corr=%zu t=(%.4f %.4f %.4f)\n", (int)ok, (size_t)icp.get_last_stats().num_iterations,
              (size_t)icp.get_last_stats().num_correspondences, M(0, 3), M(1, 3), M(2, 3));
  ok = ok && std::fabs(M(0, 3) - 0.3f) < 0.02f && std::fabs(M(1, 3)) < 0.02f;
  Eigen::Matrix4f Mf = frame->get_pose().Matrix();   // optimize() leaves the pose on the frame as well (ICP.cpp:458-460)
  ok = ok && std::fabs(Mf(0, 3) - M(0, 3)) < 1e-6f;
  auto l0 = vmap.GetPointCloud();
  frame->set_local_map(l0);                           // Estimator.cpp:469-470
  ok = ok && l0 && l0->size() == vmap.GetVoxelCount();
  auto kf_m = make_frame(0, raw0, ds0);
  auto kf_c = make_frame(2, raw1, ds1);
  Eigen::Matrix4f Pc = Eigen::Matrix4f::Identity(); Pc(0, 3) = 0.2f;
  kf_c->set_pose(SE3f(Pc));
  SE3f rel; float ratio = 0.0f;
  bool lok = icp.optimize_loop(kf_c, kf_m, rel, ratio);
  Eigen::Matrix4f Mr = rel.Matrix();
  std::printf("optimize_loop: ok=%d inliers=%.3f rel t=(%.4f %.4f %.4f)\n", (int)lok, ratio, Mr(0, 3), Mr(1, 3), Mr(2, 3));
  ok = ok && lok && std::fabs(Mr(0, 3) - 0.1f) < 0.03f;
  std::printf(ok ? "DROPIN-REALHDR PASS\n" : "DROPIN-REALHDR FAIL\n");
  return ok ? 0 : 1;
}
