// Stand-alone check of b2lo_dropin.h: compiles against the stub types, links libb2lo.so, and (on a GPU box) runs the
// Estimator's per-scan call sequence on a synthetic planar scene: filter -> UpdateVoxelMap -> optimize -> GetPointCloud.
#define B2LO_SHIM_STUBS
#include <cstdio>
#include <cstdlib>
#include <random>
#include "b2lo_dropin.h"

using namespace lidar_slam;

int main() {
  std::mt19937 gen(7);
  std::uniform_real_distribution<float> u(-20.0f, 20.0f);
  std::normal_distribution<float> nz(0.0f, 0.01f);
  auto scene = [&](float dx) {
    auto c = std::make_shared<util::PointCloud>();
    for (int i = 0; i < 30000; ++i) { float x = u(gen), y = u(gen); c->push_back(x - dx, y, -1.7f + nz(gen)); }                 // ground
    for (int i = 0; i < 15000; ++i) { float x = u(gen), z = 0.15f * u(gen) + 1.0f; c->push_back(x - dx, 8.0f + nz(gen), z); }   // wall
    for (int i = 0; i < 15000; ++i) { float y = u(gen), z = 0.15f * u(gen) + 1.0f; c->push_back(15.0f - dx + nz(gen), y, z); }  // end wall
    return c;
  };
  map::FastVoxelGrid grid(0.5f);
  map::VoxelMap vmap(0.5f);
  vmap.SetHierarchyFactor(3);
  vmap.SetPlanarityThreshold(0.1f);
  vmap.SetComputeSurfels(true);
  auto raw0 = scene(0.0f);
  auto ds0 = std::make_shared<util::PointCloud>();
  grid.filter(*raw0, *ds0, 2);
  std::printf("filter: %zu -> %zu voxels\n", raw0->size(), grid.getVoxelCount());
  vmap.UpdateVoxelMap(ds0, Eigen::Vector3d(0, 0, 0), 120.0, true);
  std::printf("map: L0 %zu L1 %zu surfels %zu\n", vmap.GetVoxelCount(), vmap.GetL1VoxelCount(), vmap.GetSurfelCount());
  auto raw1 = scene(0.3f);  // the sensor moved 0.3 m along x
  auto ds1 = std::make_shared<util::PointCloud>();
  grid.filter(*raw1, *ds1, 2);
  optimization::ICPConfig cfg;
  cfg.max_iterations = 10; cfg.translation_tolerance = 1e-4; cfg.rotation_tolerance = 1e-4;
  optimization::IterativeClosestPointOptimizer icp(cfg, std::make_shared<optimization::AdaptiveMEstimator>());
  auto frame = std::make_shared<database::LidarFrame>(ds1);
  SE3f init, out;
  bool ok = icp.optimize(&vmap, frame, init, out);
  Eigen::Matrix4f M = out.Matrix();
  std::printf("optimize: ok=%d iters=%zu corr=%zu t=(%.4f %.4f %.4f)\n", (int)ok, icp.get_last_stats().num_iterations,
              icp.get_last_stats().num_correspondences, M(0, 3), M(1, 3), M(2, 3));
  auto l0 = vmap.GetPointCloud();
  auto surfels = vmap.GetL1Surfels();
  Eigen::Vector3f n, c;
  bool hit = vmap.GetSurfelAtPoint(Eigen::Vector3f(1.0f, 1.0f, -1.7f), n, c);
  std::printf("export: %zu centroids, %zu surfels, lookup hit=%d n=(%.2f %.2f %.2f)\n", l0->size(), surfels.size(), (int)hit, n.x(), n.y(), n.z());
  bool pass = ok && std::fabs(M(0, 3) - 0.3f) < 0.02f && std::fabs(M(1, 3)) < 0.02f && l0->size() == vmap.GetVoxelCount() && hit && std::fabs(std::fabs(n.z()) - 1.0f) < 0.05f;
  std::printf(pass ? "DROPIN PASS\n" : "DROPIN FAIL\n");
  return pass ? 0 : 1;
}
