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
  // a default-constructed ICPConfig (max_iterations = 50, tolerances 1e-6: ICP.h:55-76) must register as well, not throw
  optimization::IterativeClosestPointOptimizer icp_default(optimization::ICPConfig(), std::make_shared<optimization::AdaptiveMEstimator>());
  SE3f out_d;
  bool ok_d = icp_default.optimize(&vmap, std::make_shared<database::LidarFrame>(ds1), init, out_d);
  Eigen::Matrix4f Md = out_d.Matrix();
  std::printf("optimize (default ICPConfig): ok=%d iters=%zu t=(%.4f %.4f %.4f)\n", (int)ok_d, icp_default.get_last_stats().num_iterations, Md(0, 3), Md(1, 3), Md(2, 3));
  ok = ok && ok_d && std::fabs(Md(0, 3) - 0.3f) < 0.02f && icp_default.get_last_stats().num_iterations >= 1;
  auto l0 = vmap.GetPointCloud();
  auto surfels = vmap.GetL1Surfels();
  Eigen::Vector3f n, c;
  bool hit = vmap.GetSurfelAtPoint(Eigen::Vector3f(1.0f, 1.0f, -1.7f), n, c);
  std::printf("export: %zu centroids, %zu surfels, lookup hit=%d n=(%.2f %.2f %.2f)\n", l0->size(), surfels.size(), (int)hit, n.x(), n.y(), n.z());
  // loop-closure ICP between the two "keyframes": ds1 sits 0.3 m further along x; the matched keyframe is ds0 at the origin, the
  // current one starts 0.1 m off its true pose, so the relative correction must bring back ~(+0.1, 0, 0)
  auto kf_m = std::make_shared<database::LidarFrame>(ds0);
  auto kf_c = std::make_shared<database::LidarFrame>(ds1);
  Eigen::Matrix4f Pc; Pc(0, 3) = 0.2f;
  kf_c->set_pose(SE3f(Pc));
  SE3f rel; float ratio = 0.0f;
  bool lok = icp.optimize_loop(kf_c, kf_m, rel, ratio);
  Eigen::Matrix4f Mr = rel.Matrix();
  std::printf("optimize_loop: ok=%d inliers=%.3f rel t=(%.4f %.4f %.4f)\n", (int)lok, ratio, Mr(0, 3), Mr(1, 3), Mr(2, 3));
  // final-map export (util::VoxelGrid) and KITTI-image ingest
  b2lo::VoxelGrid vg;
  vg.setLeafSize(1.0f);
  vg.setInputCloud(ds0);
  util::PointCloud coarse;
  vg.filter(coarse);
  bool sorted = coarse.size() > 100 && coarse.size() < ds0->size();
  for (size_t i = 1; i < coarse.size() && sorted; ++i) sorted = std::floor(coarse[i - 1].x) <= std::floor(coarse[i].x);
  std::vector<float> bin;
  for (size_t i = 0; i < raw0->size(); ++i) { bin.push_back((*raw0)[i].x); bin.push_back((*raw0)[i].y); bin.push_back((*raw0)[i].z); bin.push_back(0.5f); }
  util::PointCloud from_image;
  bool iok = b2lo::filter_scan_file_image(bin.data(), bin.size() * sizeof(float), false, 0.5f, 2, from_image);
  bool same = iok && from_image.size() == ds0->size();
  for (size_t i = 0; i < from_image.size() && same; ++i) same = from_image[i].x == (*ds0)[i].x && from_image[i].y == (*ds0)[i].y && from_image[i].z == (*ds0)[i].z;
  std::printf("voxel grid: %zu -> %zu (x-sorted %d); .bin image ingest identical to filter(): %d\n", ds0->size(), coarse.size(), (int)sorted, (int)same);
  bool pass = lok && ratio > 0.9f && std::fabs(Mr(0, 3) - 0.1f) < 0.03f && std::fabs(Mr(1, 3)) < 0.03f && sorted && same &&
              ok && std::fabs(M(0, 3) - 0.3f) < 0.02f && std::fabs(M(1, 3)) < 0.02f && l0->size() == vmap.GetVoxelCount() && hit && std::fabs(std::fabs(n.z()) - 1.0f) < 0.05f;
  std::printf(pass ? "DROPIN PASS\n" : "DROPIN FAIL\n");
  return pass ? 0 : 1;
}
