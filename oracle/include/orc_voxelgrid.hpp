// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of util::VoxelGrid::filter
// (/root/reference/src/util/PointCloudUtils.h:462-557), the final-map downsample of Estimator::save_map_to_ply
// (/root/reference/src/processing/Estimator.cpp:1290-1298): ordered map keyed by (floor(x/leaf), floor(y/leaf), floor(z/leaf)),
// per voxel a running weighted mean in f32 fed in input order, output in key order.  Parity unpinned (no reference fixtures).
#pragma once
#include <cmath>
#include <map>
#include <tuple>
#include <vector>

namespace orc {

inline std::vector<float> voxel_grid_filter(const float* xyz, size_t n, float leaf) {
  std::vector<float> out;
  if (!xyz || n == 0 || leaf <= 0) return out;
  struct Acc { float c[3] = {0, 0, 0}; float w = 0.0f; };
  std::map<std::tuple<int, int, int>, Acc> cells;   // tuple order == VoxelKey::operator< (:531-535)
  for (size_t i = 0; i < n; ++i) {
    const float* p = xyz + 3 * i;
    const auto key = std::make_tuple((int)std::floor(p[0] / leaf), (int)std::floor(p[1] / leaf), (int)std::floor(p[2] / leaf));
    Acc& a = cells[key];
    if (a.w == 0.0f) { a.c[0] = p[0]; a.c[1] = p[1]; a.c[2] = p[2]; a.w = 1.0f; }   // :503-506
    else {                                                                          // :507-519
      const float tot = a.w + 1.0f, ro = a.w / tot, rn = 1.0f / tot;
      for (int k = 0; k < 3; ++k) a.c[k] = ro * a.c[k] + rn * p[k];
      a.w = tot;
    }
  }
  out.reserve(3 * cells.size());
  for (const auto& kv : cells) out.insert(out.end(), kv.second.c, kv.second.c + 3);
  return out;
}

}  // namespace orc
