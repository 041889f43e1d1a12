// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_voxel.hpp — CPU restatement of
//   map::FastVoxelFilter            /root/reference/src/database/VoxelMap.h:53-140
//   map::VoxelKey / VoxelKeyHash    /root/reference/src/database/VoxelMap.h:152-183
//   map::VoxelMap                   /root/reference/src/database/VoxelMap.h:188-332, VoxelMap.cpp:20-438
// Arithmetic is IEEE f32 without FMA contraction (the reference builds for baseline x86-64,
// CMakeLists.txt:19-21); compile this file with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <array>
#include "orc_dense_map.hpp"
#include "orc_eigen.hpp"

namespace orc {

struct P3 { float x, y, z; };

// ---- FastVoxelFilter --------------------------------------------------------------------------
static inline uint64_t expand_bits21(uint64_t v) {  // VoxelMap.h:114-122
  v = v & 0x1FFFFF;
  v = (v | (v << 32)) & 0x1F00000000FFFFULL;
  v = (v | (v << 16)) & 0x1F0000FF0000FFULL;
  v = (v | (v << 8)) & 0x100F00F00F00F00FULL;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ULL;
  v = (v | (v << 2)) & 0x1249249249249249ULL;
  return v;
}
static inline uint64_t filter_morton_key(float x, float y, float z, float inv_voxel) {  // VoxelMap.h:124-135
  const int64_t OFFSET = (1 << 20);
  int64_t ix = (int64_t)std::floor(x * inv_voxel) + OFFSET;
  int64_t iy = (int64_t)std::floor(y * inv_voxel) + OFFSET;
  int64_t iz = (int64_t)std::floor(z * inv_voxel) + OFFSET;
  ix = std::max<int64_t>(0, std::min<int64_t>(ix, (1 << 21) - 1));
  iy = std::max<int64_t>(0, std::min<int64_t>(iy, (1 << 21) - 1));
  iz = std::max<int64_t>(0, std::min<int64_t>(iz, (1 << 21) - 1));
  return expand_bits21((uint64_t)ix) | (expand_bits21((uint64_t)iy) << 1) | (expand_bits21((uint64_t)iz) << 2);
}
struct U64Hash { uint64_t operator()(uint64_t k) const { return k; } };

class FastVoxelFilter {
 public:
  explicit FastVoxelFilter(float voxel_size = 0.5f) : voxel_(voxel_size), inv_(1.0f / voxel_size) {}
  void setVoxelSize(float v) { voxel_ = v; inv_ = 1.0f / v; }
  float getVoxelSize() const { return voxel_; }
  size_t getVoxelCount() const { return map_.size(); }
  // VoxelMap.h:73-104.  Optional out_keys receives the Morton key of each output voxel.
  void filter(const P3* in, size_t n, std::vector<P3>& out, int stride = 1, std::vector<uint64_t>* out_keys = nullptr) {
    out.clear();
    if (out_keys) out_keys->clear();
    if (n == 0 || stride < 1) return;
    map_.clear();
    map_.reserve(n / ((size_t)stride * 8));
    for (size_t i = 0; i < n; i += (size_t)stride) {
      const P3& pt = in[i];
      if (!std::isfinite(pt.x) || !std::isfinite(pt.y) || !std::isfinite(pt.z)) continue;
      uint64_t key = filter_morton_key(pt.x, pt.y, pt.z, inv_);
      Acc& v = map_[key];
      v.sx += pt.x; v.sy += pt.y; v.sz += pt.z; v.count++;
    }
    out.reserve(map_.size());
    for (const auto& kv : map_.values) {
      float ic = 1.0f / (float)kv.second.count;
      out.push_back(P3{kv.second.sx * ic, kv.second.sy * ic, kv.second.sz * ic});
      if (out_keys) out_keys->push_back(kv.first);
    }
  }
 private:
  struct Acc { float sx = 0, sy = 0, sz = 0; uint32_t count = 0; };
  float voxel_, inv_;
  DenseMap<uint64_t, Acc, U64Hash> map_;
};

// ---- VoxelKey ---------------------------------------------------------------------------------
struct VoxelKey {
  int x = 0, y = 0, z = 0;
  bool operator==(const VoxelKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
static inline uint64_t voxel_key_morton(const VoxelKey& k) {  // VoxelMap.h:166-183 (wraps, does not clamp)
  auto ex = [](int32_t v) { return expand_bits21((uint64_t)(int64_t)(v + (1 << 20)) & 0x1fffff); };
  return ex(k.x) | (ex(k.y) << 1) | (ex(k.z) << 2);
}
struct VoxelKeyHash { uint64_t operator()(const VoxelKey& k) const { return voxel_key_morton(k); } };

// ---- VoxelMap ---------------------------------------------------------------------------------
class VoxelMap {
 public:
  struct L0 { float c[3] = {0, 0, 0}; int hit_count = 1; int point_count = 0; };
  struct L1 {
    DenseSet<VoxelKey, VoxelKeyHash> children;
    bool has_surfel = false;
    float normal[3] = {0, 0, 0};
    float centroid[3] = {0, 0, 0};
    float planarity = 1.0f;
    int last_child_count = 0;
  };

  explicit VoxelMap(float voxel_size = 0.5f) : voxel_(voxel_size) {}
  // VoxelMap.cpp:27-48
  bool SetVoxelSize(float s) { if (s <= 0.0f) return false; if (std::abs(voxel_ - s) > 1e-6f) { voxel_ = s; Clear(); } return true; }
  bool SetHierarchyFactor(int f) { if (f <= 0 || f % 2 == 0) return false; if (factor_ != f) { factor_ = f; Clear(); } return true; }
  void SetPlanarityThreshold(float t) { planarity_thr_ = t; }
  void SetComputeSurfels(bool c) { compute_surfels_ = c; }
  void SetInitHitCount(int c) { init_hit_ = c; }
  float GetVoxelSize() const { return voxel_; }
  int GetHierarchyFactor() const { return factor_; }
  size_t GetVoxelCount() const { return l0_.size(); }
  size_t GetL1VoxelCount() const { return l1_.size(); }
  bool empty() const { return l0_.empty(); }
  size_t GetSurfelCount() const { size_t c = 0; for (auto& kv : l1_.values) if (kv.second.has_surfel) c++; return c; }
  void Clear() { l0_.clear(); l1_.clear(); }

  VoxelKey PointToVoxelKey(const float* p, int level) const {  // VoxelMap.cpp:50-58
    float scale = voxel_;
    if (level == 1) scale *= (float)factor_;
    return VoxelKey{(int)std::floor(p[0] / scale), (int)std::floor(p[1] / scale), (int)std::floor(p[2] / scale)};
  }
  VoxelKey GetParentKey(const VoxelKey& k) const {  // VoxelMap.cpp:60-67
    int f = factor_;
    return VoxelKey{k.x >= 0 ? k.x / f : (k.x - (f - 1)) / f, k.y >= 0 ? k.y / f : (k.y - (f - 1)) / f,
                    k.z >= 0 ? k.z / f : (k.z - (f - 1)) / f};
  }

  // VoxelMap.cpp:128-262
  void UpdateVoxelMap(const P3* cloud, size_t n, const double sensor_position[3], double max_distance, bool is_keyframe) {
    if (!cloud || n == 0) return;
    if (!is_keyframe) return;
    float sensor[3] = {(float)sensor_position[0], (float)sensor_position[1], (float)sensor_position[2]};
    float radius_sq = (float)(max_distance * max_distance);
    std::vector<VoxelKey> to_remove;
    for (const auto& kv : l0_.values) {
      float d[3] = {kv.second.c[0] - sensor[0], kv.second.c[1] - sensor[1], kv.second.c[2] - sensor[2]};
      float dist_sq = sqnorm3<float>(d);
      if (dist_sq > radius_sq) to_remove.push_back(kv.first);
    }
    for (const auto& k : to_remove) { UnregisterFromParent(k); l0_.erase(k); }
    std::vector<VoxelKey> l1_remove;
    for (const auto& kv : l1_.values) if (kv.second.children.empty()) l1_remove.push_back(kv.first);
    for (const auto& k : l1_remove) l1_.erase(k);

    DenseSet<VoxelKey, VoxelKeyHash> affected;
    for (size_t i = 0; i < n; ++i) {
      float p[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
      AddPoint(p);
      affected.insert(PointToVoxelKey(p, 1));
    }
    if (!compute_surfels_) return;
    const int MIN_OCC = 5;
    for (size_t ai = 0; ai < affected.size(); ++ai) {
      const VoxelKey key_L1 = affected.at(ai);
      int64_t it = l1_.find(key_L1);
      if (it < 0) continue;
      L1& node = l1_.values[(size_t)it].second;
      int cur = (int)node.children.size();
      if (cur < MIN_OCC) { node.has_surfel = false; continue; }
      if (node.has_surfel && node.last_child_count == cur) continue;
      std::vector<std::array<float, 3>> cents;
      cents.reserve(node.children.size());
      for (size_t ci = 0; ci < node.children.size(); ++ci) {
        int64_t i0 = l0_.find(node.children.at(ci));
        if (i0 >= 0) { const L0& v = l0_.values[(size_t)i0].second; cents.push_back({v.c[0], v.c[1], v.c[2]}); }
      }
      if (cents.size() < 3) { node.has_surfel = false; continue; }
      float mu[3], nrm[3], plan;
      FitPlane(cents, mu, nrm, plan);
      if (plan > planarity_thr_) {
        node.has_surfel = false;
        for (size_t ci = 0; ci < node.children.size(); ++ci) l0_.erase(node.children.at(ci));
        l1_.erase_at((size_t)it);
        continue;
      }
      node.has_surfel = true;
      for (int a = 0; a < 3; ++a) { node.normal[a] = nrm[a]; node.centroid[a] = mu[a]; }
      node.planarity = plan;
      node.last_child_count = cur;
    }
  }

  // VoxelMap.cpp:264-302 (+ RecomputeAllSurfels :304-366)
  void ApplyTransformAndRehash(const float* T /*row-major 4x4*/) {
    float R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    float t[3] = {T[3], T[7], T[11]};
    std::vector<std::pair<VoxelKey, L0>> transformed;
    transformed.reserve(l0_.size());
    for (auto& kv : l0_.values) {
      L0 nn = kv.second;
      float rc[3];
      mat3_mul_vec<float>(R, kv.second.c, rc);
      for (int a = 0; a < 3; ++a) nn.c[a] = rc[a] + t[a];
      transformed.emplace_back(PointToVoxelKey(nn.c, 0), nn);
    }
    l0_.clear(); l1_.clear();
    for (auto& kn : transformed) {
      L0& ex = l0_[kn.first];
      if (ex.point_count == 0) ex = kn.second;
      else {
        float n1 = (float)ex.point_count, n2 = (float)kn.second.point_count;
        for (int a = 0; a < 3; ++a) ex.c[a] = (ex.c[a] * n1 + kn.second.c[a] * n2) / (n1 + n2);
        ex.point_count += kn.second.point_count;
      }
      RegisterToParent(kn.first);
    }
    RecomputeAllSurfels();
  }

  // VoxelMap.cpp:368-386
  bool GetSurfelAtPoint(const float* p, float* normal, float* centroid) const {
    VoxelKey k = PointToVoxelKey(p, 1);
    int64_t it = l1_.find(k);
    if (it < 0) return false;
    const L1& node = l1_.values[(size_t)it].second;
    if (!node.has_surfel) return false;
    for (int a = 0; a < 3; ++a) { normal[a] = node.normal[a]; centroid[a] = node.centroid[a]; }
    return true;
  }
  // VoxelMap.cpp:388-403
  void GetPointCloud(std::vector<P3>& out) const {
    out.clear(); out.reserve(l0_.size());
    for (const auto& kv : l0_.values) out.push_back(P3{kv.second.c[0], kv.second.c[1], kv.second.c[2]});
  }

  const DenseMap<VoxelKey, L0, VoxelKeyHash>& l0() const { return l0_; }
  const DenseMap<VoxelKey, L1, VoxelKeyHash>& l1() const { return l1_; }

  // plane fit shared by UpdateVoxelMap (:223-242) and RecomputeAllSurfels (:332-353)
  static void FitPlane(const std::vector<std::array<float, 3>>& cents, float* mu, float* nrm, float& planarity) {
    float c[3] = {0, 0, 0};
    for (const auto& p : cents) { c[0] += p[0]; c[1] += p[1]; c[2] += p[2]; }
    float nf = (float)cents.size();
    c[0] /= nf; c[1] /= nf; c[2] /= nf;
    float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (const auto& p : cents) {
      float d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[i * 3 + j] += d[i] * d[j];
    }
    for (int i = 0; i < 9; ++i) cov[i] /= nf;
    float U[9], S[3], V[9];
    jacobi_svd3<float>(cov, U, S, V);
    nrm[0] = U[0 * 3 + 2]; nrm[1] = U[1 * 3 + 2]; nrm[2] = U[2 * 3 + 2];
    planarity = S[2] / (S[0] + 1e-6f);
    mu[0] = c[0]; mu[1] = c[1]; mu[2] = c[2];
  }

 private:
  void RegisterToParent(const VoxelKey& k0) { l1_[GetParentKey(k0)].children.insert(k0); }  // :77-80
  void UnregisterFromParent(const VoxelKey& k0) {  // :82-97
    VoxelKey parent = GetParentKey(k0);
    int64_t it = l1_.find(parent);
    if (it < 0) return;
    L1& node = l1_.values[(size_t)it].second;
    node.children.erase(k0);
    if (node.children.size() < 5) node.has_surfel = false;
    if (node.children.empty()) l1_.erase_at((size_t)it);
  }
  void AddPoint(const float* p) {  // :99-120
    VoxelKey key = PointToVoxelKey(p, 0);
    auto r = l0_.try_emplace(key);
    L0& v = l0_.values[r.first].second;
    int n = v.point_count;
    if (n == 0) { v.c[0] = p[0]; v.c[1] = p[1]; v.c[2] = p[2]; v.hit_count = init_hit_; v.point_count = 1; }
    else {
      float fn = (float)n, fn1 = (float)(n + 1);
      for (int a = 0; a < 3; ++a) v.c[a] = (v.c[a] * fn + p[a]) / fn1;
      v.point_count++;
    }
    if (r.second) RegisterToParent(key);
  }
  void RecomputeAllSurfels() {  // :304-366
    const int MIN_OCC = 5;
    for (auto& kv : l1_.values) {
      L1& node = kv.second;
      int cur = (int)node.children.size();
      if (cur < MIN_OCC) { node.has_surfel = false; continue; }
      std::vector<std::array<float, 3>> cents;
      for (size_t ci = 0; ci < node.children.size(); ++ci) {
        int64_t i0 = l0_.find(node.children.at(ci));
        if (i0 >= 0) { const L0& v = l0_.values[(size_t)i0].second; cents.push_back({v.c[0], v.c[1], v.c[2]}); }
      }
      if (cents.size() < (size_t)MIN_OCC) { node.has_surfel = false; continue; }
      float mu[3], nrm[3], plan;
      FitPlane(cents, mu, nrm, plan);
      if (plan > planarity_thr_) { node.has_surfel = false; continue; }
      node.has_surfel = true;
      for (int a = 0; a < 3; ++a) { node.normal[a] = nrm[a]; node.centroid[a] = mu[a]; }
      node.planarity = plan;
      node.last_child_count = cur;
    }
  }

  float voxel_;
  int factor_ = 3;
  int init_hit_ = 1;
  float planarity_thr_ = 0.1f;
  bool compute_surfels_ = true;
  DenseMap<VoxelKey, L0, VoxelKeyHash> l0_;
  DenseMap<VoxelKey, L1, VoxelKeyHash> l1_;
};

}  // namespace orc
