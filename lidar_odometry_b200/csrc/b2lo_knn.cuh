// b2lo_knn.cuh — K3: exact 5-NN over the L0 centroids + per-query plane fit ("KDTree" correspondence mode).
//
// Replaces IterativeClosestPointOptimizer::find_correspondences_kdtree
// (/root/reference/src/optimization/IterativeClosestPointOptimizer.cpp:647-767), is_collinear (:785-792),
// util::KdTree::nearestKSearch (src/util/PointCloudUtils.h:398-423) and VoxelMap::RebuildKdTree
// (src/database/VoxelMap.cpp:420-438).  nanoflann returns the exact k nearest by f32 squared L2
// ((dx^2 + dy^2) + dz^2, thirdparty/nanoflann/nanoflann.hpp:638-649) in ascending order, so any exact
// search yields the same indices except on exact f32 distance ties (ties here break towards the smaller
// GetPointCloud index; nanoflann keeps the first visited).
//
// No separate index is built: every L0 centroid lies inside the cell of its own Z-order key, so the L0 hash
// IS a uniform grid.  A query walks Chebyshev shells of cells around its own cell and stops as soon as the
// 5th-best distance is smaller than the distance to the boundary of the visited cube (one warp per query, one
// lane per cell).  Queries still unresolved after KNN_MAX_RING shells (farther than ~2 cells from five map points)
// are finished by a warp-per-query exact scan of the dense L0 centroid stream (coalesced float4, 16 B / voxel).
#pragma once
#include "b2lo_dev.cuh"

namespace b2 {

constexpr int KNN_K = 5;
constexpr int KNN_MAX_RING = 2;   // shells of L0 cells probed before a query falls back to the exact full scan (measured: more shells
                                  // do not pay - queries that miss two shells are usually far from the map altogether)

struct Top5 {
  float d[KNN_K]; int id[KNN_K]; int n;
  __device__ __forceinline__ void init() { n = 0; for (int i = 0; i < KNN_K; ++i) { d[i] = 3.402823466e+38f; id[i] = -1; } }
  __device__ __forceinline__ void push(float dd, int idx) {
    if (n == KNN_K && !(dd < d[KNN_K - 1] || (dd == d[KNN_K - 1] && idx < id[KNN_K - 1]))) return;
    int j = n < KNN_K ? n : KNN_K - 1;
    while (j > 0 && (d[j - 1] > dd || (d[j - 1] == dd && id[j - 1] > idx))) { d[j] = d[j - 1]; id[j] = id[j - 1]; --j; }
    d[j] = dd; id[j] = idx;
    if (n < KNN_K) ++n;
  }
};

__device__ __forceinline__ float knn_dist2(const float* w, float cx, float cy, float cz) {
  float dx = w[0] - cx, dy = w[1] - cy, dz = w[2] - cz;
  return (dx * dx + dy * dy) + dz * dz;
}

// merge the per-lane sorted candidate lists into the warp-wide top 5 (ascending (d2, index)); result valid in every lane
__device__ __forceinline__ void knn_warp_merge(const Top5& mine, Top5& top) {
  top.init();
  int head = 0;
  for (int k = 0; k < KNN_K; ++k) {
    float d = head < mine.n ? mine.d[head] : 3.402823466e+38f;
    int id = head < mine.n ? mine.id[head] : 0x7fffffff;
    float bd = d; int bi = id;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, bd, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (bi == 0x7fffffff) break;
    if (bi == id && head < mine.n) ++head;
    top.d[top.n] = bd; top.id[top.n] = bi; ++top.n;
  }
}

// warp-cooperative exact scan of all n0 centroids; result valid in every lane
__device__ __forceinline__ void knn_brute_warp(const MapDev& M, int n0, const float* w, Top5& top) {
  const int lane = threadIdx.x & 31;
  Top5 mine;
  mine.init();
  // 8 independent 16 B loads in flight per lane (the loop is latency-bound otherwise), candidates rejected against the current
  // 5th-best before the sorted insert
  for (int pos0 = lane; pos0 < n0; pos0 += 32 * 8) {
    float4 c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int pos = pos0 + 32 * u; c[u] = (pos < n0) ? M.l0_cent[pos] : make_float4(3e18f, 3e18f, 3e18f, 0.0f); }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int pos = pos0 + 32 * u;
      const float dd = knn_dist2(w, c[u].x, c[u].y, c[u].z);
      if (pos < n0 && (mine.n < KNN_K || dd <= mine.d[KNN_K - 1])) mine.push(dd, pos);
    }
  }
  knn_warp_merge(mine, top);
}

__device__ __forceinline__ bool knn_exact_within(const MapDev& M, const float* w, int kx, int ky, int kz, int r, const Top5& top, float slop) {
  if (top.n < KNN_K) return false;
  const float lo = (float)r * M.voxel, hi = (float)(r + 1) * M.voxel;
  float m = 3.402823466e+38f;
  const int k3[3] = {kx, ky, kz};
  for (int a = 0; a < 3; ++a) {
    float base = (float)k3[a] * M.voxel;
    m = fminf(m, fminf((w[a] - base) + lo, (base - w[a]) + hi));
  }
  m -= slop;
  return m > 0.0f && top.d[KNN_K - 1] < m * m;
}
// one WARP per query: lane c probes cell c of the 3x3x3 cube around the query's cell (27 independent probes in flight), the
// candidates are merged with shuffles; while the 5th distance does not clear the visited cube's faces, the next shell of cells
// follows (98, 218, 386 cells, one lane per cell).  Returns true when the top-5 is provably exact.  Result in every lane.
__device__ __forceinline__ bool knn_rings_warp(const MapDev& M, const float* w, Top5& top) {
  const int lane = threadIdx.x & 31;
  top.init();
  int kx = voxel_coord(w[0], M.voxel), ky = voxel_coord(w[1], M.voxel), kz = voxel_coord(w[2], M.voxel);
  if (!key_in_range(kx, ky, kz)) return false;
  const float slop = 1e-3f * M.voxel + 4e-6f * fmaxf(fabsf(w[0]), fmaxf(fabsf(w[1]), fabsf(w[2])));
  Top5 mine;
  mine.init();
  if (lane < 27) {
    int x = kx + lane % 3 - 1, y = ky + (lane / 3) % 3 - 1, z = kz + lane / 9 - 1;
    if (key_in_range(x, y, z)) {
      int s0 = l0_find(M, key_pack(x, y, z));
      if (s0 >= 0) { int pos = (int)M.l0_tab[s0].pos; float4 c = M.l0_cent[pos]; mine.push(knn_dist2(w, c.x, c.y, c.z), pos); }
    }
  }
  knn_warp_merge(mine, top);
  if (knn_exact_within(M, w, kx, ky, kz, 1, top, slop)) return true;
  for (int r = 2; r <= KNN_MAX_RING; ++r) {
    const int side = 2 * r + 1, cells = side * side * side;
    for (int c = lane; c < cells; c += 32) {
      int dx = c % side - r, dy = (c / side) % side - r, dz = c / (side * side) - r;
      if (dx > -r && dx < r && dy > -r && dy < r && dz > -r && dz < r) continue;   // inner cube already visited
      int x = kx + dx, y = ky + dy, z = kz + dz;
      if (!key_in_range(x, y, z)) continue;
      int s0 = l0_find(M, key_pack(x, y, z));
      if (s0 < 0) continue;
      int pos = (int)M.l0_tab[s0].pos;
      float4 cc = M.l0_cent[pos];
      mine.push(knn_dist2(w, cc.x, cc.y, cc.z), pos);
    }
    knn_warp_merge(mine, top);
    if (knn_exact_within(M, w, kx, ky, kz, r, top, slop)) return true;
  }
  return false;
}

// is_collinear (ICP.cpp:785-792), f64, Eigen normalized() / cross / norm
__device__ __forceinline__ bool knn_collinear(const double* p1, const double* p2, const double* p3, double thr) {
  double a[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, b[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
  double na = add3(a[0] * a[0], a[1] * a[1], a[2] * a[2]), nb = add3(b[0] * b[0], b[1] * b[1], b[2] * b[2]);
  if (na > 0.0) { double s = sqrt(na); a[0] /= s; a[1] /= s; a[2] /= s; }
  if (nb > 0.0) { double s = sqrt(nb); b[0] /= s; b[1] /= s; b[2] /= s; }
  double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  return sqrt(add3(c[0] * c[0], c[1] * c[1], c[2] * c[2])) < thr;
}

// smallest right-singular vector of a 5x3 f64 matrix: one-sided (Hestenes) Jacobi on the columns — the
// engine's stated replacement for JacobiSVD<MatrixXd>(A, ComputeFullV) (ICP.cpp:745); sign irrelevant downstream
__device__ __forceinline__ void knn_plane_normal(const double* A /*5x3 row-major*/, double* normal) {
  double B[15], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 15; ++i) B[i] = A[i];
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 5; ++i) { alpha += B[i * 3 + p] * B[i * 3 + p]; beta += B[i * 3 + q] * B[i * 3 + q]; gamma += B[i * 3 + p] * B[i * 3 + q]; }
        if (gamma == 0.0 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 5; ++i) { double bp = B[i * 3 + p], bq = B[i * 3 + q]; B[i * 3 + p] = c * bp - s * bq; B[i * 3 + q] = s * bp + c * bq; }
        for (int i = 0; i < 3; ++i) { double vp = V[i * 3 + p], vq = V[i * 3 + q]; V[i * 3 + p] = c * vp - s * vq; V[i * 3 + q] = s * vp + c * vq; }
      }
    if (!rotated) break;
  }
  double nrm[3] = {0, 0, 0};
  for (int j = 0; j < 3; ++j) for (int i = 0; i < 5; ++i) nrm[j] += B[i * 3 + j] * B[i * 3 + j];
  int best = 0;
  for (int j = 1; j < 3; ++j) if (nrm[j] < nrm[best]) best = j;
  for (int i = 0; i < 3; ++i) normal[i] = V[i * 3 + best];
}

// plane through the 5 neighbours; returns state 0 (<5 found or collinear), 1 (gated out), 2 (accepted)
__device__ __forceinline__ int knn_fit(const MapDev& M, const Top5& top, const float* w, double max_dist, float* n_out, float* c_out, double* res) {
  if (top.n < KNN_K) return 0;
  double sel[KNN_K][3];
  for (int k = 0; k < KNN_K; ++k) { float4 c = M.l0_cent[top.id[k]]; sel[k][0] = (double)c.x; sel[k][1] = (double)c.y; sel[k][2] = (double)c.z; }
  if (knn_collinear(sel[0], sel[1], sel[2], 0.5)) return 0;
  double cen[3] = {0, 0, 0};
  for (int k = 0; k < KNN_K; ++k) { cen[0] += sel[k][0]; cen[1] += sel[k][1]; cen[2] += sel[k][2]; }
  cen[0] /= (double)KNN_K; cen[1] /= (double)KNN_K; cen[2] /= (double)KNN_K;
  double A[15], nrm[3];
  for (int k = 0; k < KNN_K; ++k) for (int a = 0; a < 3; ++a) A[k * 3 + a] = sel[k][a] - cen[a];
  knn_plane_normal(A, nrm);
  double plane_d = -add3(nrm[0] * cen[0], nrm[1] * cen[1], nrm[2] * cen[2]);
  double dist = fabs(add3(nrm[0] * (double)w[0], nrm[1] * (double)w[1], nrm[2] * (double)w[2]) + plane_d);
  *res = dist;
  for (int a = 0; a < 3; ++a) { n_out[a] = (float)nrm[a]; c_out[a] = (float)cen[a]; }
  return dist > max_dist ? 1 : 2;
}

}  // namespace b2
