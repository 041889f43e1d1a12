// b2lo_icp.cu — K2/K4/K5: scan-to-map ICP with the Gauss-Newton loop resident on the device.
//
// Replaces IterativeClosestPointOptimizer::optimize (/root/reference/src/optimization/IterativeClosestPointOptimizer.cpp:255-463)
// with find_correspondences (:587-645), VoxelMap::GetSurfelAtPoint (src/database/VoxelMap.cpp:368-386),
// the residual normalisation (:304-316), AdaptiveMEstimator::calculate_scale_factor
// (src/optimization/AdaptiveMEstimator.cpp:243-291, 294-485, 710-787), the normal equations (:345-410)
// and the solve + SE(3) right update (:417-448).
//
// One Gauss-Newton iteration = four launches on the context stream, no host round trip:
//   k_icp_corr   : per query: T*p (f32, no FMA) -> L1 key -> one 32 B hash-sector probe -> f64 residual gate;
//                  per 256-query tile: compacted accepted-query list + count            (48 B / query algorithmic)
//   k_icp_pko1   : one CTA: tile-count scan, C < min check, iteration-0 sigma/6 scale, libstdc++-exact
//                  sample draw from the hit tables, k-means, EM (3-component GMM)
//   k_icp_pko2   : one CTA per alpha candidate: JS divergence; last CTA takes the arg-min -> Huber delta
//   k_icp_gn     : per accepted query: residual, Jacobian, weight, 28 unique sums; warp-shuffle + shared-memory
//                  block reduction in f64; the last CTA adds the block partials in fixed order, solves the 6x6
//                  system (pivoted LDL^T, f32), applies T <- T * (Exp(dw), dt) and sets the convergence flag.
// Later iterations turn into no-ops once the state's `done` flag is set.
#include <climits>
#include <cstdlib>
#include <cstring>
#define B2LO_TL_FILE 2
#include "b2lo_internal.h"
#include "b2lo_launch.cuh"
#include "b2lo_knn.cuh"

namespace b2 {

constexpr int TILE = 256;
// the slice of the scan parameter block an optimize outside the per-scan driver refreshes: initial pose + forced scale
constexpr size_t SP_POSE_BYTES = offsetof(ScanParams, force_scale) + sizeof(double) - offsetof(ScanParams, T_init);
constexpr int PKO_THREADS = 256;
constexpr int MAXS = 128;  // max GMM sample size / alpha candidates handled

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void transform_point(const float* R, const float* t, float x, float y, float z, float* w) {
  // Matrix4f * (x,y,z,1): ((c0*x + c1*y) + c2*z) + c3*1   (src/util/PointCloudUtils.cpp:120-121)
  w[0] = ((R[0] * x + R[1] * y) + R[2] * z) + t[0] * 1.0f;
  w[1] = ((R[3] * x + R[4] * y) + R[5] * z) + t[1] * 1.0f;
  w[2] = ((R[6] * x + R[7] * y) + R[8] * z) + t[2] * 1.0f;
}

// state of a fresh optimize (one thread): pose = initial transform, counters cleared
__device__ __forceinline__ void icp_state_begin(IcpState* st, const ScanParams* sp) {
  for (int i = 0; i < 16; ++i) st->T_init[i] = sp->T_init[i];
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) st->R[i * 3 + j] = st->T_init[i * 4 + j]; st->t[i] = st->T_init[i * 4 + 3]; }
  const double fs = sp->force_scale;
  st->iter = 0; st->done = 0; st->status = B2LO_OK; st->n_corr = 0; st->scale = fs > 0.0 ? fs : 1.0; st->scale_forced = fs > 0.0 ? 1 : 0;
  st->delta = 0.0; st->ticket = 0u; st->ticket_corr = 0u;
  st->num_iterations = 0; st->converged = 0; st->initial_cost = 0.0; st->final_cost = 0.0; st->em_iters = 0; st->kmeans_iters = 0;
}

// returns slot (>=0) if the L1 voxel of w holds a surfel; fills n, c and the key taps
__device__ __forceinline__ int surfel_probe(const MapDev& M, const float* w, float* n, float* c, int* key3, unsigned long long* morton) {
  int kx = voxel_coord(w[0], M.scale1), ky = voxel_coord(w[1], M.scale1), kz = voxel_coord(w[2], M.scale1);
  if (key3) { key3[0] = kx; key3[1] = ky; key3[2] = kz; }
  unsigned long long key = key_pack(kx, ky, kz);
  if (morton) *morton = key_morton(kx, ky, kz);   // debug tap: the reference's VoxelKeyHash value
  if (!key_in_range(kx, ky, kz)) return -1;
  uint32_t mask = (1u << M.l1_log2cap) - 1u;
  uint32_t s = hash_slot(key, M.l1_log2cap);
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    const float4* e = reinterpret_cast<const float4*>(&M.l1_tab[s]);
    float4 a = __ldg(e);  // key (8 B), n.x, n.y
    unsigned long long k = ((unsigned long long)__float_as_uint(a.y) << 32) | (unsigned long long)__float_as_uint(a.x);
    if (k == KEY_EMPTY) return -1;
    if (k != KEY_TOMB && (k & KEY_MASK) == key) {
      if (!(k & SURFEL_BIT)) return -1;
      float4 b = __ldg(e + 1);  // n.z, c.x, c.y, c.z
      n[0] = a.z; n[1] = a.w; n[2] = b.x; c[0] = b.y; c[1] = b.z; c[2] = b.w;
      return (int)s;
    }
    s = (s + 1) & mask;
  }
  return -1;
}

__device__ __forceinline__ double gate_residual(const float* n, const float* c, const float* w) {
  // |normal.dot(p_world_d - plane_point)| in f64 (ICP.cpp:623-628); Vector3d dot reduces as (a + b) + c
  double d0 = (double)w[0] - (double)c[0], d1 = (double)w[1] - (double)c[1], d2 = (double)w[2] - (double)c[2];
  return fabs(((double)n[0] * d0 + (double)n[1] * d1) + (double)n[2] * d2);
}

// block-wide sums of two doubles (warp shuffles + one shared exchange); result valid in thread 0
__device__ __forceinline__ void block_sum2_t0(double& a, double& b, double* smd /*>= 16*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (lane == 0) { smd[2 * w] = a; smd[2 * w + 1] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < nw; ++i) { sa += smd[2 * i]; sb += smd[2 * i + 1]; }
    a = sa; b = sb;
  }
  __syncthreads();
}

// K2.  One query per thread, TILE queries per compaction tile, persistent CTAs striding over the tiles.  Each query is a
// chain of two dependent memory round trips (16 B query point -> key -> 32 B surfel sector), so the loop is software-pipelined
// three deep: while tile t is gated and compacted, the sector loads of tile t+G and the point loads of tile t+2G are in flight
// (G = gridDim.x).  Both 16 B halves of a sector are requested together, all loop-invariant loads (done flag, pose, count,
// first point) are issued before the first use, and compaction costs one barrier per tile (warp ballots + one shared exchange).
// Accepted queries are written in ascending query order per tile (cidx) with the tile count (tilecnt) and the tile's raw
// residual moments (tilesum); k_icp_pko1 scans the tile counts.
// Algorithmic traffic per query: 16 B query + 32 B surfel sector (+ 12 B of per-query results: slot, f64 residual).
struct CorrStage {   // a query between its key computation and its gate
  float w[3];
  unsigned long long key;
  uint32_t hs;
  float4 ea, eb;
};
__device__ __forceinline__ void corr_issue(const MapDev& M, const float* sR, const float* sT, float4 p, bool valid, CorrStage& q) {
  transform_point(sR, sT, p.x, p.y, p.z, q.w);
  // (true f32 divisions: a division-free floor with an exact fallback near cell faces - voxel_coord_fast in b2lo_dev.cuh - saves ~40
  // instructions per query but measured SLOWER here, 41.2 vs 37.6 us per 2^20 random probes on the 10^7-voxel map: the three fallback
  // branches disturb the schedule of this software-pipelined loop more than the shorter common path gains)
  int kx = voxel_coord(q.w[0], M.scale1), ky = voxel_coord(q.w[1], M.scale1), kz = voxel_coord(q.w[2], M.scale1);
  q.key = (valid && key_in_range(kx, ky, kz)) ? key_pack(kx, ky, kz) : KEY_TOMB;   // TOMB never matches: no surfel
  q.hs = hash_slot(q.key, M.l1_log2cap);
  const float4* e = reinterpret_cast<const float4*>(&M.l1_tab[q.hs]);
  q.ea = __ldg(e);        // key (8 B), n.x, n.y
  q.eb = __ldg(e + 1);    // n.z, c.x, c.y, c.z  (same 32 B sector)
}
struct PkoTables;
// the PKO fit (definition below, at k_icp_pko1): with FUSE the CTA that finishes LAST continues straight into it, so one Gauss-Newton
// iteration is three launches instead of four (one kernel boundary and one launch less on the critical path; a converged iteration's
// no-op launch disappears with it).  The body reads what the other CTAs wrote through L2 (__ldcg) behind a fence and a ticket.
__device__ __noinline__ void pko1_body(const int* d_npts, IcpState* st, IcpParams prm, const double* res, const int* cidx, const int* tilecnt,
                                       int* tileoff, const PkoTables* T, const int* hits, double* gmm_out, const double* ext_sample, int ext_C,
                                       double ext_scale, const double* tilesum, const double* ext_plan = nullptr);
template <int DEPTH, int MINB, bool FUSE>
struct k_icp_corr { static __device__ __forceinline__ void run(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_npts, IcpState* st,
                                                   IcpParams prm, double* res, int* slot_out, int* cidx, int* tilecnt, double* tilesum,
                                                   int* tileoff, const PkoTables* T, const int* hits, double* gmm_out, const ScanParams* sp_first) { TL_START();
  // sp_first != nullptr (FUSE only): this is the first correspondence pass of an optimize and no k_icp_begin ran - the pose comes from the
  // scan parameter block, a stale `done` flag is ignored, and the elected last CTA initialises the state before it enters the fit
  __shared__ float sR[9], sT[3];
  __shared__ int s_cnt[2][TILE / 32];
  __shared__ double s_sum[TILE / 32][2];
  __shared__ double s_tsum[FUSE ? 2 : 1][TILE / 32][2];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int G = gridDim.x;
  double m1 = 0.0, m2 = 0.0;
  // every loop-invariant load and the first point go out together (the buffers are sized in whole tiles, see ctx_reserve_points)
  const bool first = FUSE && sp_first != nullptr;
  const int done = first ? 0 : st->done;
  const int npts = *d_npts;
  float pose_v = 0.0f;
  if (tid < 12) {
    if (first) pose_v = tid < 9 ? sp_first->T_init[(tid / 3) * 4 + tid % 3] : sp_first->T_init[(tid - 9) * 4 + 3];
    else pose_v = tid < 9 ? st->R[tid] : st->t[tid - 9];
  }
  int tile = blockIdx.x;
  float4 p1 = pts[tile * TILE + tid];
  if (done) return;
  if (tid < 9) sR[tid] = pose_v; else if (tid < 12) sT[tid - 9] = pose_v;
  __syncthreads();
  const int ntiles = (npts + TILE - 1) / TILE;
  // CTAs without a tile leave at once (also from the fused kernel's election, which counts the active CTAs only; see k_icp_gn)
  const unsigned nact = (unsigned)(ntiles < 1 ? 1 : (ntiles < G ? ntiles : G));
  if (blockIdx.x >= nact) return;
  const uint32_t mask = (1u << M.l1_log2cap) - 1u;
  CorrStage cur;
  if (tile < ntiles) corr_issue(M, sR, sT, p1, tile * TILE + tid < npts, cur);
  if (DEPTH >= 3 && tile + G < ntiles) p1 = pts[(tile + G) * TILE + tid];
  int ph = 0;
  for (; tile < ntiles; tile += G) {
    const int i = tile * TILE + tid;
    // point of tile + 2G
    float4 p2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (DEPTH >= 3) { if (tile + 2 * G < ntiles) p2 = pts[(tile + 2 * G) * TILE + tid]; }
    else if (tile + G < ntiles) p1 = pts[(tile + G) * TILE + tid];
    // gate of this tile
    int s = -1;
    double r = 0.0;
    if (cur.key != KEY_TOMB) {
      uint32_t h = cur.hs;
      float4 a = cur.ea, b = cur.eb;
      for (uint32_t probe = 0; probe <= mask; ++probe) {
        unsigned long long k = ((unsigned long long)__float_as_uint(a.y) << 32) | (unsigned long long)__float_as_uint(a.x);
        if (k == KEY_EMPTY) break;
        if (k != KEY_TOMB && (k & KEY_MASK) == cur.key) {
          if (k & SURFEL_BIT) {
            float n[3] = {a.z, a.w, b.x}, c[3] = {b.y, b.z, b.w};
            r = gate_residual(n, c, cur.w);
            if (!(r > prm.max_dist)) s = (int)h;
          }
          break;
        }
        h = (h + 1) & mask;
        const float4* e = reinterpret_cast<const float4*>(&M.l1_tab[h]);
        a = __ldg(e); b = __ldg(e + 1);
      }
    }
    const bool in = i < npts;
    // sector loads of tile + G
    CorrStage nxt;
    const bool more = tile + G < ntiles;
    if (more) corr_issue(M, sR, sT, p1, (tile + G) * TILE + tid < npts, nxt);
    // results + compaction of this tile: ascending query order, one barrier
    if (in) { slot_out[i] = s; res[i] = r; }
    const bool ok = s >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    // sum r, sum r^2 over the accepted queries (residual scale, ICP.cpp:304-316): carried per thread across this CTA's tiles and reduced
    // ONCE behind the loop (20 shuffles + 10 f64 adds per tile and warp less); the CTA's total is filed under its first tile
    if (ok) { m1 += r; m2 += r * r; }
    if (FUSE) {
      // scan-sized inputs (the per-scan graph): the moments are filed per TILE, so that the result does not depend on how many tiles a
      // CTA walks - a lock-step batch starts fewer CTAs per sequence than a lone sequence and must produce the same bits
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
      if (lane == 0) { s_tsum[ph][wrp][0] = m1; s_tsum[ph][wrp][1] = m2; }
      m1 = 0.0; m2 = 0.0;
    }
    if (lane == 0) s_cnt[ph][wrp] = __popc(bal);
    __syncthreads();
    const int cw = lane < TILE / 32 ? s_cnt[ph][lane] : 0;            // lane l holds the count of warp l
    const int total = __reduce_add_sync(0xffffffffu, cw);
    const int before = __reduce_add_sync(0xffffffffu, lane < wrp ? cw : 0);
    if (ok) cidx[tile * TILE + before + __popc(bal & ((1u << lane) - 1u))] = i;
    if (tid == 0) {
      tilecnt[tile] = total;
      if (FUSE) {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < TILE / 32; ++w2) { sa += s_tsum[ph][w2][0]; sb += s_tsum[ph][w2][1]; }
        tilesum[2 * tile] = sa; tilesum[2 * tile + 1] = sb;
      } else if (tile != (int)blockIdx.x) { tilesum[2 * tile] = 0.0; tilesum[2 * tile + 1] = 0.0; }
    }
    ph ^= 1;   // the other buffer is rewritten only after the next barrier: nobody still reads it then
    if (more) cur = nxt;
    p1 = p2;
  }
  if (!FUSE && (int)blockIdx.x < ntiles) {   // this CTA's residual moments -> the slot of its first tile
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    __syncthreads();   // the last tile's readers of s_sum's neighbours are done
    if (lane == 0) { s_sum[wrp][0] = m1; s_sum[wrp][1] = m2; }
    __syncthreads();
    if (tid == 0) {
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < TILE / 32; ++w2) { sa += s_sum[w2][0]; sb += s_sum[w2][1]; }
      tilesum[2 * blockIdx.x] = sa; tilesum[2 * blockIdx.x + 1] = sb;
    }
  }
  if (FUSE) {
    __shared__ int s_last;
    __threadfence();                 // this CTA's tiles are visible before its ticket is
    __syncthreads();
    if (tid == 0) {
      const unsigned int t = atomicAdd(&st->ticket_corr, 1u);
      s_last = (t == nact - 1);
      if (s_last) st->ticket_corr = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (first) {   // k_icp_begin, by the one CTA that is still running
      if (tid == 0) icp_state_begin(st, sp_first);
      __syncthreads();
    }
    pko1_body(d_npts, st, prm, res, cidx, tilecnt, tileoff, T, hits, gmm_out, nullptr, 0, 0.0, tilesum);
  }
} };

// ---- KDTree-mode correspondence (K3) ----------------------------------------------------------------
// search: one warp per query probes the cell shells of the L0 hash (27, then 98 cells, one lane per cell); unresolved queries are queued
__global__ void __launch_bounds__(TILE) k_knn_search(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_npts, IcpState* st,
                                                     int* knn_idx, int* knn_n, int* unres, int* n_unres) { TL_START();
  if (st->done) return;
  __shared__ float sR[9], sT[3];
  if (threadIdx.x < 9) sR[threadIdx.x] = st->R[threadIdx.x];
  if (threadIdx.x < 3) sT[threadIdx.x] = st->t[threadIdx.x];
  __syncthreads();
  const int npts = *d_npts;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < npts; i += nwarps) {
    float4 p = pts[i];
    float w[3];
    transform_point(sR, sT, p.x, p.y, p.z, w);
    Top5 top;
    bool exact = knn_rings_warp(M, w, top);
    if (lane == 0) {
      if (exact) {
        for (int k = 0; k < KNN_K; ++k) knn_idx[i * KNN_K + k] = top.id[k];
        knn_n[i] = top.n;
      } else {
        knn_n[i] = -1;
        unres[atomicAdd(n_unres, 1)] = i;
      }
    }
  }
}
// exact scan for the queued queries: one warp per query over the dense L0 centroid stream
__global__ void __launch_bounds__(256) k_knn_brute(MapDev M, const float4* __restrict__ pts, IcpState* st, int* knn_idx, int* knn_n,
                                                   const int* __restrict__ unres, const int* __restrict__ n_unres) { TL_START();
  if (st->done) return;
  const int nu = *n_unres;
  const int n0 = M.ctr[0];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int u = warp; u < nu; u += nwarps) {
    int i = unres[u];
    float4 p = pts[i];
    float w[3];
    transform_point(st->R, st->t, p.x, p.y, p.z, w);
    Top5 top;
    knn_brute_warp(M, n0, w, top);
    if ((threadIdx.x & 31) == 0) {
      for (int k = 0; k < KNN_K; ++k) knn_idx[i * KNN_K + k] = top.id[k];
      knn_n[i] = top.n;
    }
  }
}
// plane fit + gate + per-tile compaction (same outputs as k_icp_corr, plus the per-query plane)
__global__ void __launch_bounds__(TILE) k_knn_gate(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_npts, IcpState* st, IcpParams prm,
                                                   const int* __restrict__ knn_idx, const int* __restrict__ knn_n, int* n_unres, double* res,
                                                   int* slot_out, int* cidx, int* tilecnt, float4* plane, double* tilesum) { TL_START();
  if (st->done) return;
  __shared__ int sm[40];
  __shared__ double smd2[16];
  __shared__ float sR[9], sT[3];
  if (threadIdx.x < 9) sR[threadIdx.x] = st->R[threadIdx.x];
  if (threadIdx.x < 3) sT[threadIdx.x] = st->t[threadIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_unres = 0;  // k_knn_brute of this iteration has finished
  __syncthreads();
  const int npts = *d_npts;
  const int ntiles = (npts + TILE - 1) / TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int i = tile * TILE + threadIdx.x;
    int ok = 0;
    double a1 = 0.0, a2 = 0.0;
    if (i < npts) {
      float4 p = pts[i];
      float w[3], n[3] = {0, 0, 0}, c[3] = {0, 0, 0};
      transform_point(sR, sT, p.x, p.y, p.z, w);
      Top5 top;
      top.init();
      top.n = knn_n[i] < 0 ? 0 : knn_n[i];
      for (int k = 0; k < top.n; ++k) top.id[k] = knn_idx[i * KNN_K + k];
      double r = 0.0;
      int state = knn_fit(M, top, w, prm.max_dist, n, c, &r);
      ok = (state == 2);
      if (ok) { a1 = r; a2 = r * r; }
      slot_out[i] = ok ? 0 : -1;
      res[i] = r;
      plane[2 * i] = make_float4(n[0], n[1], n[2], c[0]);
      plane[2 * i + 1] = make_float4(c[1], c[2], 0.0f, 0.0f);
    }
    int total;
    int off = block_excl_scan(ok, &total, sm);
    if (ok) cidx[tile * TILE + off] = i;
    block_sum2_t0(a1, a2, smd2);
    if (threadIdx.x == 0) { tilecnt[tile] = total; tilesum[2 * tile] = a1; tilesum[2 * tile + 1] = a2; }
  }
}
// parity tap of the KDTree-mode correspondence at the pose held in st
__global__ void k_knn_taps(MapDev M, const float4* __restrict__ pts, int npts, IcpState* st, double max_dist, const int* __restrict__ knn_idx,
                           const int* __restrict__ knn_n, int* idx_out, float* d2_out, int* found, int* state, float* nout, float* cout, double* res) { TL_START();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  float4 p = pts[i];
  float w[3], n[3] = {0, 0, 0}, c[3] = {0, 0, 0};
  transform_point(st->R, st->t, p.x, p.y, p.z, w);
  Top5 top;
  top.init();
  top.n = knn_n[i] < 0 ? 0 : knn_n[i];
  for (int k = 0; k < KNN_K; ++k) {
    int id = k < top.n ? knn_idx[i * KNN_K + k] : -1;
    top.id[k] = id;
    idx_out[i * KNN_K + k] = id;
    float d2 = 0.0f;
    if (id >= 0) { float4 cc = M.l0_cent[id]; d2 = knn_dist2(w, cc.x, cc.y, cc.z); }
    d2_out[i * KNN_K + k] = d2;
  }
  found[i] = top.n;
  double r = 0.0;
  int stt = knn_fit(M, top, w, max_dist, n, c, &r);
  state[i] = stt; res[i] = r;
  for (int a = 0; a < 3; ++a) { nout[i * 3 + a] = n[a]; cout[i * 3 + a] = c[a]; }
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_d(double v, double* smd /*>=33*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) smd[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double x = lane < nw ? smd[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) smd[32] = x;
  }
  __syncthreads();
  double r = smd[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ double pko_kernel(int type, double r, double delta) {
  if (type == 0) { double a = fabs(r); return a <= delta ? 1.0 : delta / a; }
  double e2 = r * r, d2 = delta * delta;
  return d2 / (d2 + e2);
}
__device__ __forceinline__ double gauss_pdf(double x, double mean, double var) {
  if (var <= 0.0) return 0.0;
  double diff = x - mean;
  double ex = -0.5 * (diff * diff) / var;
  double nrm = 1.0 / sqrt(2.0 * 3.14159265358979323846 * var);
  return nrm * exp(ex);
}

// sample position j of std::shuffle(iota(n), mt19937(42)) from the hit tables (see b2lo_pko_host.cpp)
__device__ int shuffled_head(const PkoTables* T, const int* hits, int n, int j) {
  int mode = n >= 65536 ? 2 : ((n & 1) ? 1 : 0);
  int lo = T->hit_off[mode][j], hi = T->hit_off[mode][j + 1];
  // largest hit < n (lists are ascending and short)
  int best = -1;
  for (int q = lo; q < hi; ++q) { int v = hits[q]; if (v < n) best = v; else break; }
  if (best >= 0) return best;
  int pos = j;
  int top = n - 1 < 127 ? n - 1 : 127;
  for (int i = top; i >= 1; --i) {
    int r = T->head_r[mode][i];
    if (pos == i) pos = r; else if (pos == r) pos = i;
  }
  return pos;
}

// exp(x) for the Gaussian terms of the EM (x <= 0).  Branch-free and short: the argument is clamped at -708 (exp = 3e-308, an
// additive nothing next to the O(1) mixture sums); x = k ln2/2^10 + r with |r| <= ln2/2^11 = 3.4e-4 by the 1.5*2^52 rounding trick (no
// float->int conversion) and a two-step Cody-Waite reduction; exp(x) = 2^q * T1[j1] * T2[j2] * (1 + r + r^2/2 + r^3/6 + r^4/24),
// k = 2^10 q + 2^5 j1 + j2, with two 32-entry tables in shared memory (2^(j/32), 2^(j/1024): at most 2-way bank conflicts on the
// per-lane random look-ups; 256-entry tables would allow a quadratic but conflict ~6-way); truncation 4e-20.
// 11 f64 instructions with a dependent chain of 9, against ~30 / ~25 for the library routine - the EM fixed point is one long chain
// of dependent f64 operations.  Relative error ~3e-16.  Same-box A/B against the previous one-table degree-6 form (14 instructions),
// together with the third-order reciprocal and the fused accumulations below: 1250 -> 1202 SM cycles per EM iteration; the loop is
// bound by the dependent chain (rsqrt -> exp -> exchange -> rcp -> 5 shuffle levels -> rcp), not by f64 issue slots.
__device__ __forceinline__ double em_exp(double x, const double* __restrict__ s_t1, const double* __restrict__ s_t2) {
  x = fmax(x, -708.0);
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double tm = fma(x, 1477.3197218702985, MAGIC);        // 2^10 / ln 2
  const double kd = tm - MAGIC;
  const int k = __double2loint(tm);                           // round(2^10 x / ln 2) in two's complement
  double r = fma(-kd, 6.769015435155716e-04, x);             // ln2/2^10 hi
  r = fma(-kd, 2.264694154146777e-20, r);                       // ln2/2^10 lo
  const double t = s_t1[(k >> 5) & 31] * s_t2[k & 31];
  const double r2 = r * r;
  const double u = fma(r, 1.0 / 6.0, 0.5);
  const double v = fma(r2, 1.0 / 24.0, u);
  const double q = fma(r2, v, r);                             // e^r - 1 up to r^5/120 (4e-20)
  const double pr = fma(t, q, t);
  return __hiloint2double(__double2hiint(pr) + (k >> 10) * 1048576, __double2loint(pr));
}

// Reciprocal / reciprocal square root for the EM chain: hardware seed (rcp/rsqrt.approx.ftz.f64, ~2^-20) + one third-order step /
// two Newton steps (error ~1e-16, not correctly rounded).  4 / 9 dependent f64 operations instead of the ~20-40 instructions of the IEEE
// library routines; the EM only needs the ~1e-14 agreement discussed at k_icp_pko1.
__device__ __forceinline__ double em_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);       // one third-order step: r (1 + e + e^2), error e^3 ~ 2^-60
  return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double em_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);   // y += y (1/2 - x y^2 / 2)
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
}

constexpr int PKO_TOFF = 1024;
// Exclusive scan of the per-tile accepted counts by ONE block when there are many tiles (dense scans: thousands).  Warp w owns a
// contiguous run of tiles; pass 1 sums it with independent loads, one barrier orders the warp totals, pass 2 re-reads the run (L2 hits)
// and writes the offsets with a warp-shuffle scan per 128 tiles - two barriers in all instead of two per blockDim tiles, and no load
// sits on the critical path of a barrier.  Integer sums: the result is the same in any order.  Returns the total.
__device__ __forceinline__ int tile_scan_warps(const int* tilecnt, int* tileoff, int ntiles, int* sm /* >= 32 ints */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int per = ((ntiles + nw * 128 - 1) / (nw * 128)) * 128;
  const int b = wid * per, e = (b + per < ntiles) ? b + per : ntiles;
  int s = 0;
  for (int t = b + lane * 4; t < e; t += 128) {
#pragma unroll
    for (int u = 0; u < 4; ++u) if (t + u < e) s += __ldcg(&tilecnt[t + u]);
  }
  s = __reduce_add_sync(0xffffffffu, s);
  __syncthreads();
  if (lane == 0) sm[wid] = s;
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < nw; ++w) { const int v = sm[w]; if (w < wid) base += v; tot += v; }
  for (int t0 = b; t0 < e; t0 += 128) {
    const int t = t0 + lane * 4;
    int c[4], ls = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { c[u] = (t + u < e) ? __ldcg(&tilecnt[t + u]) : 0; ls += c[u]; }
    int incl = ls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    int ex = base + incl - ls;
#pragma unroll
    for (int u = 0; u < 4; ++u) if (t + u < e) { tileoff[t + u] = ex; ex += c[u]; }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncthreads();
  return tot;
}
// last tile whose offset is <= ci (offsets are non-decreasing, tileoff[0] = 0 <= ci).  Eight-way search: the seven probes of a step are
// independent loads, so a table of thousands of tiles in L2 costs 4 memory round trips instead of the 12-13 of a binary search.
__device__ __forceinline__ int tile_of(const int* tileoff, int ntiles, int ci) {
  int lo = 0, hi = ntiles - 1;
  while (lo < hi) {
    const int step = (hi - lo + 8) >> 3;
    int v[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) { const int idx = lo + (u + 1) * step; v[u] = tileoff[idx < hi ? idx : hi]; }
    int nl = lo, nh = hi;
#pragma unroll
    for (int u = 6; u >= 0; --u) { const int idx = lo + (u + 1) * step < hi ? lo + (u + 1) * step : hi; if (v[u] > ci) nh = idx - 1; }
#pragma unroll
    for (int u = 0; u < 7; ++u) { const int idx = lo + (u + 1) * step < hi ? lo + (u + 1) * step : hi; if (v[u] <= ci && idx <= nh) nl = idx; }
    lo = nl; hi = nh;
  }
  return lo;
}
constexpr int EM_SPL = MAXS / 32;   // samples per lane in the EM (4): sample index = lane + 32 k
static_assert(MAXS % 32 == 0, "EM layout needs a multiple of 32 samples");

template <int V> __device__ __forceinline__ void warp_sum_multi(double (&v)[V]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}

// One CTA: accepted-count scan, the C < min test, the iteration-0 residual scale, the libstdc++-exact sample draw,
// k-means and the 3-component EM of AdaptiveMEstimator::fit_gmm (AdaptiveMEstimator.cpp:294-485).
// The EM fixed point is a chain of dependent f64 operations (30-100 iterations per call, capped at 100), so the
// kernel is laid out for LATENCY (measured on B200: DFMA 8.5, SHFL64+DADD 36, f64 rcp/rsqrt ~75, bar.sync ~45 cycles):
//   * warp c owns mixture component c; each lane carries 4 samples (independent chains hide the exp latency);
//   * per iteration ONE shared-memory exchange + ONE named barrier: the three warps publish their unnormalised
//     densities p_c(x_i) (and the mean shift of their previous M-step), then every warp normalises its own
//     responsibilities and reduces its own three sums with warp shuffles only;
//   * the convergence test of iteration t rides the exchange of iteration t+1 (the speculative E-step is dropped);
//   * the variance update uses sum r x^2 / nk - mean^2 (algebraically the reference's second pass), reciprocals are
//     computed once per component, and responsibilities are p * rcp(sum) (a true f64 divide of the tiny far-component
//     terms takes the slow denormal path).
// The reference sums left to right and divides; the two evaluation orders agree to ~1e-14 relative, which can move
// the discrete outputs (iteration counts, arg-min alpha) only on near-exact ties - tests/ assert they match the oracle.
__device__ __noinline__ void pko1_body(const int* d_npts, IcpState* st, IcpParams prm, const double* res, const int* cidx, const int* tilecnt,
                                       int* tileoff, const PkoTables* T, const int* hits, double* gmm_out, const double* ext_sample, int ext_C,
                                       double ext_scale, const double* tilesum, const double* ext_plan) {
  if (st->done) return;
  if (ext_plan) { ext_C = (int)ext_plan[1]; ext_scale = ext_plan[2]; }   // device-ordered point-sharded mode: the plan of k_shard_plan
  if (threadIdx.x == 0) TL_HERE();   // fit begins
  __shared__ int sm[40];
  __shared__ double smd[40];
  __shared__ double s_x[MAXS];
  __shared__ int s_head[MAXS];
  __shared__ double s_p[2][3][MAXS];
  __shared__ double s_t1[32], s_t2[32];   // exp tables: 2^(j/32), 2^(j/1024)
  __shared__ double s_dm[2][4];
  __shared__ double s_par[3][4];
  __shared__ int s_toff[PKO_TOFF];   // tile offsets of the first PKO_TOFF tiles: the sample draw's binary search stays on chip
  const int tid = threadIdx.x;
  const long long c0 = clock64();
  long long c1 = c0, c2 = c0;
  int ns;
  if (ext_sample) {
    // point-sharded mode: the globally drawn, already normalised sample arrives from the all-reduce (b2lo_icp_shard_*)
    if (!prm.use_pko) { if (tid == 0) { st->delta = prm.robust_delta; st->em_iters = 0; st->kmeans_iters = 0; st->scale = ext_scale; st->n_corr = ext_C; } return; }
    ns = T->sample_size < ext_C ? T->sample_size : ext_C;
    if (tid < ns) s_x[tid] = ext_sample[tid];
    if (tid >= MAXS && tid < MAXS + 32) { s_t1[tid - MAXS] = T->exp2_t1[tid - MAXS]; s_t2[tid - MAXS] = T->exp2_t2[tid - MAXS]; }
    if (tid == 0) { st->scale = ext_scale; st->n_corr = ext_C; }
    __syncthreads();
  } else {
  const int npts = *d_npts;
  const int ctile = prm.ctile;
  const int ntiles = (npts + ctile - 1) / ctile;
  // 1. exclusive scan of the per-tile accepted counts
  int base = 0;
  if (ntiles > PKO_TOFF) base = tile_scan_warps(tilecnt, tileoff, ntiles, sm);   // dense scans; s_toff is not used beyond PKO_TOFF tiles
  else
  for (int t0 = 0; t0 < ntiles; t0 += blockDim.x) {
    int t = t0 + tid;
    int c = t < ntiles ? __ldcg(&tilecnt[t]) : 0;
    int tot;
    int e = block_excl_scan(c, &tot, sm);
    if (t < ntiles) { tileoff[t] = base + e; if (t < PKO_TOFF) s_toff[t] = base + e; }
    base += tot;
  }
  const int C = base;
  if (tid == 0) { st->n_corr = C; st->n_blocks = ntiles; }
  if (C < prm.min_corr) {  // ICP.cpp:298-302
    if (tid == 0) { st->done = 2; st->status = B2LO_S_INSUFFICIENT; }
    return;
  }
  // the precomputed image of std::shuffle's swap partners (the sample draw below) covers index vectors of up to 2^22 correspondences
  if (prm.use_pko && C > (1 << 22)) { if (tid == 0) { st->done = 2; st->status = B2LO_E_CAPACITY; } return; }
  // 2. residual normalisation scale, first iteration only (ICP.cpp:304-316)
  c1 = clock64();
  double scale = st->scale;
  if (st->iter == 0 && !st->scale_forced) {
    // population sigma / 6 from the per-tile raw moments K2 left behind (var = E[r^2] - mean^2)
    double a1 = 0.0, a2 = 0.0;
    for (int t = tid; t < ntiles; t += blockDim.x) { a1 += __ldcg(&tilesum[2 * t]); a2 += __ldcg(&tilesum[2 * t + 1]); }
    a1 = block_sum_d(a1, smd);
    a2 = block_sum_d(a2, smd);
    const double mean = a1 / (double)C;
    const double var = fmax(a2 / (double)C - mean * mean, 0.0);
    scale = sqrt(var) / 6.0;
    if (tid == 0) st->scale = scale;
  }
  if (!prm.use_pko) { if (tid == 0) { st->delta = prm.robust_delta; st->em_iters = 0; st->kmeans_iters = 0; } return; }
  const double sdiv = fmax(scale, 1e-6);
  c2 = clock64();
  // 3. the sample: residuals[idx[0..ns)] of the shuffled index vector (AdaptiveMEstimator.cpp:319-331)
  ns = T->sample_size < C ? T->sample_size : C;
  const int mode = C >= 65536 ? 2 : ((C & 1) ? 1 : 0);
  if (tid < MAXS) s_head[tid] = T->head_r[mode][tid];
  if (tid >= MAXS && tid < MAXS + 32) { s_t1[tid - MAXS] = T->exp2_t1[tid - MAXS]; s_t2[tid - MAXS] = T->exp2_t2[tid - MAXS]; }
  __syncthreads();
  if (tid < ns) {
    // position tid of std::shuffle(iota(C), mt19937(42)): the LARGEST swap partner i in [128, C) that hit position tid
    // (ascending hit list), else a backward trace through the first min(C,128)-1 swaps
    const int lo = T->hit_off[mode][tid], hi = T->hit_off[mode][tid + 1];
    int best = -1;
    for (int q0 = lo; q0 < hi; q0 += 8) {
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (q0 + u < hi) ? hits[q0 + u] : 0x7fffffff;
#pragma unroll
      for (int u = 0; u < 8; ++u) if (v[u] < C) best = v[u];
      if (v[7] >= C) break;
    }
    int ci = best;
    if (best < 0) {
      if (C >= MAXS) ci = T->head_pos[mode][tid];   // all 127 head swaps apply: the trace is a constant of the mode
      else {
        int pos = tid;
        for (int i = C - 1; i >= 1; --i) { int r = s_head[i]; pos = (pos == i) ? r : ((pos == r) ? i : pos); }
        ci = pos;
      }
    }
    int tl = 0, th = ntiles - 1;  // last tile with tileoff <= ci
    if (ntiles <= PKO_TOFF) {
      while (tl < th) { int mid = (tl + th + 1) >> 1; if (s_toff[mid] <= ci) tl = mid; else th = mid - 1; }
    } else tl = tile_of(tileoff, ntiles, ci);
    int q = __ldcg(&cidx[tl * ctile + (ci - tileoff[tl])]);
    s_x[tid] = __ldcg(&res[q]) / sdiv;
  }
  __syncthreads();
  }
  const long long c3 = clock64();
  if (tid >= 96) return;
  const int lane = tid & 31, c = tid >> 5;   // warp c <-> mixture component / k-means cluster c
  double x[EM_SPL], xx[EM_SPL];
  bool act[EM_SPL];
#pragma unroll
  for (int k = 0; k < EM_SPL; ++k) { const int i = lane + 32 * k; act[k] = i < ns; x[k] = act[k] ? s_x[i] : 0.0; xx[k] = x[k] * x[k]; }
  const double inv_ns = 1.0 / (double)ns;
  // 4. k-means (:336-389): mean0 = 0, mean1/2 = sample[dis(gen)]; every warp runs it redundantly (no exchange needed)
  double m1 = s_x[T->kmeans_seed[ns][0]], m2 = s_x[T->kmeans_seed[ns][1]];
  int km_iters = 0;
  double cnt[3] = {0.0, 0.0, 0.0};
  for (;;) {
    ++km_iters;
    double v[2] = {0.0, 0.0};   // sum1, sum2
    int cn = 0;                 // count1 | count2 << 16 (exact in any order, so they travel as integers)
#pragma unroll
    for (int k = 0; k < EM_SPL; ++k) {
      double md = fabs(x[k]);  // |x - mean0|, mean0 = 0
      const double d1 = fabs(x[k] - m1), d2 = fabs(x[k] - m2);
      int cl = 0;
      if (d1 < md) { md = d1; cl = 1; }
      if (d2 < md) { md = d2; cl = 2; }
      cl = act[k] ? cl : 0;
      v[0] += (cl == 1) ? x[k] : 0.0; v[1] += (cl == 2) ? x[k] : 0.0;   // idle lanes hold x = 0 and add +0.0
      cn += (cl == 1 ? 1 : 0) + (cl == 2 ? 65536 : 0);
    }
    warp_sum_multi<2>(v);
    cn = __reduce_add_sync(0xffffffffu, cn);
    const double k1 = (double)(cn & 0xffff), k2 = (double)(cn >> 16);
    const double n1 = k1 > 0.0 ? v[0] / k1 : v[0];
    const double n2 = k2 > 0.0 ? v[1] / k2 : v[1];
    cnt[1] = k1; cnt[2] = k2; cnt[0] = (double)ns - k1 - k2;
    const bool same = (n1 == m1 && n2 == m2);
    if (same || km_iters >= 100000) break;
    m1 = n1; m2 = n2;
  }
  const long long c4 = clock64();
  // 5. initial variance = population variance of the sample (:392-399); weights = cluster fractions (:402-410)
  double mean = (c == 1) ? m1 : ((c == 2) ? m2 : 0.0), var, wgt = cnt[c] * inv_ns;
  {
    double v2[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EM_SPL; ++k) if (act[k]) { v2[0] += x[k]; v2[1] += xx[k]; }
    warp_sum_multi<2>(v2);
    const double mu = v2[0] * inv_ns;
    var = v2[1] * inv_ns - mu * mu;   // population variance
  }
  // 6. EM (:418-484)
  int em_iters = 100, ph = 0;
  double dm = 1.0;   // |mean shift| of this component in the previous M-step
  for (int it = 0; it < 100; ++it) {
    // E-step, own component: p_c(x_i) = w_c N(x_i; mean_c, var_c)   (gaussian_pdf :675-685: 0 for variance <= 0)
    const bool ok = var > 0.0;
    const double sq = ok ? em_rsqrt(6.283185307179586 * var) : 0.0;  // one rsqrt feeds the normalisation and 1/var = 2 pi sq^2
    const double hiv = -0.5 * (6.283185307179586 * (sq * sq));
    const double coef = wgt * sq;
    double p[EM_SPL];
#pragma unroll
    for (int k = 0; k < EM_SPL; ++k) {
      const double d = x[k] - mean;
      p[k] = (act[k] ? coef : 0.0) * em_exp((d * d) * hiv, s_t1, s_t2);   // idle samples: p = 0, so their responsibility is 0 below (no branch)
      s_p[ph][c][lane + 32 * k] = p[k];
    }
    if (lane == 0) s_dm[ph][c] = dm;
    asm volatile("bar.sync 1, 96;" ::: "memory");
    if (it > 0) {  // convergence test of the previous iteration (:476-482); this iteration's E-step was speculative
      const double change = s_dm[ph][1] + s_dm[ph][2];
      if (change < 1e-6) { em_iters = it; break; }
    }
    // branch-free so that the four reciprocal chains of this lane interleave: idle samples normalise by 1 and are masked by 0
    double sum[EM_SPL], a3[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EM_SPL; ++k) {
      const int i = lane + 32 * k;
      const double t = (s_p[ph][0][i] + s_p[ph][1][i]) + s_p[ph][2][i];
      sum[k] = act[k] ? t : 1.0;
    }
#pragma unroll
    for (int k = 0; k < EM_SPL; ++k) sum[k] = em_rcp(sum[k]);   // sum 0 -> NaN propagates as in the reference
#pragma unroll
    for (int k = 0; k < EM_SPL; ++k) {
      const double r = p[k] * sum[k];
      a3[0] += r; a3[1] = fma(r, x[k], a3[1]); a3[2] = fma(r, xx[k], a3[2]);
    }
    warp_sum_multi<3>(a3);
    // M-step, own component (:436-474); mean0 stays pinned at 0
    const double rnk = em_rcp(a3[0]);
    const double nm = (c == 0) ? 0.0 : a3[1] * rnk;
    const double nv = fmax(a3[2] * rnk - 2.0 * nm * (a3[1] * rnk) + nm * nm, 1e-6);   // sum r (x - nm)^2 / nk
    dm = fabs(nm - mean);
    wgt = a3[0] * inv_ns; mean = nm; var = nv;
    ph ^= 1;
  }
  if (lane == 0) { s_par[c][0] = mean; s_par[c][1] = var; s_par[c][2] = wgt; }
  asm volatile("bar.sync 1, 96;" ::: "memory");
  if (tid == 0) {
    for (int j = 0; j < 3; ++j) { gmm_out[j] = s_par[j][0]; gmm_out[3 + j] = s_par[j][1]; gmm_out[6 + j] = s_par[j][2]; }
    st->em_iters = em_iters; st->kmeans_iters = km_iters;
    const long long c5 = clock64();
    long long* dg = st->dbg + 8 * (st->iter & 1);
    dg[0] = c1 - c0; dg[1] = c2 - c1; dg[2] = c3 - c2; dg[3] = c4 - c3; dg[4] = c5 - c4; dg[5] = em_iters; dg[6] = km_iters;
    TL_HERE();   // EM done
  }
  // 7. P(r_k) of the fitted mixture on the JS grid r_k = dr * (1 + k) (:741-752), shared by all alpha candidates
  for (int k = tid; k < 100; k += 96) {
    const double rr = (T->trunc / 100.0) * (1.0 + (double)k);
    double Pr = 0.0;
    for (int m = 0; m < 3; ++m) Pr += s_par[m][2] * gauss_pdf(rr, s_par[m][0], s_par[m][1]);
    gmm_out[16 + k] = Pr + 1e-10;
  }
}
// the fit as a launch of its own (KDTree mode, loop-closure ICP, point-sharded mode, profiling runs)
struct k_icp_pko1 { static __device__ __forceinline__ void run(const int* __restrict__ d_npts, IcpState* st, IcpParams prm, const double* __restrict__ res,
                                                           const int* __restrict__ slot, const int* __restrict__ cidx, const int* __restrict__ tilecnt,
                                                           int* tileoff, const PkoTables* __restrict__ T, const int* __restrict__ hits, double* gmm_out,
                                                           const double* __restrict__ ext_sample, int ext_C, double ext_scale,
                                                           const double* __restrict__ tilesum, const double* __restrict__ ext_plan) { TL_START();
  (void)slot;
  pko1_body(d_npts, st, prm, res, cidx, tilecnt, tileoff, T, hits, gmm_out, ext_sample, ext_C, ext_scale, tilesum, ext_plan);
} };

// one CTA per alpha candidate i = 1..S (blockIdx.x + 1); thread k handles r_k = dr * (1 + k); the last CTA takes the arg-min
struct k_icp_pko2 { static __device__ __forceinline__ void run(IcpState* st, IcpParams prm, const PkoTables* __restrict__ T, const double* __restrict__ gmm,
                                                   double* js, unsigned int* ticket) { TL_START();
  __shared__ double s_c[4], s_n[4];
  __shared__ int s_i[4];
  __shared__ int s_last;
  const int k = threadIdx.x, lane = k & 31, w = k >> 5;
  // the done flag and this thread's operands go out together
  const int done_v = st->done;
  const int na_all = T->n_alpha;
  const double dr = T->trunc / 100.0;
  const double Pr_v = k < 100 ? __ldcg(&gmm[16 + k]) : 0.0;
  if (done_v || !prm.use_pko) return;
  // one candidate per CTA for a lone sequence (gridDim.x = segments); a lock-step batch starts fewer CTAs per sequence and each walks
  // several candidates - the per-candidate sums are the same either way
  for (int ai = blockIdx.x + 1; ai < na_all; ai += gridDim.x) {
    const double alpha = T->alpha[ai];
    const double pf = T->Z[ai];
    double v = 0.0, c = 0.0;
    if (k < 100) {
      double r = dr * (1.0 + (double)k);
      double Pr = Pr_v;
      double Q = pko_kernel(T->kernel_type, r, alpha) / (pf + 1e-10) + 1e-10;
      double Mx = 0.5 * (Pr + Q);
      double jsd = 0.5 * (Pr * log(Pr / Mx) + Q * log(Q / Mx));
      if (jsd == jsd) { v = jsd; c = 1.0; }   // NaN terms are skipped (:777-779)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane == 0) { s_c[w] = v; s_n[w] = c; }
    __syncthreads();
    if (k == 0) {
      double cost = (s_c[0] + s_c[1]) + (s_c[2] + s_c[3]), cnt = (s_n[0] + s_n[1]) + (s_n[2] + s_n[3]);
      js[ai] = (pf < 1e-10 || cnt == 0.0) ? 1.7976931348623157e308 : cost / cnt;
    }
    __syncthreads();
  }
  if (k == 0) {
    __threadfence();
    unsigned int old = atomicAdd(ticket, 1u);
    s_last = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // arg-min with strict '<' in candidate order (:259-275): smallest cost, lowest index on ties
  const int na = T->n_alpha;
  double best = 1.7976931348623157e308; int bi = 0x7fffffff;
  for (int i = 1 + k; i < na; i += blockDim.x) { double cc = ((volatile double*)js)[i]; if (cc < best) { best = cc; bi = i; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { s_c[w] = best; s_i[w] = bi; }
  __syncthreads();
  if (k == 0) {
    for (int ww = 1; ww < 4; ++ww) if (s_c[ww] < best || (s_c[ww] == best && s_i[ww] < bi)) { best = s_c[ww]; bi = s_i[ww]; }
    // best_cost starts at DBL_MAX and only a strictly smaller cost replaces min_scale_factor
    st->delta = (bi != 0x7fffffff && best < 1.7976931348623157e308) ? T->alpha[bi] : T->min_sf;
    *ticket = 0u;
  }
} };

// ---------------------------------------------------------------------------------------------------
__device__ void gn_finish(IcpState* st, const IcpParams& prm, const double* acc /*28*/) {
  // unpack: 21 lower-triangle entries (row-major a>=b), 6 g, 1 cost
  float H[36], g[6];
  double H64[36];
  int q = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b <= a; ++b) { double v = acc[q++]; H64[a * 6 + b] = v; H64[b * 6 + a] = v; H[a * 6 + b] = (float)v; H[b * 6 + a] = (float)v; }
  for (int a = 0; a < 6; ++a) g[a] = (float)acc[21 + a];
  const int it = st->iter;
  b2lo_iter_trace* tr = (it < B2LO_MAX_ITERS) ? &st->trace[it] : nullptr;
  Pose cur;
  for (int i = 0; i < 9; ++i) cur.R.m[i] = st->R[i];
  for (int i = 0; i < 3; ++i) cur.t[i] = st->t[i];
  if (tr) {
    tr->n_corr = st->n_corr; tr->scale = st->scale; tr->delta = st->delta;
    for (int i = 0; i < 36; ++i) tr->H[i] = H64[i];
    for (int i = 0; i < 6; ++i) tr->g[i] = acc[21 + i];
    tr->cost = acc[27];
    pose_to_T16(cur, tr->T_in);
    tr->em_iters = st->em_iters; tr->kmeans_iters = st->kmeans_iters;
  }
  float mg[6], dx[6];
  for (int a = 0; a < 6; ++a) mg[a] = -g[a];
  TL_HERE();   // finish: unpacked, trace head written
  ldlt6_solve(H, mg, dx);
  TL_HERE();   // finish: LDLT solved
  float dt[3] = {dx[0], dx[1], dx[2]}, dw[3] = {dx[3], dx[4], dx[5]};
  Pose d;
  float wn = sqrtf(sqn3(dw));
  // T <- T * SE3(Exp(dw), dt) (ICP.cpp:426-434, MathUtils.h:144-147).  The reference re-projects the exponential and the
  // product onto SO(3) with an SVD each; both inputs are rotations up to rounding, so one Newton-Schulz step gives the
  // same nearest rotation (b2lo_math.cuh, so3_project_near) without the two Jacobi-sweep chains on this single thread.
  if (wn < 1e-10f) d.R = mat3_identity();
  else {
    Mat3 I = mat3_identity(), Mx;
    if (wn < 1e-6f) { Mat3 K = hat(dw); for (int i = 0; i < 9; ++i) Mx.m[i] = I.m[i] + K.m[i]; }
    else {
      float ti = 1.0f / wn;
      float kv[3] = {dw[0] * ti, dw[1] * ti, dw[2] * ti};
      Mat3 K = hat(kv);
      float sn = sin_f32(wn), omc = 1.0f - cos_f32(wn);
      Mat3 oK;
      for (int i = 0; i < 9; ++i) oK.m[i] = omc * K.m[i];
      Mat3 KK = mat3_mul(oK, K);
      for (int i = 0; i < 9; ++i) Mx.m[i] = (I.m[i] + sn * K.m[i]) + KK.m[i];
    }
    d.R = so3_project_near(Mx);
  }
  d.t[0] = dt[0]; d.t[1] = dt[1]; d.t[2] = dt[2];
  Pose nxt;
  {
    float rt[3];
    mat3_vec(cur.R.m, d.t, rt);
    nxt.t[0] = cur.t[0] + rt[0]; nxt.t[1] = cur.t[1] + rt[1]; nxt.t[2] = cur.t[2] + rt[2];
    nxt.R = so3_project_near(mat3_mul(cur.R, d.R));
  }
  TL_HERE();   // finish: pose updated
  for (int i = 0; i < 9; ++i) st->R[i] = nxt.R.m[i];
  for (int i = 0; i < 3; ++i) st->t[i] = nxt.t[i];
  if (tr) { for (int i = 0; i < 6; ++i) tr->dx[i] = dx[i]; pose_to_T16(nxt, tr->T_out); }
  float tn = sqrtf(sqn3(dt));
  bool conv = ((double)tn < prm.tol_t) && ((double)wn < prm.tol_r);
  if (it == 0) st->initial_cost = (double)(float)acc[27];
  st->final_cost = (double)(float)acc[27];
  st->num_iterations = it + 1;
  st->iter = it + 1;
  if (conv) { st->done = 1; st->converged = 1; }
  else if (it + 1 >= prm.max_iterations) st->done = 1;
}

template <bool SURFEL>
struct k_icp_gn { static __device__ __forceinline__ void run(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_npts, IcpState* st, IcpParams prm,
                                                 const double* __restrict__ res, const int* __restrict__ slot, const float4* __restrict__ plane,
                                                 double* partial, double* ext_out) { TL_START();
  __shared__ double red[8][28];
  __shared__ float sR[9], sT[3];
  __shared__ int s_last;
  // the done flag, the pose and the loop-invariant scalars go out together (one memory round trip ahead of the work, not three)
  const int done_v = st->done;
  const int npts = *d_npts;
  const double scale_v = st->scale, delta_v = st->delta;
  float pose_v = 0.0f;
  if (threadIdx.x < 12) pose_v = threadIdx.x < 9 ? st->R[threadIdx.x] : st->t[threadIdx.x - 9];
  if (done_v) return;
  // CTAs beyond the last tile of this scan have nothing to add: they leave before the reduction, and the election and the final sum
  // run over the active CTAs only (their partial sums would be exact zeros: the result is bit-identical).  The grid is sized by the
  // buffer capacity (graph-replayable), so for a 4 k-point scan 48 of 64 CTAs are of this kind - harmless for a lone sequence, but a
  // lock-step batch of 128 sequences would push 6000 of them through the SMs per launch.
  // The partial sums are partitioned over VIRTUAL blocks: VG = the grid a lone sequence starts (then virtual = real and the loop below
  // runs once).  A lock-step batch starts fewer CTAs per sequence (its share of the GPU) and each walks several virtual blocks, one
  // reduction each - the same additions in the same order, without pushing 48 idle CTAs per sequence through the SMs.
  const int ntile_v = (npts + (int)blockDim.x - 1) / (int)blockDim.x;
  const int VG = prm.gn_vgrid > 0 ? prm.gn_vgrid : (int)gridDim.x;
  const unsigned nact = (unsigned)(ntile_v < 1 ? 1 : (ntile_v < VG ? ntile_v : VG));                  // virtual blocks with work
  const unsigned nreal = nact < gridDim.x ? nact : gridDim.x;                                          // CTAs taking part
  if (blockIdx.x >= nreal) return;
  if (threadIdx.x < 9) sR[threadIdx.x] = pose_v; else if (threadIdx.x < 12) sT[threadIdx.x - 9] = pose_v;
  __syncthreads();
  const double sdiv = fmax(scale_v, 1e-6);
  const float delta = (float)delta_v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (unsigned vb = blockIdx.x; vb < nact; vb += gridDim.x) {
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  for (int i = (int)vb * blockDim.x + threadIdx.x; i < npts; i += VG * blockDim.x) {
    int s = slot[i];
    if (s < 0) continue;
    float4 P = pts[i];
    float n[3], qc[3];
    if (SURFEL) {
      const float4* e = reinterpret_cast<const float4*>(&M.l1_tab[s]);
      float4 ea = __ldg(e), eb = __ldg(e + 1);
      n[0] = ea.z; n[1] = ea.w; n[2] = eb.x; qc[0] = eb.y; qc[1] = eb.z; qc[2] = eb.w;
    } else {
      float4 ea = plane[2 * i], eb = plane[2 * i + 1];
      n[0] = ea.x; n[1] = ea.y; n[2] = ea.z; qc[0] = ea.w; qc[1] = eb.x; qc[2] = eb.y;
    }
    float p[3] = {P.x, P.y, P.z};
    // p_world = R * p + t ; residual = n . (p_world - q)            (ICP.cpp:368-371)
    float Rp[3]; mat3_vec(sR, p, Rp);
    float d[3] = {(Rp[0] + sT[0]) - qc[0], (Rp[1] + sT[1]) - qc[1], (Rp[2] + sT[2]) - qc[2]};
    float r = dot3(n, d);
    float J[6];
    for (int j = 0; j < 3; ++j) J[j] = add3(n[0] * sR[j], n[1] * sR[3 + j], n[2] * sR[6 + j]);               // n^T R
    float u[3];
    for (int j = 0; j < 3; ++j) u[j] = add3((-n[0]) * sR[j], (-n[1]) * sR[3 + j], (-n[2]) * sR[6 + j]);      // (-n)^T R
    J[3] = add3(u[0] * 0.0f, u[1] * p[2], u[2] * (-p[1]));                                                   // ... * [p]x
    J[4] = add3(u[0] * (-p[2]), u[1] * 0.0f, u[2] * p[0]);
    J[5] = add3(u[0] * p[1], u[1] * (-p[0]), u[2] * 0.0f);
    float w = 1.0f;
    if (prm.use_robust) {  // ICP.cpp:388-404
      float an = fabsf((float)(res[i] / sdiv));
      if (prm.loss_type == 1) { float ratio = an / delta; w = 1.0f / (1.0f + ratio * ratio); }
      else if (an > delta) w = delta / an;
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      float wJ = w * J[a];
#pragma unroll
      for (int b = 0; b <= a; ++b) acc[q++] += (double)(wJ * J[b]);
    }
    float wr = w * r;
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += (double)(wr * J[a]);
    acc[27] += (double)(wr * r);
  }
  // block reduction: warp shuffle, then shared memory across the 8 warps.  The 28 sums are reduced by RECURSIVE HALVING: in the stage
  // with partner lane ^ s a lane keeps the half of its values whose index has bit s equal to its own and hands the other half over,
  // so 16 + 8 + 4 + 2 + 1 = 31 shuffle-adds leave the total of value L on lane L - instead of 28 five-stage butterflies (140).
  // Every total is built by the same tree of pairwise additions as in the butterfly (lane L's view of it; fp addition commutes), so
  // the bits are unchanged.
  {
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < 28 ? acc[k] : 0.0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool upper = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const double send = upper ? v[i] : v[i + s];
        const double keep = upper ? v[i + s] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    if (lane < 28) red[wid][lane] = v[0];
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    double v = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][threadIdx.x];
    partial[vb * 28 + threadIdx.x] = v;
  }
  if (vb + gridDim.x < nact) __syncthreads();   // red is rewritten by the next virtual block
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { unsigned int old = atomicAdd(&st->ticket, 1u); s_last = (old == nreal - 1); }
  __syncthreads();
  if (!s_last) return;
  const long long g0 = clock64();
  __threadfence();
  // fixed-order sum of the per-block partials: 8 warps x 28 lanes take every 8th block (independent L2 loads in flight),
  // then the 8 warp totals are added in order -> bit-reproducible from run to run
  {
    double v = 0.0;
    if (lane < 28) for (unsigned b = wid; b < nact; b += 8) v += __ldcg(&partial[b * 28 + lane]);
    __syncthreads();
    if (lane < 28) red[wid][lane] = v;
    __syncthreads();
    if (threadIdx.x < 28) {
      double t = red[0][threadIdx.x];
      for (int w8 = 1; w8 < 8; ++w8) t += red[w8][threadIdx.x];
      red[0][threadIdx.x] = t;
    }
  }
  __syncthreads();
  if (ext_out) {  // point-sharded mode: hand this rank's 28 partial sums to the all-reduce; b2lo_icp_shard_finish solves
    if (threadIdx.x < 28) ext_out[threadIdx.x] = red[0][threadIdx.x];
    if (threadIdx.x == 0) st->ticket = 0u;
    return;
  }
  if (threadIdx.x == 0) {
    const long long g1 = clock64();
    st->ticket = 0u;
    TL_HERE();   // partials summed, finish begins
    gn_finish(st, prm, red[0]);
    TL_HERE();   // finish done
    st->dbg[16] = g1 - g0; st->dbg[17] = clock64() - g1;
  }
} };

struct k_icp_begin { static __device__ __forceinline__ void run(IcpState* st, const ScanParams* __restrict__ sp) { TL_START();
  if (threadIdx.x == 0 && blockIdx.x == 0) icp_state_begin(st, sp);
} };
// on failure the reference leaves optimized_transform = initial (ICP.cpp:266,301)
struct k_icp_end { static __device__ __forceinline__ void run(IcpState* st) { TL_START();
  if (threadIdx.x == 0 && blockIdx.x == 0 && st->done == 2) {
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) st->R[i * 3 + j] = st->T_init[i * 4 + j]; st->t[i] = st->T_init[i * 4 + 3]; }
  }
} };

// parity tap: per-query correspondence state at a fixed pose
__global__ void k_icp_taps(MapDev M, const float4* __restrict__ pts, int npts, const float* __restrict__ T16, double max_dist, int* state,
                           int* key3, unsigned long long* morton, float* nout, float* cout, double* res) { TL_START();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  float R[9], t[3];
  for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) R[a * 3 + b] = T16[a * 4 + b]; t[a] = T16[a * 4 + 3]; }
  float4 p = pts[i];
  float w[3], n[3] = {0, 0, 0}, c[3] = {0, 0, 0};
  transform_point(R, t, p.x, p.y, p.z, w);
  int s = surfel_probe(M, w, n, c, key3 + i * 3, morton + i);
  int stt = 0; double r = 0.0;
  if (s >= 0) { r = gate_residual(n, c, w); stt = (r > max_dist) ? 1 : 2; }
  state[i] = stt; res[i] = r;
  for (int a = 0; a < 3; ++a) { nout[i * 3 + a] = n[a]; cout[i * 3 + a] = c[a]; }
}

// ---------------------------------------------------------------------------------------------------
// K2 launch geometry for a sequence built for `ctiles_cap` tiles: the kernel is persistent and software-pipelined, so exactly one
// resident wave.  (A cp.async-streamed variant with deeper thread-private shared-memory rings was tried for dense clouds and was
// slower: the 16 B LDGSTS copies saturate the MIO queue - ncu: mio_throttle 9.4 stalls per issue, 44.8 us vs 33.2 us per 2^20 probes.)
#define B2_CORR_ARGS MapDev, const float4*, const int*, IcpState*, IcpParams, double*, int*, int*, int*, double*, int*, const PkoTables*, const int*, double*, const ScanParams*
struct CorrLaunch { bool fuse; int grid; size_t smem; };
static CorrLaunch corr_launch(b2lo_ctx* ctx, int ctiles_cap, bool fuse_pko = false) {
  static int per_sm[2] = {0, 0};   // a property of the compiled kernel, identical on every device of this build
  const int f = fuse_pko ? 1 : 0;
  if (per_sm[f] == 0) {
    // 80 registers, 3 CTAs/SM; capping at 64 registers for 4 CTAs/SM measured the same (37.1 vs 37.7 us)
    int n = 0;
    cudaError_t e = fuse_pko ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_one<k_icp_corr<3, 1, true>, TILE, 1, B2_CORR_ARGS>, TILE, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_one<k_icp_corr<3, 1, false>, TILE, 1, B2_CORR_ARGS>, TILE, 0);
    if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = 2; }
    per_sm[f] = n;
  }
  if (ctiles_cap < 1) ctiles_cap = 1;
  const int resident = ctx->sm_count * per_sm[f];
  CorrLaunch L;
  L.fuse = fuse_pko; L.grid = ctiles_cap < resident ? ctiles_cap : resident; L.smem = 0;
  // lock-step batch: the fused kernel's outputs are per tile (independent of the grid), so a sequence starts only as many CTAs as its
  // share of the GPU; the CTAs walk the tiles.  (A lone sequence keeps one CTA per tile of the buffer capacity: 48 of 64 leave at once,
  // but 128 sequences x 48 idle CTAs took a third of the SM slot time of the batched launch.)
  if (fuse_pko) L.grid = batch_grid(ctx, L.grid);
  return L;
}
// start K2 (the argument types are spelled out so that the instantiation is the one the occupancy query above looked at)
static void corr_start(b2lo_ctx* ctx, const CorrLaunch& cl, cudaStream_t s, MapDev M, const float4* pts, const int* d_npts, IcpState* st, IcpParams prm, double* res,
                       int* slot, int* cidx, int* tilecnt, double* tilesum, int* tileoff, const PkoTables* T, const int* hits, double* gmm, const ScanParams* sp_first) {
  if (cl.fuse) launch<k_icp_corr<3, 1, true>, TILE, 1, 1, B2_CORR_ARGS>(ctx, dim3((unsigned)cl.grid), dim3(TILE), cl.smem, s, M, pts, d_npts, st, prm, res, slot, cidx, tilecnt, tilesum,
                                                                     tileoff, T, hits, gmm, sp_first);
  else launch<k_icp_corr<3, 1, false>, TILE, 1, 1, B2_CORR_ARGS>(ctx, dim3((unsigned)cl.grid), dim3(TILE), cl.smem, s, M, pts, d_npts, st, prm, res, slot, cidx, tilecnt, tilesum,
                                                              tileoff, T, hits, gmm, sp_first);
}

static int knn_reserve(b2lo_ctx* ctx) {
  if (ctx->k_cap >= ctx->pts_cap) return B2LO_OK;
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  void* old[] = {ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_plane};
  for (void* p : old) if (p) cudaFree(p);
  size_t cap = ctx->pts_cap;
  B2_CUDA(cudaMalloc((void**)&ctx->k_idx, cap * KNN_K * sizeof(int)));
  B2_CUDA(cudaMalloc((void**)&ctx->k_n, cap * sizeof(int)));
  B2_CUDA(cudaMalloc((void**)&ctx->k_unres, cap * sizeof(int)));
  B2_CUDA(cudaMalloc((void**)&ctx->k_plane, cap * 2 * sizeof(float4)));
  if (!ctx->k_nunres) { B2_CUDA(cudaMalloc((void**)&ctx->k_nunres, sizeof(int))); B2_CUDA(cudaMemsetAsync(ctx->k_nunres, 0, sizeof(int), ctx->stream)); }
  ctx->k_cap = cap;
  ctx->alloc_epoch++;
  return B2LO_OK;
}

int icp_build_pko(b2lo_ctx* ctx, const b2lo_icp_cfg* cfg) {
  if (cfg->num_alpha_segments < 1 || cfg->num_alpha_segments > 128 || cfg->gmm_sample_size < 1 || cfg->gmm_sample_size > 128 ||
      cfg->gmm_components != 3 || cfg->pko_kernel_type < 0 || cfg->pko_kernel_type > 1) {
    set_error("PKO config outside the supported domain (segments<=128, sample<=128, 3 components, huber|cauchy kernel)");
    return B2LO_E_ARG;
  }
  const b2lo_icp_cfg& o = ctx->pko_cfg_built;
  if (ctx->pko_built && o.num_alpha_segments == cfg->num_alpha_segments && o.gmm_sample_size == cfg->gmm_sample_size &&
      o.pko_kernel_type == cfg->pko_kernel_type && o.min_scale_factor == cfg->min_scale_factor && o.max_scale_factor == cfg->max_scale_factor &&
      o.truncated_threshold == cfg->truncated_threshold)
    return B2LO_OK;
  bool first = ctx->h_pko_hits.empty();
  if (first) pko_build_host(cfg, &ctx->h_pko, &ctx->h_pko_hits);
  else { std::vector<int> tmp; pko_build_host(cfg, &ctx->h_pko, &tmp); }
  if (!ctx->d_pko) B2_CUDA(cudaMalloc(&ctx->d_pko, sizeof(PkoTables)));
  if (!ctx->d_pko_hits) B2_CUDA(cudaMalloc(&ctx->d_pko_hits, sizeof(int) * (ctx->h_pko_hits.size() + 1)));
  B2_CUDA(cudaMemcpyAsync(ctx->d_pko, &ctx->h_pko, sizeof(PkoTables), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(ctx->d_pko_hits, ctx->h_pko_hits.data(), sizeof(int) * ctx->h_pko_hits.size(), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->pko_cfg_built = *cfg;
  ctx->pko_built = true;
  return B2LO_OK;
}

// everything icp_run may have to allocate or upload, done ahead of a stream capture
int icp_prepare(b2lo_ctx* ctx, const b2lo_icp_cfg* cfg) {
  int rc = icp_build_pko(ctx, cfg);
  if (!rc && !cfg->use_surfel_correspondence) rc = knn_reserve(ctx);
  return rc;
}

// enqueue a whole optimize() on the context stream.  T_init16 (host) is copied through the pinned state
// block unless init_pose_on_device (then st->T_init was written by a previous kernel).
int icp_run(b2lo_map* map, const float4* d_pts, const int* d_npts, size_t npts_cap, const float* T_init16, const b2lo_icp_cfg* cfg,
            bool init_pose_on_device, bool restore_on_failure) {
  b2lo_ctx* ctx = map->ctx;
  // any iteration count >= 1 (ICPConfig's default is 50, ICP.h:57); only the per-iteration trace is limited to the first B2LO_MAX_ITERS
  if (cfg->max_iterations < 1) { set_error("max_iterations must be >= 1"); return B2LO_E_ARG; }
  const bool surfel = cfg->use_surfel_correspondence != 0;
  int rc = icp_build_pko(ctx, cfg);
  if (rc) return rc;
  if (!surfel && (rc = knn_reserve(ctx))) return rc;
  cudaStream_t s = ctx->stream;
  IcpParams prm;
  prm.max_iterations = cfg->max_iterations; prm.min_corr = cfg->min_correspondence_points; prm.use_robust = cfg->use_robust_loss;
  prm.loss_type = cfg->loss_type; prm.use_pko = cfg->use_adaptive_m_estimator; prm.use_surfel = cfg->use_surfel_correspondence;
  prm.tol_t = cfg->translation_tolerance; prm.tol_r = cfg->rotation_tolerance; prm.max_dist = cfg->max_correspondence_distance;
  prm.robust_delta = cfg->robust_loss_delta;
  const int qpt = 1;   // measured on the 10^7-voxel map: 1 query/thread 45.5 us, 2: 47.7 us, 4: 52.4 us per 2^20 random probes
  prm.ctile = surfel ? TILE * qpt : TILE;
  (void)init_pose_on_device;
  if (!ctx->sp_preloaded) {
    if ((rc = sp_begin_write(ctx))) return rc;
    for (int i = 0; i < 16; ++i) ctx->h_sp->T_init[i] = T_init16[i];
    ctx->h_sp->force_scale = ctx->force_scale;
    if ((rc = sp_upload(ctx, offsetof(ScanParams, T_init), SP_POSE_BYTES))) return rc;
  }
  ctx->force_scale = 0.0;
  int grid_knn = (int)((npts_cap + 7) / 8);   // one warp per query, 8 per CTA
  if (grid_knn > ctx->sm_count * 8) grid_knn = ctx->sm_count * 8;
  if (grid_knn < 1) grid_knn = 1;
  int ctiles_cap = (int)((npts_cap + prm.ctile - 1) / prm.ctile);
  // the last CTA of the correspondence kernel runs the PKO fit itself; per-kernel profiling keeps the two launches apart
  // (scan-sized clouds only: the fused kernel needs 88 registers, which would cost a dense cloud one resident CTA per SM)
  const bool fuse = surfel && !(ctx->prof && ctx->prof->on) && npts_cap <= 65536;
  const CorrLaunch cl = corr_launch(ctx, ctiles_cap, fuse);
  if (!fuse) { launch<k_icp_begin, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, s, ctx->d_icp, ctx->d_sp); ctx->launches++; }   // fused: the first correspondence pass initialises the state
  cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap_status);
  const bool capturing = cap_status != cudaStreamCaptureStatusNone;
  int ntiles_cap = (int)((npts_cap + TILE - 1) / TILE);
  int grid = ntiles_cap < 1 ? 1 : (ntiles_cap > ctx->i_max_blocks ? ctx->i_max_blocks : ntiles_cap);
  double* gmm = ctx->i_partial + (size_t)ctx->i_max_blocks * 28;      // 9 GMM doubles, then P(r_k) at [16, 116)
  double* js = gmm + 120;                                             // 129 doubles
  unsigned int* tk = reinterpret_cast<unsigned int*>(js + 132);       // PKO arg-min ticket (zeroed at context creation)
  for (int it = 0; it < cfg->max_iterations; ++it) {
    if (surfel) {
      prof_begin(ctx, PS_CORR);
      corr_start(ctx, cl, s, map->d, d_pts, d_npts, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_tilesum,
                                          ctx->i_blkoff, ctx->d_pko, ctx->d_pko_hits, gmm, (fuse && it == 0) ? ctx->d_sp : nullptr);
      prof_end(ctx);
    } else {
      prof_begin(ctx, PS_KNN);
      k_knn_search<<<grid_knn, TILE, 0, s>>>(map->d, d_pts, d_npts, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres);
      k_knn_brute<<<ctx->sm_count * 2, 256, 0, s>>>(map->d, d_pts, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres);
      k_knn_gate<<<grid, TILE, 0, s>>>(map->d, d_pts, d_npts, ctx->d_icp, prm, ctx->k_idx, ctx->k_n, ctx->k_nunres, ctx->i_res, ctx->i_slot,
                                       ctx->i_cidx, ctx->i_blkcnt, ctx->k_plane, ctx->i_tilesum);
      prof_end(ctx);
      ctx->launches += 2;
    }
    if (!fuse) {
      prof_begin(ctx, PS_PKO1);
      launch<k_icp_pko1, PKO_THREADS, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(PKO_THREADS)), 0, s, d_npts, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_blkoff, ctx->d_pko,
                                           ctx->d_pko_hits, gmm, nullptr, 0, 0.0, ctx->i_tilesum, nullptr);
      prof_end(ctx);
    }
    if (cfg->use_adaptive_m_estimator) {
      prof_begin(ctx, PS_PKO2);
      // (a warp-per-candidate variant for batches - four candidates in flight per CTA, no block barrier - measured the same: at 128
      // sequences the 2.6 M f64 logarithms of a step's arg-min are FP64-throughput-bound, 21 us per launch either way)
      launch<k_icp_pko2, 128, 1>(ctx, dim3((unsigned)(batch_grid(ctx, cfg->num_alpha_segments))), dim3((unsigned)(128)), 0, s, ctx->d_icp, prm, ctx->d_pko, gmm, js, tk);
      prof_end(ctx);
    }
    prof_begin(ctx, PS_GN);
    IcpParams pg = prm;
    pg.gn_vgrid = grid;                       // the partition of the partial sums is the lone sequence's, whatever grid is started
    const int gstart = batch_grid(ctx, grid);
    if (surfel) launch<k_icp_gn<true>, TILE, 2>(ctx, dim3((unsigned)(gstart)), dim3((unsigned)(TILE)), 0, s, map->d, d_pts, d_npts, ctx->d_icp, pg, ctx->i_res, ctx->i_slot, nullptr, ctx->i_partial, nullptr);
    else launch<k_icp_gn<false>, TILE, 1>(ctx, dim3((unsigned)(gstart)), dim3((unsigned)(TILE)), 0, s, map->d, d_pts, d_npts, ctx->d_icp, pg, ctx->i_res, ctx->i_slot, ctx->k_plane, ctx->i_partial, nullptr);
    prof_end(ctx);
    ctx->launches += (cfg->use_adaptive_m_estimator ? 4 : 3) - (fuse ? 1 : 0);
    // long runs (max_iterations beyond the usual 4): look at the done flag every 8 iterations instead of enqueueing dozens of no-op launches
    if (!capturing && (it & 7) == 7 && it + 1 < cfg->max_iterations) {
      B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 48, &ctx->d_icp->done, sizeof(int), cudaMemcpyDeviceToHost, s));
      B2_CUDA(cudaStreamSynchronize(s));
      if (ctx->h_counts[48]) break;
    }
  }
  // the per-scan driver's decision kernel never reads the pose of a failed optimize (it falls back to the motion-model guess itself)
  if (restore_on_failure) { launch<k_icp_end, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, s, ctx->d_icp); ctx->launches++; }
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}

// ---- point-sharded mode: queries split across ranks, map replicated ---------------------------------------
// local statistics of this rank's shard after K2: tile offsets, accepted count, sum r, sum r^2
__device__ __forceinline__ void shard_stats_body(const int* __restrict__ d_npts, IcpState* st, const IcpParams& prm, const int* __restrict__ tilecnt, int* tileoff,
                                                 double* stats3, const double* __restrict__ tilesum) {
  __shared__ int sm[40];
  __shared__ double smd[40];
  const int tid = threadIdx.x;
  const int npts = *d_npts;
  const int ntiles = (npts + prm.ctile - 1) / prm.ctile;
  const int base = tile_scan_warps(tilecnt, tileoff, ntiles, sm);
  double a1 = 0.0, a2 = 0.0;
  if (st->iter == 0) {   // the moments only feed the iteration-0 residual scale (shard_plan_body)
    for (int t = tid; t < ntiles; t += blockDim.x) { a1 += tilesum[2 * t]; a2 += tilesum[2 * t + 1]; }
    a1 = block_sum_d(a1, smd);
    a2 = block_sum_d(a2, smd);
  }
  if (tid == 0) { st->n_blocks = base; stats3[0] = (double)base; stats3[1] = a1; stats3[2] = a2; }
}
__global__ void __launch_bounds__(256) k_shard_stats(const int* __restrict__ d_npts, IcpState* st, IcpParams prm, const double* __restrict__ res,
                                                      const int* __restrict__ slot, const int* __restrict__ tilecnt, int* tileoff, double* stats3,
                                                      const double* __restrict__ tilesum) { TL_START();
  shard_stats_body(d_npts, st, prm, tilecnt, tileoff, stats3, tilesum);
}
// this rank's contribution to the global GMM sample: position j of shuffle(iota(C_total)) is a global compacted index;
// the rank owning it ([offset, offset + C_local)) writes the normalised residual, everybody else writes 0
// (called by every thread of the block; threads >= MAXS only take part in the barrier)
__device__ __forceinline__ void shard_sample_body(const int* __restrict__ d_npts, IcpState* st, const IcpParams& prm, const double* __restrict__ res,
                                                  const int* __restrict__ cidx, const int* __restrict__ tileoff, const PkoTables* __restrict__ T,
                                                  const int* __restrict__ hits, long long offset, long long c_total, double scale, double* sample,
                                                  const double* plan) {
  __shared__ int s_head[MAXS];
  const int tid = threadIdx.x;
  if (plan) {   // device-ordered mode: offset / total / scale come from k_shard_plan; a finished or failed optimize contributes nothing
    offset = (long long)plan[0]; c_total = (long long)plan[1]; scale = plan[2];
    if (st->done || c_total < 1) { if (tid < MAXS) sample[tid] = 0.0; return; }
  }
  const int C = (int)c_total;
  const int c_local = st->n_blocks;  // accepted count of this shard (k_shard_stats)
  const int npts = *d_npts;
  const int ntiles = (npts + prm.ctile - 1) / prm.ctile;
  const int mode = C >= 65536 ? 2 : ((C & 1) ? 1 : 0);
  if (tid < MAXS) s_head[tid] = T->head_r[mode][tid];
  __syncthreads();
  const int ns = T->sample_size < C ? T->sample_size : C;
  double out = 0.0;
  if (tid < ns) {
    const int lo = T->hit_off[mode][tid], hi = T->hit_off[mode][tid + 1];
    int best = -1;
    for (int q0 = lo; q0 < hi; q0 += 8) {   // ascending hit list, eight independent loads per round (as in pko1_body)
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (q0 + u < hi) ? hits[q0 + u] : 0x7fffffff;
#pragma unroll
      for (int u = 0; u < 8; ++u) if (v[u] < C) best = v[u];
      if (v[7] >= C) break;
    }
    long long ci = best;
    if (best < 0) {
      if (C >= MAXS) ci = T->head_pos[mode][tid];   // all 127 head swaps apply: the trace is a constant of the mode
      else {
        int pos = tid;
        for (int i = C - 1; i >= 1; --i) { int r = s_head[i]; pos = (pos == i) ? r : ((pos == r) ? i : pos); }
        ci = pos;
      }
    }
    ci -= offset;
    if (ci >= 0 && ci < c_local) {
      const int tl = tile_of(tileoff, ntiles, (int)ci);
      int q = cidx[tl * prm.ctile + ((int)ci - tileoff[tl])];
      out = res[q] / fmax(scale, 1e-6);
    }
  }
  if (tid < MAXS) sample[tid] = out;
}
__global__ void __launch_bounds__(MAXS) k_shard_sample(const int* __restrict__ d_npts, IcpState* st, IcpParams prm, const double* __restrict__ res,
                                                        const int* __restrict__ cidx, const int* __restrict__ tileoff, const PkoTables* __restrict__ T,
                                                        const int* __restrict__ hits, long long offset, long long c_total, double scale, double* sample,
                                                        const double* __restrict__ plan) { TL_START();
  shard_sample_body(d_npts, st, prm, res, cidx, tileoff, T, hits, offset, c_total, scale, sample, plan);
}
__global__ void k_shard_finish(IcpState* st, IcpParams prm, const double* __restrict__ acc28) { TL_START();
  if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) {
    double acc[28];
    for (int i = 0; i < 28; ++i) acc[i] = acc28[i];
    gn_finish(st, prm, acc);
  }
}

// device-ordered point-sharded mode: every rank turns the all-gathered (C_r, sum r, sum r^2) triples into the same plan -
// plan[0] = global offset of this rank's accepted correspondences, plan[1] = global count C, plan[2] = residual scale - and applies
// the reference's C < min_correspondence_points test (ICP.cpp:298-302) and iteration-0 scale (ICP.cpp:304-316) to the GLOBAL values
__device__ __forceinline__ void shard_plan_body(IcpState* st, const IcpParams& prm, const double* gathered, int world, int rank, double* plan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (st->done) { plan[0] = 0.0; plan[1] = 0.0; plan[2] = st->scale; return; }
  long long off = 0, total = 0;
  double a1 = 0.0, a2 = 0.0;
  for (int r = 0; r < world; ++r) {
    const long long c = (long long)gathered[3 * r];
    if (r < rank) off += c;
    total += c; a1 += gathered[3 * r + 1]; a2 += gathered[3 * r + 2];
  }
  double scale = st->scale;
  st->n_corr = (int)total;
  if (total < prm.min_corr) { st->done = 2; st->status = B2LO_S_INSUFFICIENT; }
  else if (prm.use_pko && total > (1ll << 22)) { st->done = 2; st->status = B2LO_E_CAPACITY; }   // the GLOBAL count left the PKO sample tables
  else if (st->iter == 0 && !st->scale_forced) {
    const double mean = a1 / (double)total;
    scale = sqrt(fmax(a2 / (double)total - mean * mean, 0.0)) / 6.0;
    st->scale = scale;
  }
  plan[0] = (double)off; plan[1] = (double)total; plan[2] = scale;
}
__global__ void k_shard_plan(IcpState* st, IcpParams prm, const double* __restrict__ gathered, int world, int rank, double* plan) { TL_START();
  shard_plan_body(st, prm, gathered, world, rank, plan);
}

// ---- peer-memory exchange over NVLink (the B200-native alternative to the three NCCL launches per iteration) -------------------------
// Every rank owns a MAILBOX in its HBM that the peers write into directly (cudaIpc-mapped pointers, one process per GPU): slot
// [exchange][source rank] holds the source's payload tagged with the epoch of the exchange.  One exchange: store the own payload
// into every peer's mailbox, then wait for the current epoch from all sources in the OWN mailbox and add / gather the payloads in
// rank order (the same order on every rank: bit-identical results everywhere).  No NCCL launch, no proxy thread, no ring.  A slot is
// reused every third exchange at the earliest, and every exchange in between is a barrier of all ranks, so no payload is overwritten
// before it is read.
// Because an exchange is a few lines inside a CTA, the single-CTA steps around it fuse: per iteration the peer path launches
//   K2 -> k_shard_pre_peer (shard statistics, all-gather, plan, sample, all-reduce, GMM fit) -> arg-min -> K5 -> k_shard_post_peer
// (all-reduce of the 28 sums + the 6x6 solve) = 5 kernels instead of the 11 of the NCCL path.
constexpr int P2P_MAXW = 8, P2P_MAXN = 128;
// A payload double travels as two 8-byte words {32 data bits, 32-bit epoch}: an aligned 8-byte store arrives whole, so the reader
// needs no separate flag and the writer no fence between payload and flag (the "low-latency" wire format NCCL also uses for small
// messages) - one NVLink store latency per exchange.
struct ShardMailbox { unsigned long long w[4][P2P_MAXW][2 * P2P_MAXN]; };   // exchange kinds 0..2 of an iteration + 3 = the start line of an optimize
struct PeerTable { ShardMailbox* p[P2P_MAXW]; };
struct PeerArgs { PeerTable peers; int rank, world; unsigned long long epoch; int* err; };
// mode 0: dst[r * n + t] = payload of rank r (all-gather); mode 1: dst[t] = sum over ranks in rank order (all-reduce).
// Called by every thread of a block of >= P2P_MAXN threads; src/dst are read/written by this block only.
__device__ __forceinline__ void p2p_exchange_body(const PeerArgs& pa, const double* src, double* dst, int n, int e, unsigned long long epoch, int mode) {
  const int t = threadIdx.x;
  __syncthreads();                   // src was written by other threads of this block
  unsigned long long g0 = 0;
  if (t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const unsigned long long tag = (epoch & 0xffffffffull) << 32;
  if (t < n) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(src[t]);
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    for (int p = 0; p < pa.world; ++p) {
      volatile unsigned long long* q = pa.peers.p[p]->w[e][pa.rank];
      q[2 * t] = w0; q[2 * t + 1] = w1;
    }
    const volatile unsigned long long* mine = pa.peers.p[pa.rank]->w[e][0];
    const long long t0 = clock64();
    double a = 0.0;
    for (int r = 0; r < pa.world; ++r) {   // rank order: the same sum on every rank
      unsigned long long v0, v1;
      for (;;) {                           // bounded: a lost peer must not hang the GPU
        v0 = mine[(size_t)r * 2 * P2P_MAXN + 2 * t]; v1 = mine[(size_t)r * 2 * P2P_MAXN + 2 * t + 1];
        if ((v0 >> 32) == (tag >> 32) && (v1 >> 32) == (tag >> 32)) break;
        if (clock64() - t0 > 6000000000ll) { atomicExch(pa.err, 1); break; }
      }
      const double x = __longlong_as_double((long long)((v0 & 0xffffffffull) | (v1 << 32)));
      if (mode == 0) dst[r * n + t] = x; else a += x;
    }
    if (mode != 0) dst[t] = a;
  }
  __syncthreads();
  if (t == 0) {                      // latency report: ns this exchange took on this rank, waiting for the slowest peer included
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (e < 3) reinterpret_cast<unsigned long long*>(pa.err + 2)[e] = g1 - g0;
  }
}
__global__ void __launch_bounds__(P2P_MAXN) k_p2p_exchange(PeerArgs pa, const double* src, double* dst, int n, int e, int mode) {
  p2p_exchange_body(pa, src, dst, n, e, pa.epoch, mode);
}
// everything between K2 and the arg-min of one point-sharded iteration in ONE single-CTA kernel (epochs pa.epoch and pa.epoch + 1)
__global__ void __launch_bounds__(PKO_THREADS) k_shard_pre_peer(PeerArgs pa, const int* d_npts, IcpState* st, IcpParams prm, const double* res, const int* cidx,
                                                                const int* tilecnt, int* tileoff, const double* tilesum, const PkoTables* T, const int* hits,
                                                                double* stats3, double* gathered, double* plan, double* sample, double* gmm_out) { TL_START();
  shard_stats_body(d_npts, st, prm, tilecnt, tileoff, stats3, tilesum);
  if (threadIdx.x == 0) TL_HERE();   // shard statistics done
  p2p_exchange_body(pa, stats3, gathered, 3, 0, pa.epoch, 0);
  if (threadIdx.x == 0) TL_HERE();   // statistics gathered
  shard_plan_body(st, prm, gathered, pa.world, pa.rank, plan);
  __syncthreads();
  shard_sample_body(d_npts, st, prm, res, cidx, tileoff, T, hits, 0, 0, 0.0, sample, plan);
  if (threadIdx.x == 0) TL_HERE();   // sample share drawn
  p2p_exchange_body(pa, sample, sample, MAXS, 1, pa.epoch + 1, 1);
  pko1_body(d_npts, st, prm, res, cidx, tilecnt, tileoff, T, hits, gmm_out, sample, 0, 0.0, tilesum, plan);
}
// all-reduce of this rank's 28 Gauss-Newton sums + the 6x6 solve and pose update (every rank: identical bits)
__global__ void __launch_bounds__(P2P_MAXN) k_shard_post_peer(PeerArgs pa, IcpState* st, IcpParams prm, double* acc28) { TL_START();
  p2p_exchange_body(pa, acc28, acc28, 28, 2, pa.epoch, 1);
  if (threadIdx.x == 0 && !st->done) {
    double acc[28];
    for (int i = 0; i < 28; ++i) acc[i] = acc28[i];
    gn_finish(st, prm, acc);
  }
}

__global__ void k_lookup(MapDev M, float px, float py, float pz, float* out7) { TL_START();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float w[3] = {px, py, pz}, n[3] = {0, 0, 0}, c[3] = {0, 0, 0};
  int s = surfel_probe(M, w, n, c, nullptr, nullptr);
  out7[0] = s >= 0 ? 1.0f : 0.0f;
  for (int a = 0; a < 3; ++a) { out7[1 + a] = n[a]; out7[4 + a] = c[a]; }
}

}  // namespace b2

using namespace b2;

extern "C" int b2lo_icp_correspondences(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T16[16],
                                        double max_distance, int* state, int* l1key, uint64_t* morton, float* normal, float* centroid,
                                        double* residual, size_t* n_accepted) {
  if (!map || !T16 || !state || !l1key || !morton || !normal || !centroid || !residual) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  if (n_accepted) *n_accepted = 0;
  if (!local_xyz || m == 0) return B2LO_S_EMPTY;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  { int rr = ctx_reserve_points(ctx, m); if (rr) return rr; }  // may reallocate d_query: reserve before taking the pointer
  int rc = ctx_stage_h2d(ctx, local_xyz, m, stride_floats, 1, ctx->d_query, ctx->d_nquery);
  if (rc) return rc;
  int* d_state = nullptr; int* d_key = nullptr; unsigned long long* d_mor = nullptr; float* d_n = nullptr; float* d_c = nullptr; double* d_r = nullptr;
  float* d_T = nullptr;
  B2_CUDA(cudaMalloc(&d_state, m * sizeof(int)));
  B2_CUDA(cudaMalloc(&d_key, m * 3 * sizeof(int)));
  B2_CUDA(cudaMalloc(&d_mor, m * sizeof(unsigned long long)));
  B2_CUDA(cudaMalloc(&d_n, m * 3 * sizeof(float)));
  B2_CUDA(cudaMalloc(&d_c, m * 3 * sizeof(float)));
  B2_CUDA(cudaMalloc(&d_r, m * sizeof(double)));
  B2_CUDA(cudaMalloc(&d_T, 16 * sizeof(float)));
  B2_CUDA(cudaMemcpyAsync(d_T, T16, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_icp_taps<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(map->d, ctx->d_query, (int)m, d_T, max_distance, d_state, d_key, d_mor, d_n, d_c, d_r);
  ctx->launches++;
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  B2_CUDA(cudaMemcpy(state, d_state, m * sizeof(int), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(l1key, d_key, m * 3 * sizeof(int), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(morton, d_mor, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(normal, d_n, m * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(centroid, d_c, m * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(residual, d_r, m * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d_state); cudaFree(d_key); cudaFree(d_mor); cudaFree(d_n); cudaFree(d_c); cudaFree(d_r); cudaFree(d_T);
  if (n_accepted) { size_t c = 0; for (size_t i = 0; i < m; ++i) c += (state[i] == 2); *n_accepted = c; }
  return B2LO_OK;
}

extern "C" int b2lo_map_lookup(b2lo_map* map, const float p[3], float n[3], float c[3]) {
  if (!map || !p || !n || !c) return B2LO_E_ARG;
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  cudaSetDevice(ctx->device);
  float* d_out = reinterpret_cast<float*>(ctx->i_partial + (size_t)ctx->i_max_blocks * 28 + 260);
  k_lookup<<<1, 32, 0, ctx->stream>>>(map->d, p[0], p[1], p[2], d_out);
  ctx->launches++;
  float h[8];
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 24, d_out, 7 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(h, ctx->h_counts + 24, 7 * sizeof(float));
  if (h[0] == 0.0f) return 0;
  n[0] = h[1]; n[1] = h[2]; n[2] = h[3]; c[0] = h[4]; c[1] = h[5]; c[2] = h[6];
  return 1;
}


// KDTree-mode parity tap: exact 5-NN (indices into the b2lo_map_export_l0 order), found count, plane, gate
extern "C" int b2lo_icp_correspondences_knn(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T16[16],
                                            double max_distance, int* knn, float* d2, int* found, int* state, float* normal, float* centroid,
                                            double* residual, size_t* n_accepted, size_t* n_scanned) {
  if (!map || !T16 || !knn || !d2 || !found || !state || !normal || !centroid || !residual) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  if (n_accepted) *n_accepted = 0;
  if (n_scanned) *n_scanned = 0;
  if (!local_xyz || m == 0) return B2LO_S_EMPTY;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  { int rr = ctx_reserve_points(ctx, m); if (rr) return rr; }  // may reallocate d_query: reserve before taking the pointer
  int rc = ctx_stage_h2d(ctx, local_xyz, m, stride_floats, 1, ctx->d_query, ctx->d_nquery);
  if (rc) return rc;
  if ((rc = knn_reserve(ctx))) return rc;
  cudaStream_t s = ctx->stream;
  if ((rc = sp_begin_write(ctx))) return rc;
  for (int i = 0; i < 16; ++i) ctx->h_sp->T_init[i] = T16[i];
  ctx->h_sp->force_scale = 0.0;
  if ((rc = sp_upload(ctx, offsetof(ScanParams, T_init), SP_POSE_BYTES))) return rc;
  launch<k_icp_begin, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, s, ctx->d_icp, ctx->d_sp);
  int grid = (int)((m + 7) / 8);
  if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  k_knn_search<<<grid, TILE, 0, s>>>(map->d, ctx->d_query, ctx->d_nquery, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres);
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 40, ctx->k_nunres, sizeof(int), cudaMemcpyDeviceToHost, s));
  k_knn_brute<<<ctx->sm_count * 2, 256, 0, s>>>(map->d, ctx->d_query, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres);
  B2_CUDA(cudaMemsetAsync(ctx->k_nunres, 0, sizeof(int), s));
  int *d_idx = nullptr, *d_found = nullptr, *d_state = nullptr; float *d_d2 = nullptr, *d_n = nullptr, *d_c = nullptr; double* d_r = nullptr;
  B2_CUDA(cudaMalloc(&d_idx, m * KNN_K * sizeof(int)));
  B2_CUDA(cudaMalloc(&d_d2, m * KNN_K * sizeof(float)));
  B2_CUDA(cudaMalloc(&d_found, m * sizeof(int)));
  B2_CUDA(cudaMalloc(&d_state, m * sizeof(int)));
  B2_CUDA(cudaMalloc(&d_n, m * 3 * sizeof(float)));
  B2_CUDA(cudaMalloc(&d_c, m * 3 * sizeof(float)));
  B2_CUDA(cudaMalloc(&d_r, m * sizeof(double)));
  k_knn_taps<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(map->d, ctx->d_query, (int)m, ctx->d_icp, max_distance, ctx->k_idx, ctx->k_n, d_idx, d_d2, d_found,
                                                         d_state, d_n, d_c, d_r);
  ctx->launches += 4;
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaStreamSynchronize(s));
  B2_CUDA(cudaMemcpy(knn, d_idx, m * KNN_K * sizeof(int), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(d2, d_d2, m * KNN_K * sizeof(float), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(found, d_found, m * sizeof(int), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(state, d_state, m * sizeof(int), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(normal, d_n, m * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(centroid, d_c, m * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(residual, d_r, m * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d_idx); cudaFree(d_d2); cudaFree(d_found); cudaFree(d_state); cudaFree(d_n); cudaFree(d_c); cudaFree(d_r);
  if (n_accepted) { size_t c = 0; for (size_t i = 0; i < m; ++i) c += (state[i] == 2); *n_accepted = c; }
  if (n_scanned) *n_scanned = (size_t)ctx->h_counts[40];
  return B2LO_OK;
}

// ---- point-sharded scan-to-map ICP: per-phase entry points, the collectives run between them on the caller's side -------------
static void shard_params(const b2lo_icp_cfg* cfg, size_t npts_cap, IcpParams& prm, int& qpt) {
  prm.max_iterations = cfg->max_iterations; prm.min_corr = cfg->min_correspondence_points; prm.use_robust = cfg->use_robust_loss;
  prm.loss_type = cfg->loss_type; prm.use_pko = cfg->use_adaptive_m_estimator; prm.use_surfel = 1;
  prm.tol_t = cfg->translation_tolerance; prm.tol_r = cfg->rotation_tolerance; prm.max_dist = cfg->max_correspondence_distance;
  prm.robust_delta = cfg->robust_loss_delta;
  qpt = 1;
  (void)npts_cap;
  prm.ctile = TILE * qpt;
}
static int shard_check(b2lo_map* map, const b2lo_icp_cfg* cfg) {
  if (!map || !cfg) return B2LO_E_ARG;
  if (!cfg->use_surfel_correspondence) { set_error("the point-sharded mode supports surfel correspondence only"); return B2LO_E_ARG; }
  // any iteration count >= 1 (ICPConfig's default is 50, ICP.h:57); only the per-iteration trace is limited to the first B2LO_MAX_ITERS
  if (cfg->max_iterations < 1) { set_error("max_iterations must be >= 1"); return B2LO_E_ARG; }
  return B2LO_OK;
}
extern "C" int b2lo_icp_shard_begin(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T_init[16], const b2lo_icp_cfg* cfg) {
  int rc = shard_check(map, cfg);
  if (rc) return rc;
  if (!T_init || stride_floats < 3) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  if ((rc = icp_build_pko(ctx, cfg))) return rc;
  { int rr = ctx_reserve_points(ctx, m); if (rr) return rr; }  // may reallocate d_query: reserve before taking the pointer
  if ((rc = ctx_stage_h2d(ctx, local_xyz, m, stride_floats, 1, ctx->d_query, ctx->d_nquery))) return rc;
  ctx->shard_m = m;
  if ((rc = sp_begin_write(ctx))) return rc;
  for (int i = 0; i < 16; ++i) ctx->h_sp->T_init[i] = T_init[i];
  ctx->h_sp->force_scale = 0.0;
  if ((rc = sp_upload(ctx, offsetof(ScanParams, T_init), SP_POSE_BYTES))) return rc;
  launch<k_icp_begin, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, ctx->stream, ctx->d_icp, ctx->d_sp);
  ctx->launches++;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}
extern "C" int b2lo_icp_shard_corr(b2lo_map* map, const b2lo_icp_cfg* cfg, double* stats3_dev) {
  int rc = shard_check(map, cfg);
  if (rc) return rc;
  if (!stats3_dev) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  cudaSetDevice(ctx->device);
  IcpParams prm; int qpt;
  const size_t m = ctx->shard_m ? ctx->shard_m : 1;
  shard_params(cfg, m, prm, qpt);
  int ctiles = (int)((m + prm.ctile - 1) / prm.ctile);
  const CorrLaunch cl = corr_launch(ctx, ctiles);
  cudaStream_t s = ctx->stream;
  prof_begin(ctx, PS_CORR);
  corr_start(ctx, cl, s, map->d, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_tilesum,
                                      nullptr, nullptr, nullptr, nullptr, nullptr);
  prof_end(ctx);
  k_shard_stats<<<1, 256, 0, s>>>(ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_blkcnt, ctx->i_blkoff, stats3_dev, ctx->i_tilesum);
  ctx->launches += 2;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}
extern "C" int b2lo_icp_shard_sample(b2lo_map* map, const b2lo_icp_cfg* cfg, long long offset, long long c_total, double scale, double* sample_dev) {
  int rc = shard_check(map, cfg);
  if (rc) return rc;
  if (!sample_dev || c_total < 1 || c_total > (1ll << 22)) { set_error("shard_sample: total correspondence count outside [1, 2^22]"); return B2LO_E_ARG; }
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  cudaSetDevice(ctx->device);
  IcpParams prm; int qpt;
  shard_params(cfg, ctx->shard_m ? ctx->shard_m : 1, prm, qpt);
  k_shard_sample<<<1, MAXS, 0, ctx->stream>>>(ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_cidx, ctx->i_blkoff, ctx->d_pko, ctx->d_pko_hits, offset, c_total,
                                              scale, sample_dev, nullptr);
  ctx->launches++;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}
extern "C" int b2lo_icp_shard_accumulate(b2lo_map* map, const b2lo_icp_cfg* cfg, long long c_total, double scale, const double* sample_dev, double* acc28_dev) {
  int rc = shard_check(map, cfg);
  if (rc) return rc;
  if (!sample_dev || !acc28_dev) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  cudaSetDevice(ctx->device);
  IcpParams prm; int qpt;
  const size_t m = ctx->shard_m ? ctx->shard_m : 1;
  shard_params(cfg, m, prm, qpt);
  cudaStream_t s = ctx->stream;
  double* gmm = ctx->i_partial + (size_t)ctx->i_max_blocks * 28;
  double* js = gmm + 120;
  unsigned int* tk = reinterpret_cast<unsigned int*>(js + 132);
  int ntiles = (int)((m + TILE - 1) / TILE);
  int grid = ntiles < 1 ? 1 : (ntiles > ctx->i_max_blocks ? ctx->i_max_blocks : ntiles);
  prof_begin(ctx, PS_PKO1);
  launch<k_icp_pko1, PKO_THREADS, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(PKO_THREADS)), 0, s, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_blkoff, ctx->d_pko,
                                       ctx->d_pko_hits, gmm, sample_dev, (int)c_total, scale, ctx->i_tilesum, nullptr);
  prof_end(ctx);
  if (cfg->use_adaptive_m_estimator) { launch<k_icp_pko2, 128, 1>(ctx, dim3((unsigned)(cfg->num_alpha_segments)), dim3((unsigned)(128)), 0, s, ctx->d_icp, prm, ctx->d_pko, gmm, js, tk); ctx->launches++; }
  prof_begin(ctx, PS_GN);
  launch<k_icp_gn<true>, TILE, 2>(ctx, dim3((unsigned)(grid)), dim3((unsigned)(TILE)), 0, s, map->d, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, nullptr, ctx->i_partial, acc28_dev);
  prof_end(ctx);
  ctx->launches += 2;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}
extern "C" int b2lo_icp_shard_finish(b2lo_map* map, const b2lo_icp_cfg* cfg, const double* acc28_dev, float T_out[16], int* done, b2lo_icp_stats* stats) {
  int rc = shard_check(map, cfg);
  if (rc) return rc;
  if (!acc28_dev || !T_out || !done) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  cudaSetDevice(ctx->device);
  IcpParams prm; int qpt;
  shard_params(cfg, ctx->shard_m ? ctx->shard_m : 1, prm, qpt);
  k_shard_finish<<<1, 32, 0, ctx->stream>>>(ctx->d_icp, prm, acc28_dev);
  ctx->launches++;
  B2_CUDA(cudaMemcpyAsync(ctx->h_icp, ctx->d_icp, sizeof(IcpState), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->d2h_bytes += sizeof(IcpState);
  const IcpState* h = ctx->h_icp;
  Pose p;
  for (int i = 0; i < 9; ++i) p.R.m[i] = h->R[i];
  for (int i = 0; i < 3; ++i) p.t[i] = h->t[i];
  pose_to_T16(p, T_out);
  *done = h->done;
  if (stats) {
    stats->status = h->status; stats->num_iterations = h->num_iterations; stats->num_correspondences = h->n_corr; stats->converged = h->converged;
    stats->initial_cost = h->initial_cost; stats->final_cost = h->final_cost; stats->device_ms = 0.0f;
    std::memcpy(stats->it, h->trace, sizeof(stats->it));
  }
  return B2LO_OK;
}

// ---- device-ordered point-sharded optimize: the three exchanges of a Gauss-Newton iteration are NCCL calls
// enqueued on the context stream between the kernels - no host synchronisation inside the loop, one at the end.
//   K2 -> k_shard_stats -> ncclAllGather(3 doubles / rank) -> k_shard_plan -> k_shard_sample -> ncclAllReduce(128 doubles)
//      -> PKO fit + arg-min (replicated: every rank fits the identical global sample) -> K5 partial sums -> ncclAllReduce(28 doubles)
//      -> k_shard_finish (every rank solves the identical 6x6 system and keeps the identical pose)
// NCCL is taken from the process (torch has it loaded) with dlopen: the library itself carries no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>
namespace b2 {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) { a.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
  if (!a.h) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
  a.AllGather = (decltype(a.AllGather))dlsym(a.h, "ncclAllGather");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.GetErrorString;
  return a;
}
}  // namespace b2
struct b2lo_shard_comm {
  b2lo_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  // peer-memory exchange (b2lo_shard_comm_ipc_handle / _open_peers)
  b2::ShardMailbox* d_mail = nullptr; b2::PeerTable peers{}; bool p2p = false; unsigned long long epoch = 0; int* d_err = nullptr;
  b2::PeerArgs peer_args(unsigned long long first_epoch) const { b2::PeerArgs a; a.peers = peers; a.rank = rank; a.world = world; a.epoch = first_epoch; a.err = d_err; return a; }
  void* opened[b2::P2P_MAXW] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* d_buf = nullptr;   // [0,3) this rank's stats | [8, 8 + 3 world) gathered | plan (3) | sample (128) | acc (28)
  double *d_stats = nullptr, *d_gathered = nullptr, *d_plan = nullptr, *d_sample = nullptr, *d_acc = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
#define B2_NCCL(expr)                                                                                          \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != ncclSuccess) { b2::set_error("%s failed: %s", #expr, b2::nccl_api().GetErrorString(_r)); return B2LO_E_CUDA; } \
  } while (0)

extern "C" int b2lo_shard_unique_id(void* out, size_t bytes) {
  NcclApi& N = nccl_api();
  if (!N.ok) { set_error("NCCL is not available in this process (dlopen libnccl.so.2 failed)"); return B2LO_E_CUDA; }
  if (!out || bytes < sizeof(ncclUniqueId)) return B2LO_E_ARG;
  ncclUniqueId id;
  B2_NCCL(N.GetUniqueId(&id));
  std::memcpy(out, &id, sizeof id);
  return B2LO_OK;
}
extern "C" int b2lo_shard_comm_create(b2lo_ctx* ctx, int world, int rank, const void* unique_id, size_t bytes, b2lo_shard_comm** out) {
  if (!ctx || !out || world < 1 || rank < 0 || rank >= world) return B2LO_E_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  b2lo_shard_comm* c = new b2lo_shard_comm();
  c->ctx = ctx; c->world = world; c->rank = rank;
  if (world > P2P_MAXW) { delete c; set_error("shard communicator: at most %d ranks", P2P_MAXW); return B2LO_E_ARG; }
  if (world > 1 && unique_id) {   // NCCL is optional: a communicator that only uses the peer-memory exchange passes unique_id = NULL
    NcclApi& N = nccl_api();
    if (!N.ok) { delete c; set_error("NCCL is not available in this process (dlopen libnccl.so.2 failed)"); return B2LO_E_CUDA; }
    if (bytes < sizeof(ncclUniqueId)) { delete c; return B2LO_E_ARG; }
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    ncclResult_t r = N.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", N.GetErrorString(r)); delete c; return B2LO_E_CUDA; }
  }
  const size_t n = 8 + 3 * (size_t)world + 8 + 128 + 32;
  if (cudaMalloc(&c->d_buf, n * sizeof(double)) != cudaSuccess) { set_error("cudaMalloc(shard exchange buffer) failed"); delete c; return B2LO_E_NOMEM; }
  cudaMemset(c->d_buf, 0, n * sizeof(double));
  c->d_stats = c->d_buf; c->d_gathered = c->d_buf + 8; c->d_plan = c->d_gathered + 3 * world; c->d_sample = c->d_plan + 8; c->d_acc = c->d_sample + 128;
  for (auto& e : c->ev) cudaEventCreate(&e);
  if (cudaMalloc((void**)&c->d_mail, sizeof(ShardMailbox)) != cudaSuccess || cudaMalloc((void**)&c->d_err, 8 * sizeof(int)) != cudaSuccess) {
    set_error("cudaMalloc(shard mailbox) failed"); delete c; return B2LO_E_NOMEM;
  }
  cudaMemset(c->d_mail, 0, sizeof(ShardMailbox));
  cudaMemset(c->d_err, 0, 8 * sizeof(int));
  *out = c;
  return B2LO_OK;
}
// the cudaIpc handle of this rank's mailbox (64 bytes), to be handed to every peer by the host
extern "C" int b2lo_shard_comm_ipc_handle(b2lo_shard_comm* c, void* out, size_t bytes) {
  if (!c || !out || bytes < sizeof(cudaIpcMemHandle_t)) return B2LO_E_ARG;
  cudaSetDevice(c->ctx->device);
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, c->d_mail));
  std::memcpy(out, &h, sizeof h);
  return B2LO_OK;
}
// handles: world x bytes_each, in rank order (this rank's own entry is ignored).  After this call the exchanges of
// b2lo_icp_shard_optimize go through the peers' mailboxes over NVLink instead of NCCL.
extern "C" int b2lo_shard_comm_open_peers(b2lo_shard_comm* c, const void* handles, size_t bytes_each) {
  if (!c || !handles || bytes_each < sizeof(cudaIpcMemHandle_t)) return B2LO_E_ARG;
  cudaSetDevice(c->ctx->device);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peers.p[r] = c->d_mail; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * bytes_each, sizeof h);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e)); cudaGetLastError(); return B2LO_E_CUDA; }
    c->opened[r] = ptr;
    c->peers.p[r] = static_cast<ShardMailbox*>(ptr);
  }
  c->p2p = true;
  return B2LO_OK;
}
extern "C" int b2lo_shard_comm_destroy(b2lo_shard_comm* c) {
  if (!c) return B2LO_E_ARG;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->comm) nccl_api().CommDestroy(c->comm);
  for (void* p : c->opened) if (p) cudaIpcCloseMemHandle(p);
  if (c->d_mail) cudaFree(c->d_mail);
  if (c->d_err) cudaFree(c->d_err);
  if (c->d_buf) cudaFree(c->d_buf);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  delete c;
  return B2LO_OK;
}
// optimize() on this rank's slice of a dense scan; every rank receives the same pose.  collective_ms (nullable): CUDA-event time of the
// exchanges of the LAST iteration issued (the three collectives plus the plan / sample kernels between them), for the latency report.
extern "C" int b2lo_icp_shard_optimize(b2lo_map* map, b2lo_shard_comm* c, const float* local_xyz, size_t m, size_t stride_floats, const float T_init[16],
                                       const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats, float* collective_ms) {
  int rc = shard_check(map, cfg);
  if (rc) return rc;
  if (!c || c->ctx != map->ctx || !T_init || !T_out || stride_floats < 3) return B2LO_E_ARG;
  if ((rc = b2lo_icp_shard_begin(map, local_xyz, m, stride_floats, T_init, cfg))) return rc;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  NcclApi& N = nccl_api();
  if (c->world > 1 && !c->p2p && !c->comm) { set_error("shard communicator has neither NCCL nor opened peers"); return B2LO_E_ARG; }
  IcpParams prm; int qpt;
  const size_t mm = m ? m : 1;
  shard_params(cfg, mm, prm, qpt);
  const int ctiles = (int)((mm + prm.ctile - 1) / prm.ctile);
  const CorrLaunch cl = corr_launch(ctx, ctiles);
  cudaStream_t s = ctx->stream;
  double* gmm = ctx->i_partial + (size_t)ctx->i_max_blocks * 28;
  double* js = gmm + 120;
  unsigned int* tk = reinterpret_cast<unsigned int*>(js + 132);
  const int ntiles = (int)((mm + TILE - 1) / TILE);
  const int grid = ntiles < 1 ? 1 : (ntiles > ctx->i_max_blocks ? ctx->i_max_blocks : ntiles);
  const bool peer = c->world > 1 && c->p2p;
  // start line: the ranks finish uploading their shards at different times; one tiny exchange lines them up so that the timed loop
  // (and the first real exchange) does not absorb the upload skew
  if (peer) k_p2p_exchange<<<1, P2P_MAXN, 0, s>>>(c->peer_args(++c->epoch), c->d_plan + 4, c->d_plan + 5, 1, 3, 1);
  else if (c->world > 1) B2_NCCL(N.AllReduce(c->d_plan + 4, c->d_plan + 5, 1, ncclDouble, ncclSum, c->comm, s));
  B2_CUDA(cudaEventRecord(ctx->ev0, s));
  for (int it = 0; it < cfg->max_iterations; ++it) {
    const bool last = it + 1 == cfg->max_iterations;
    corr_start(ctx, cl, s, map->d, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_tilesum,
                                        nullptr, nullptr, nullptr, nullptr, nullptr);
    if (peer) {   // 5 launches per iteration; the exchanges are stores into the peers' mailboxes inside the single-CTA kernels
      if (last) cudaEventRecord(c->ev[0], s);
      k_shard_pre_peer<<<1, PKO_THREADS, 0, s>>>(c->peer_args(c->epoch + 1), ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_cidx, ctx->i_blkcnt, ctx->i_blkoff,
                                                 ctx->i_tilesum, ctx->d_pko, ctx->d_pko_hits, c->d_stats, c->d_gathered, c->d_plan, c->d_sample, gmm);
      c->epoch += 2;
      if (last) cudaEventRecord(c->ev[1], s);
      if (cfg->use_adaptive_m_estimator) { launch<k_icp_pko2, 128, 1>(ctx, dim3((unsigned)(cfg->num_alpha_segments)), dim3((unsigned)(128)), 0, s, ctx->d_icp, prm, ctx->d_pko, gmm, js, tk); ctx->launches++; }
      launch<k_icp_gn<true>, TILE, 2>(ctx, dim3((unsigned)(grid)), dim3((unsigned)(TILE)), 0, s, map->d, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, nullptr, ctx->i_partial, c->d_acc);
      if (last) cudaEventRecord(c->ev[2], s);
      k_shard_post_peer<<<1, P2P_MAXN, 0, s>>>(c->peer_args(++c->epoch), ctx->d_icp, prm, c->d_acc);
      if (last) cudaEventRecord(c->ev[3], s);
      ctx->launches += 4;
      continue;
    }
    k_shard_stats<<<1, 256, 0, s>>>(ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_blkcnt, ctx->i_blkoff, c->d_stats, ctx->i_tilesum);
    if (last) cudaEventRecord(c->ev[0], s);
    if (c->world > 1) B2_NCCL(N.AllGather(c->d_stats, c->d_gathered, 3, ncclDouble, c->comm, s));
    else B2_CUDA(cudaMemcpyAsync(c->d_gathered, c->d_stats, 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_shard_plan<<<1, 32, 0, s>>>(ctx->d_icp, prm, c->d_gathered, c->world, c->rank, c->d_plan);
    k_shard_sample<<<1, MAXS, 0, s>>>(ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_cidx, ctx->i_blkoff, ctx->d_pko, ctx->d_pko_hits, 0, 0, 0.0, c->d_sample,
                                      c->d_plan);
    if (c->world > 1) B2_NCCL(N.AllReduce(c->d_sample, c->d_sample, 128, ncclDouble, ncclSum, c->comm, s));
    if (last) cudaEventRecord(c->ev[1], s);
    launch<k_icp_pko1, PKO_THREADS, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(PKO_THREADS)), 0, s, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_blkoff, ctx->d_pko,
                                         ctx->d_pko_hits, gmm, c->d_sample, 0, 0.0, ctx->i_tilesum, c->d_plan);
    if (cfg->use_adaptive_m_estimator) { launch<k_icp_pko2, 128, 1>(ctx, dim3((unsigned)(cfg->num_alpha_segments)), dim3((unsigned)(128)), 0, s, ctx->d_icp, prm, ctx->d_pko, gmm, js, tk); ctx->launches++; }
    launch<k_icp_gn<true>, TILE, 2>(ctx, dim3((unsigned)(grid)), dim3((unsigned)(TILE)), 0, s, map->d, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, nullptr, ctx->i_partial, c->d_acc);
    if (last) cudaEventRecord(c->ev[2], s);
    if (c->world > 1) B2_NCCL(N.AllReduce(c->d_acc, c->d_acc, 28, ncclDouble, ncclSum, c->comm, s));
    if (last) cudaEventRecord(c->ev[3], s);
    k_shard_finish<<<1, 32, 0, s>>>(ctx->d_icp, prm, c->d_acc);
    ctx->launches += 7;
  }
  launch<k_icp_end, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, s, ctx->d_icp);
  ctx->launches++;
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaEventRecord(ctx->ev1, s));
  B2_CUDA(cudaMemcpyAsync(ctx->h_icp, ctx->d_icp, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
  if (c->p2p) B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 50, c->d_err, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));   // [0] timeout flag, [2..8) ns of the last exchange of each kind
  B2_CUDA(cudaStreamSynchronize(s));
  ctx->d2h_bytes += sizeof(IcpState);
  if (c->p2p && ctx->h_counts[50]) { set_error("peer-memory exchange timed out (a rank did not reach the exchange)"); cudaMemset(c->d_err, 0, sizeof(int)); return B2LO_E_CUDA; }
  const IcpState* h = ctx->h_icp;
  Pose p;
  for (int i = 0; i < 9; ++i) p.R.m[i] = h->R[i];
  for (int i = 0; i < 3; ++i) p.t[i] = h->t[i];
  pose_to_T16(p, T_out);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  if (stats) {
    stats->status = h->status; stats->num_iterations = h->num_iterations; stats->num_correspondences = h->n_corr; stats->converged = h->converged;
    stats->initial_cost = h->initial_cost; stats->final_cost = h->final_cost; stats->device_ms = ms;
    std::memcpy(stats->it, h->trace, sizeof(stats->it));
  }
  if (collective_ms) {
    float a = 0.0f, b = 0.0f;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&b, c->ev[2], c->ev[3]);
    *collective_ms = a + b;
    if (peer) {   // peer mode: the exchanges sit inside fused kernels; the kernels report the ns of their last exchange themselves
      const unsigned long long* ns = reinterpret_cast<const unsigned long long*>(ctx->h_counts + 52);
      *collective_ms = (float)((double)(ns[0] + ns[1] + ns[2]) * 1e-6);
    }
  }
  if (h->status == B2LO_E_CAPACITY) { set_error("more than 2^22 correspondences over all ranks: outside the PKO sample tables"); return B2LO_E_CAPACITY; }
  return h->status == B2LO_S_INSUFFICIENT ? B2LO_S_INSUFFICIENT : B2LO_OK;
}

// ---- loop-closure ICP: optimize_loop / find_correspondences_loop, ICP.cpp:40-251, 465-585 -------------------------
// The matched keyframe's cloud (a few thousand voxel-filtered points, moved to world coordinates once) is the kNN target; it is an
// arbitrary cloud, not the map's one-centroid-per-cell grid, so the exact 5-NN is the warp-per-query scan of the dense target stream
// (knn_brute_warp: ~5 k x 5 k distance evaluations per iteration, microseconds on this GPU; the reference walks a nanoflann tree).
// Everything behind the correspondences - residual scale, PKO fit + arg-min, Gauss-Newton, pose update - is the scan-to-map machinery
// unchanged: the gate kernel below writes the same per-query records (plane normal + anchor point, f64 residual, tile compaction).
namespace b2 {
struct LoopPose { float Rm[9], tm[3], Ri[9], ti[3]; };   // matched keyframe pose and its rigid inverse (both f32)

__global__ void k_loop_iota(const int* __restrict__ d_npts, IcpState* st, int* unres, int* n_unres) { TL_START();
  if (st->done) return;
  const int npts = *d_npts;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) unres[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_unres = npts;
}
// plane through the 5 neighbours, NO distance gate (:571), anchor = the nearest neighbour taken through T_lw_last (f32 matrix, f64
// product, :524) and back through the matched pose in f32 (:139-142)
__global__ void __launch_bounds__(TILE) k_loop_gate(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_npts, IcpState* st, IcpParams prm,
                                                    LoopPose lp, const int* __restrict__ knn_idx, const int* __restrict__ knn_n, double* res,
                                                    int* slot_out, int* cidx, int* tilecnt, float4* plane, double* tilesum) { TL_START();
  if (st->done) return;
  __shared__ int sm[40];
  __shared__ double smd2[16];
  __shared__ float sR[9], sT[3];
  if (threadIdx.x < 9) sR[threadIdx.x] = st->R[threadIdx.x];
  if (threadIdx.x < 3) sT[threadIdx.x] = st->t[threadIdx.x];
  __syncthreads();
  const int npts = *d_npts;
  const int ntiles = (npts + TILE - 1) / TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int i = tile * TILE + threadIdx.x;
    int ok = 0;
    double a1 = 0.0, a2 = 0.0;
    if (i < npts) {
      float4 p = pts[i];
      float w[3], n[3] = {0, 0, 0}, c[3] = {0, 0, 0};
      transform_point(sR, sT, p.x, p.y, p.z, w);
      Top5 top;
      top.init();
      top.n = knn_n[i] < 0 ? 0 : knn_n[i];
      for (int k = 0; k < top.n; ++k) top.id[k] = knn_idx[i * KNN_K + k];
      double r = 0.0;
      const int state = knn_fit(M, top, w, 1.7976931348623157e308, n, c, &r);   // c = plane centroid, replaced by the anchor below
      ok = (state == 2);
      if (ok) {
        const float4 nn0 = M.l0_cent[top.id[0]];
        const double s0[3] = {(double)nn0.x, (double)nn0.y, (double)nn0.z};
        double pl[3];
        for (int a = 0; a < 3; ++a)
          pl[a] = add3((double)lp.Ri[a * 3] * s0[0], (double)lp.Ri[a * 3 + 1] * s0[1], (double)lp.Ri[a * 3 + 2] * s0[2]) + (double)lp.ti[a];
        const float pm[3] = {(float)pl[0], (float)pl[1], (float)pl[2]};
        float Rq[3];
        mat3_vec(lp.Rm, pm, Rq);
        c[0] = Rq[0] + lp.tm[0]; c[1] = Rq[1] + lp.tm[1]; c[2] = Rq[2] + lp.tm[2];
        a1 = r; a2 = r * r;
      }
      slot_out[i] = ok ? 0 : -1;
      res[i] = r;
      plane[2 * i] = make_float4(n[0], n[1], n[2], c[0]);
      plane[2 * i + 1] = make_float4(c[1], c[2], 0.0f, 0.0f);
    }
    int total;
    int off = block_excl_scan(ok, &total, sm);
    if (ok) cidx[tile * TILE + off] = i;
    block_sum2_t0(a1, a2, smd2);
    if (threadIdx.x == 0) { tilecnt[tile] = total; tilesum[2 * tile] = a1; tilesum[2 * tile + 1] = a2; }
  }
}
// validation (:214-247): share of the current keyframe's points whose nearest target point is closer than 1 m at the final pose
__global__ void k_loop_inliers(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_npts, const IcpState* st,
                               const int* __restrict__ knn_idx, const int* __restrict__ knn_n, int* count) { TL_START();
  const int npts = *d_npts;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) {
    if (knn_n[i] < 1) continue;
    float4 p = pts[i];
    float w[3];
    transform_point(st->R, st->t, p.x, p.y, p.z, w);
    const float4 c = M.l0_cent[knn_idx[i * KNN_K]];
    if (sqrtf(knn_dist2(w, c.x, c.y, c.z)) < 1.0f) atomicAdd(count, 1);
  }
}
// forces the brute-force search of every query regardless of the done flag (used for the validation pass)
__global__ void k_loop_iota_all(const int* __restrict__ d_npts, int* unres, int* n_unres) { TL_START();
  const int npts = *d_npts;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) unres[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_unres = npts;
}
__global__ void k_loop_set_done(IcpState* st, int v) { TL_START(); if (threadIdx.x == 0 && blockIdx.x == 0) st->done = v; }

}  // namespace b2

extern "C" int b2lo_icp_optimize_loop(b2lo_ctx* ctx, const float* curr_xyz, size_t m_curr, size_t curr_stride_floats, const float T_curr[16],
                                      const float* matched_xyz, size_t m_matched, size_t matched_stride_floats, const float T_matched[16],
                                      const b2lo_icp_cfg* cfg, float T_rel[16], float* inlier_ratio, b2lo_icp_stats* stats) {
  if (!ctx || !T_curr || !T_matched || !cfg || !T_rel) return B2LO_E_ARG;
  if (curr_stride_floats < 3 || matched_stride_floats < 3) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  if (inlier_ratio) *inlier_ratio = 0.0f;
  if (stats) std::memset(stats, 0, sizeof *stats);
  { Pose id; id.R = mat3_identity(); id.t[0] = id.t[1] = id.t[2] = 0.0f; pose_to_T16(id, T_rel); }
  if (!curr_xyz || !matched_xyz || m_curr == 0 || m_matched == 0) {   // empty clouds: no correspondences (:474-479) -> break -> false
    if (stats) stats->status = B2LO_S_INSUFFICIENT;
    return B2LO_S_INSUFFICIENT;
  }
  int rc = ctx_reserve_points(ctx, m_curr > m_matched ? m_curr : m_matched);
  if (rc) return rc;
  if ((rc = knn_reserve(ctx))) return rc;
  if ((rc = icp_build_pko(ctx, cfg))) return rc;
  if (!ctx->l_cnt) { B2_CUDA(cudaMalloc((void**)&ctx->l_cnt, 4 * sizeof(int))); }
  cudaStream_t s = ctx->stream;
  B2_CUDA(cudaEventRecord(ctx->ev0, s));
  // target: the matched keyframe's cloud in world coordinates (transform_point_cloud, :56-58), kept in d_world
  if ((rc = ctx_stage_h2d(ctx, matched_xyz, m_matched, matched_stride_floats, 1, ctx->d_world, ctx->l_cnt))) return rc;
  if ((rc = ctx_transform(ctx, ctx->d_world, ctx->l_cnt, m_matched, T_matched, ctx->d_world))) return rc;
  if ((rc = ctx_stage_h2d(ctx, curr_xyz, m_curr, curr_stride_floats, 1, ctx->d_query, ctx->d_nquery))) return rc;
  B2_CUDA(cudaMemsetAsync(ctx->l_cnt + 1, 0, sizeof(int), s));
  MapDev Ml;
  std::memset(&Ml, 0, sizeof Ml);
  Ml.l0_cent = ctx->d_world;
  Ml.ctr = ctx->l_cnt;
  LoopPose lp;
  {
    Pose pm = pose_from_T16(T_matched);
    for (int i = 0; i < 9; ++i) lp.Rm[i] = pm.R.m[i];
    for (int i = 0; i < 3; ++i) lp.tm[i] = pm.t[i];
    Mat3 Rt = mat3_t(pm.R);
    float rt[3];
    mat3_vec(Rt.m, pm.t, rt);
    for (int i = 0; i < 9; ++i) lp.Ri[i] = Rt.m[i];
    for (int i = 0; i < 3; ++i) lp.ti[i] = -rt[i];
  }
  IcpParams prm;
  prm.max_iterations = 100;   // :79
  prm.min_corr = cfg->min_correspondence_points; prm.use_robust = cfg->use_robust_loss;
  prm.loss_type = cfg->loss_type; prm.use_pko = cfg->use_adaptive_m_estimator; prm.use_surfel = 0;
  prm.tol_t = cfg->translation_tolerance; prm.tol_r = cfg->rotation_tolerance; prm.max_dist = 1.7976931348623157e308;
  prm.robust_delta = cfg->robust_loss_delta;
  prm.ctile = TILE;
  if ((rc = sp_begin_write(ctx))) return rc;
  for (int i = 0; i < 16; ++i) ctx->h_sp->T_init[i] = T_curr[i];
  ctx->h_sp->force_scale = 0.0;
  if ((rc = sp_upload(ctx, offsetof(ScanParams, T_init), SP_POSE_BYTES))) return rc;
  launch<k_icp_begin, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, s, ctx->d_icp, ctx->d_sp);
  ctx->launches++;
  const int ntiles = (int)((m_curr + TILE - 1) / TILE);
  const int grid = ntiles > ctx->i_max_blocks ? ctx->i_max_blocks : ntiles;
  const int gb = ctx->sm_count * 2;
  double* gmm = ctx->i_partial + (size_t)ctx->i_max_blocks * 28;
  double* js = gmm + 120;
  unsigned int* tk = reinterpret_cast<unsigned int*>(js + 132);
  const int CHUNK = 8;   // iterations enqueued between two looks at the done flag (converged iterations turn into no-ops)
  for (int it0 = 0; it0 < 100; it0 += CHUNK) {
    for (int it = it0; it < it0 + CHUNK && it < 100; ++it) {
      k_loop_iota<<<grid, TILE, 0, s>>>(ctx->d_nquery, ctx->d_icp, ctx->k_unres, ctx->k_nunres);
      k_knn_brute<<<gb, 256, 0, s>>>(Ml, ctx->d_query, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres);
      k_loop_gate<<<grid, TILE, 0, s>>>(Ml, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, lp, ctx->k_idx, ctx->k_n, ctx->i_res, ctx->i_slot, ctx->i_cidx,
                                        ctx->i_blkcnt, ctx->k_plane, ctx->i_tilesum);
      launch<k_icp_pko1, PKO_THREADS, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(PKO_THREADS)), 0, s, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->i_cidx, ctx->i_blkcnt, ctx->i_blkoff, ctx->d_pko,
                                           ctx->d_pko_hits, gmm, nullptr, 0, 0.0, ctx->i_tilesum, nullptr);
      if (cfg->use_adaptive_m_estimator) launch<k_icp_pko2, 128, 1>(ctx, dim3((unsigned)(cfg->num_alpha_segments)), dim3((unsigned)(128)), 0, s, ctx->d_icp, prm, ctx->d_pko, gmm, js, tk);
      launch<k_icp_gn<false>, TILE, 1>(ctx, dim3((unsigned)(grid)), dim3((unsigned)(TILE)), 0, s, Ml, ctx->d_query, ctx->d_nquery, ctx->d_icp, prm, ctx->i_res, ctx->i_slot, ctx->k_plane, ctx->i_partial, nullptr);
      ctx->launches += cfg->use_adaptive_m_estimator ? 6 : 5;
    }
    B2_CUDA(cudaMemcpyAsync(ctx->h_icp, ctx->d_icp, offsetof(IcpState, trace), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    ctx->d2h_bytes += offsetof(IcpState, trace);
    if (ctx->h_icp->done) break;
  }
  // success = converged within 100 iterations (:196-201); insufficient correspondences just break out of the loop (:86-89)
  const bool converged = ctx->h_icp->done == 1 && ctx->h_icp->converged;
  int inl = 0;
  if (converged) {
    k_loop_iota_all<<<grid, TILE, 0, s>>>(ctx->d_nquery, ctx->k_unres, ctx->k_nunres);
    k_loop_set_done<<<1, 32, 0, s>>>(ctx->d_icp, 0);   // the search kernel is gated by the flag
    k_knn_brute<<<gb, 256, 0, s>>>(Ml, ctx->d_query, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres);
    k_loop_set_done<<<1, 32, 0, s>>>(ctx->d_icp, 1);
    k_loop_inliers<<<grid, TILE, 0, s>>>(Ml, ctx->d_query, ctx->d_nquery, ctx->d_icp, ctx->k_idx, ctx->k_n, ctx->l_cnt + 1);
    ctx->launches += 5;
    B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 40, ctx->l_cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  B2_CUDA(cudaMemsetAsync(ctx->k_nunres, 0, sizeof(int), s));   // the scan-to-map KDTree path expects an empty queue
  B2_CUDA(cudaEventRecord(ctx->ev1, s));
  B2_CUDA(cudaMemcpyAsync(ctx->h_icp, ctx->d_icp, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  B2_CUDA(cudaGetLastError());
  ctx->d2h_bytes += sizeof(IcpState) + sizeof(int);
  if (converged) inl = ctx->h_counts[40];
  const IcpState* h = ctx->h_icp;
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  const float ratio = converged ? (float)inl / (float)m_curr : 0.0f;
  const bool success = converged && !(ratio < 0.5f);   // :249-251
  if (stats) {
    stats->status = success ? B2LO_OK : B2LO_S_INSUFFICIENT; stats->num_iterations = h->num_iterations; stats->num_correspondences = h->n_corr;
    stats->converged = h->converged; stats->initial_cost = h->initial_cost; stats->final_cost = h->final_cost; stats->device_ms = ms;
    std::memcpy(stats->it, h->trace, sizeof(stats->it));
  }
  if (inlier_ratio) *inlier_ratio = ratio;
  if (converged) {   // optimized_relative_transform = curr_pose^-1 * optimized_curr_pose (:198), set as soon as the loop converges
    Pose opt;
    for (int i = 0; i < 9; ++i) opt.R.m[i] = h->R[i];
    for (int i = 0; i < 3; ++i) opt.t[i] = h->t[i];
    Pose cur = pose_from_T16(T_curr);
    Pose inv = pose_inv(cur);
    pose_to_T16(pose_mul(inv, opt), T_rel);
  }
  return success ? B2LO_OK : B2LO_S_INSUFFICIENT;
}

#ifdef B2LO_TIMELINE
namespace b2 { int tl_fetch_icp(unsigned long long* out, int cap) { return tl_fetch(out, cap); } }
#endif
