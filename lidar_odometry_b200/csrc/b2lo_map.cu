// b2lo_map.cu — K6/K7: the GPU-resident 2-level Z-order voxel hash and its incremental update.
//
// Replaces map::VoxelMap (/root/reference/src/database/VoxelMap.h:188-332, VoxelMap.cpp):
//   UpdateVoxelMap :128-262 (radius cull :146-158, AddPoint :99-120, RegisterToParent :77-80,
//   UnregisterFromParent :82-97, surfel refit / non-planar purge :187-261), GetSurfelAtPoint :368-386,
//   GetPointCloud :388-403, GetL1Surfels :405-418, Clear :122-126.
//
// The reference is sequential and its containers leak their ORDER into the results (SURVEY.md hard
// part 4): L0 iteration order (= GetPointCloud order, = cull order) and each L1's child-set order
// (= summation order of the surfel fit).  The kernels below reproduce those orders exactly, in parallel:
//   cull    : mark -> scan -> replay of "erase in ascending position order, swap-with-last" on indices only
//             (closed form hole_t <- element n-t when no tail voxel is culled, the usual case; otherwise one
//             thread replays the k erases in shared memory), then the moves are applied in parallel;
//             per affected parent one thread replays its child-set swap-erases in removal order.
//   insert  : per-voxel lists of the new points (atomicExch) sorted by point index -> running mean
//             c = (c*n + p)/(n+1) replayed in input order; new voxels ranked by first-seen index with a
//             prefix scan and appended; children appended to their parent in creation order.
//   surfels : one thread per affected L1: gather children in child-set order, f32 covariance, Jacobi SVD,
//             planarity gate; non-planar parents are purged with all their L0 children.  The purge is an
//             arbitrary-order swap-erase on the dense L0 vector; one thread replays it on indices only
//             (shared memory), then the moves are applied in parallel.
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#define B2LO_TL_FILE 4
#include "b2lo_internal.h"
#include "b2lo_launch.cuh"

namespace b2 {

enum { US_K = 0, US_S = 1, US_NWORK = 2, US_NNEW = 3, US_NAFF = 4, US_NPURGE = 5, US_KPURGE = 6, US_N0 = 7, US_ERR = 8, US_TICKET = 9, US_COUNT = 16 };
enum { CT_N0 = 0, CT_N1 = 1, CT_TOMB0 = 2, CT_TOMB1 = 3, CT_ERR = 4, CT_SURF = 5 };
constexpr int ERR_RANGE = 1, ERR_CAP = 2, ERR_INTERNAL = 4;
constexpr int SIM_SMEM_K = 12000;  // (4k+2) ints <= 192 KB of dynamic shared memory
constexpr int SIM_SMEM_BYTES = (4 * SIM_SMEM_K + 2) * (int)sizeof(int);

__device__ __forceinline__ unsigned long long child_key(int px, int py, int pz, int f, int code) {
  return key_pack(px * f + code % 3, py * f + (code / 3) % 3, pz * f + code / 9);
}

// Replay of k swap-with-last erases (in the given order) on a dense vector of size n, on indices only; whole CTA.
// seq_pos[t-1]: ORIGINAL position of the t-th erased element.  Output fill[t] (at aux[3k+1+t]): original position of
// the survivor that finally sits in the hole the t-th erase leaves below the new size s = n - k.
//   loc[t]   : where the t-th erased element currently sits (>= s: tail position; -h: in the hole of erase h)
//   occ*[q-s]: current occupant of tail position q (original position id, its erase time or 0)
// Fast path (no erased element lives in the tail [s, n)): erase t moves original element n-t into hole t.
__device__ void swap_erase_sim(int k, int n, const int* seq_pos, int* aux, int* smem, int smem_cap_k) {
  const int s = n - k;
  int tail = 0;
  for (int t = threadIdx.x; t < k; t += blockDim.x) tail |= (seq_pos[t] >= s);
  if (!__syncthreads_or(tail)) {
    for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) aux[3 * k + 1 + t] = n - t;
    __syncthreads();
    return;
  }
  const bool in_smem = (k <= smem_cap_k);
  int* loc = in_smem ? smem : aux;               // k+1
  int* occ_id = loc + (k + 1);                   // k
  int* occ_t = occ_id + k;                       // k
  int* fill = occ_t + k;                         // k+1
  for (int q = threadIdx.x; q < k; q += blockDim.x) { occ_id[q] = s + q; occ_t[q] = 0; }
  __syncthreads();
  for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) {
    int P = seq_pos[t - 1];
    if (P >= s) { loc[t] = P; occ_t[P - s] = t; } else loc[t] = -t;
    fill[t] = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t <= k; ++t) {
      int b = n - t;
      int l = loc[t];
      int yid = occ_id[b - s], yt = occ_t[b - s];
      if (l >= 0) { if (l != b) { occ_id[l - s] = yid; occ_t[l - s] = yt; if (yt) loc[yt] = l; } }
      else { fill[-l] = yid; if (yt) loc[yt] = l; }
    }
  }
  __syncthreads();
  if (in_smem) for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) aux[3 * k + 1 + t] = fill[t];
  __syncthreads();
}
// apply the moves of a replay: hole seq_pos[t-1] (< s) <- element fill[t]; whole CTA
__device__ void swap_erase_apply(const MapDev& M, int k, int s, const int* seq_pos, const int* aux) {
  const int* fill = aux + 3 * k + 1;
  for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) {
    int dst = seq_pos[t - 1];
    if (dst < 0 || dst >= s) continue;
    int src = fill[t];
    uint32_t sl = M.l0_slot[src];
    M.l0_cent[dst] = M.l0_cent[src];
    M.l0_key[dst] = M.l0_key[src];
    M.l0_slot[dst] = sl;
    M.l0_tab[sl].pos = (uint32_t)dst;
  }
}

// ---- cull ---------------------------------------------------------------------------------------------
// mark: ||c - sensor||^2 > r^2 (VoxelMap.cpp:146-158), per-tile counts; the last CTA scans the tile counts
struct k_cull_mark { static __device__ __forceinline__ void run(MapDev M, int n0, float sx, float sy, float sz, float r2, uint8_t* flag, int* blkcnt, int* blkoff,
                                                    int* us) { TL_START();
  // the gate and the first device-side scalars are loaded together: one memory round trip instead of two ahead of the work
  const int gate_v = M.gate ? *M.gate : 1;
  if (M.sensor_dev) { sx = M.sensor_dev[3]; sy = M.sensor_dev[7]; sz = M.sensor_dev[11]; }   // translation of a row-major 4x4 pose
  if (n0 < 0) n0 = M.ctr[CT_N0];   // replayed launch sequence: the live voxel count is on the device
  if (!gate_v) return;
  __shared__ int sm[40];
  __shared__ int s_last;
  int ntiles = (n0 + 1023) / 1024;
  // 4 tiles per trip: four independent 16 B loads in flight per thread (the scan streams 16 B / voxel from HBM)
  for (int tile0 = blockIdx.x * 4; tile0 < ntiles; tile0 += gridDim.x * 4) {
    float4 c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int pos = (tile0 + u) * 1024 + threadIdx.x;
      c[u] = (pos < n0) ? M.l0_cent[pos] : make_float4(sx, sy, sz, 0.0f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int pos = (tile0 + u) * 1024 + threadIdx.x;
      float d[3] = {c[u].x - sx, c[u].y - sy, c[u].z - sz};
      int mk = (pos < n0) && (sqn3(d) > r2);  // (centroid - sensor).squaredNorm() > radius_sq  (VoxelMap.cpp:149-150)
      if (pos < n0) flag[pos] = (uint8_t)mk;
      int tot = __syncthreads_count(mk);
      if (threadIdx.x == 0 && tile0 + u < ntiles) blkcnt[tile0 + u] = tot;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&us[US_TICKET], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int base = 0;
  for (int t0 = 0; t0 < ntiles; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    int c = t < ntiles ? ((volatile int*)blkcnt)[t] : 0, tot;
    int e = block_excl_scan(c, &tot, sm);
    if (t < ntiles) blkoff[t] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) { us[US_K] = base; us[US_S] = n0 - base; us[US_N0] = n0 - base; us[US_NWORK] = 0; us[US_TICKET] = 0; }
} };
// removed[] (ascending dense position) and the list of parents that lose children
struct k_cull_lists { static __device__ __forceinline__ void run(MapDev M, int n0, const uint8_t* flag, const int* blkoff, int* us, int* removed, int* l1work) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  __shared__ int sm[40];
  const int k = us[US_K], s_keep = us[US_S];
  if (!gate_v || k == 0) return;
  if (n0 < 0) n0 = k + s_keep;   // voxel count before the cull (k_cull_mark)
  int ntiles = (n0 + 1023) / 1024;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int pos = tile * 1024 + threadIdx.x;
    int mk = (pos < n0) ? flag[pos] : 0, tot;
    int pre = blkoff[tile] + block_excl_scan(mk, &tot, sm);
    if (pos >= n0 || !mk) continue;
    removed[pre] = pos;
    int x, y, z;
    key_unpack(M.l0_key[pos], x, y, z);
    unsigned long long pk = key_pack(parent_coord(x, M.factor), parent_coord(y, M.factor), parent_coord(z, M.factor));
    int s1 = l1_find(M, pk);
    if (s1 >= 0 && atomicCAS(&M.l1_meta[s1].mark, 0, 1) == 0) l1work[atomicAdd(&us[US_NWORK], 1)] = s1;
  }
} };
// one CTA: (1) per affected parent replay occupied_children.erase() in removal (= L0 dense) order, one warp per parent,
// one lane per child; (2) replay the k dense-vector erases on indices; (3) apply the moves, drop the hash entries
struct k_cull_fix { static __device__ __forceinline__ void run(MapDev M, const uint8_t* flag, int* us, const int* l1work, const int* removed, int* aux, int n0) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  extern __shared__ int smem[];
  const int k = us[US_K];
  const int s = us[US_S];
  const int nwork = us[US_NWORK];
  if (!gate_v || k == 0) return;
  if (n0 < 0) n0 = k + s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int wi = warp; wi < nwork; wi += nwarps) {
    int s1 = l1work[wi];
    L1Meta* mt = &M.l1_meta[s1];
    unsigned long long k1 = M.l1_tab[s1].key;
    int px, py, pz;
    key_unpack(k1 & KEY_MASK, px, py, pz);
    int n = mt->nchild;
    int code = lane < n ? mt->child[lane] : -1;
    int rp = INT_MAX;  // dense position of this lane's child if it is being removed
    if (code >= 0) {
      int s0 = l0_find(M, child_key(px, py, pz, M.factor, code));
      if (s0 >= 0) { int pos = (int)M.l0_tab[s0].pos; if (flag[pos]) rp = pos; }
    }
    unsigned rem = __ballot_sync(0xffffffffu, rp != INT_MAX);
    int nr = __popc(rem);
    // removal order = ascending dense position: rank of this lane's child among the removed ones
    int myrank = 0;
    for (int l = 0; l < 32; ++l) { int op = __shfl_sync(0xffffffffu, rp, l); if (op < rp) ++myrank; }
    // gather the removed codes in removal order into lane 0 by rank
    int ord_code[27];
    for (int a = 0; a < nr; ++a) {
      unsigned m = __ballot_sync(0xffffffffu, rp != INT_MAX && myrank == a);
      int src = __ffs(m) - 1;
      ord_code[a] = __shfl_sync(0xffffffffu, code, src);
    }
    if (lane == 0) {
      for (int a = 0; a < nr; ++a) {
        int idx = 0;
        while (idx < n && mt->child[idx] != (uint8_t)ord_code[a]) ++idx;
        if (idx < n) { mt->child[idx] = mt->child[n - 1]; --n; }
      }
      mt->nchild = (uint8_t)n;
      mt->mark = 0;
      if (n < 5) k1 &= ~SURFEL_BIT;  // has_surfel = false, last_child_count kept (VoxelMap.cpp:90-92)
      if (n == 0) { k1 = KEY_TOMB; mt->last_child_count = 0; atomicSub(&M.ctr[CT_N1], 1); atomicAdd(&M.ctr[CT_TOMB1], 1); }
      M.l1_tab[s1].key = k1;
    }
  }
  __syncthreads();
  swap_erase_sim(k, n0, removed, aux, smem, SIM_SMEM_K);
  for (int r = threadIdx.x; r < k; r += blockDim.x) {  // drop the hash entries of the removed voxels (before slots move)
    L0Entry* e = &M.l0_tab[M.l0_slot[removed[r]]];
    e->key = KEY_TOMB;
  }
  __syncthreads();
  swap_erase_apply(M, k, s, removed, aux);
  if (threadIdx.x == 0) atomicAdd(&M.ctr[CT_TOMB0], k);
} };

// ---- insert ---------------------------------------------------------------------------------------------
// local != nullptr: the update's points are the feature cloud `local` moved by the row-major pose T16 (transform_point_cloud,
// PointCloudUtils.cpp:120-121, the arithmetic of k_transform_dev); they are computed here and left in `pts` for the kernels behind
struct k_ins_probe { static __device__ __forceinline__ void run(MapDev M, float4* pts, const int* __restrict__ d_m, int* us, int* pslot, int* nxt, FEntry* atab, int alog2,
                            int* alist, const float4* __restrict__ local, const float* __restrict__ T16) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  const int m = *d_m;
  float tm[12];
  if (local) for (int i = 0; i < 12; ++i) tm[i] = T16[i];
  if (!gate_v) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    float4 p;
    if (local) {
      const float4 q = local[i];
      p = make_float4(((tm[0] * q.x + tm[1] * q.y) + tm[2] * q.z) + tm[3] * 1.0f, ((tm[4] * q.x + tm[5] * q.y) + tm[6] * q.z) + tm[7] * 1.0f,
                      ((tm[8] * q.x + tm[9] * q.y) + tm[10] * q.z) + tm[11] * 1.0f, 0.0f);
      pts[i] = p;
    } else p = pts[i];
    int x = voxel_coord(p.x, M.voxel), y = voxel_coord(p.y, M.voxel), z = voxel_coord(p.z, M.voxel);
    int ax = voxel_coord(p.x, M.scale1), ay = voxel_coord(p.y, M.scale1), az = voxel_coord(p.z, M.scale1);
    if (!key_in_range(x, y, z) || !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z)) { atomicOr(&us[US_ERR], ERR_RANGE); pslot[i] = -1; continue; }
    // the affected-set probe below does not depend on the voxel probe: its first key load goes out ahead of it, so that the two
    // dependent chains of memory round trips overlap
    const unsigned long long ak = key_pack(ax, ay, az);
    const uint32_t mask = (1u << alog2) - 1u;
    uint32_t h = hash_slot(ak, alog2);
    unsigned long long kk = *((volatile unsigned long long*)&atab[h].key);
    bool ins;
    int s0 = l0_find_or_insert(M, key_pack(x, y, z), &ins);
    if (s0 < 0) { atomicOr(&us[US_ERR], ERR_CAP); pslot[i] = -1; continue; }
    L0Entry* e = &M.l0_tab[s0];
    atomicMin(&e->first, (unsigned)i);
    atomicAdd(&e->cnt, 1);                 // idle value -1: holds count - 1
    const int prev = atomicExch(&e->head, i);      // idle value -1; stored below, once the affected-set probe is on its way
    pslot[i] = s0;
    // affected_L1.insert(PointToVoxelKey(point, 1))  (VoxelMap.cpp:178-179) — float division by voxel*3
    for (;;) {
      if (kk == ak) break;
      if (kk == KEY_EMPTY) {
        unsigned long long old = atomicCAS(&atab[h].key, KEY_EMPTY, ak);
        if (old == KEY_EMPTY) { alist[atomicAdd(&us[US_NAFF], 1)] = (int)h; break; }
        if (old == ak) break;
      }
      h = (h + 1) & mask;
      kk = *((volatile unsigned long long*)&atab[h].key);
    }
    atomicMin(&atab[h].first, (unsigned)i);
    nxt[i] = prev;
  }
} };
// leader (first point of each touched voxel) replays AddPoint over the voxel's points in input order
// weighted = 1: the "points" are voxels of a re-hash (w = point_count), merged as in ApplyTransformAndRehash (VoxelMap.cpp:283-297)
struct k_ins_apply { static __device__ __forceinline__ void run(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_m, const int* pslot, const int* nxt, int* isnew,
                            float4* newc, int weighted) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  const int m = *d_m;
  if (!gate_v) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    int s0 = pslot[i];
    if (s0 < 0) { isnew[i] = 0; continue; }
    L0Entry* e = &M.l0_tab[s0];
    if (e->first != (unsigned)i) { isnew[i] = 0; continue; }
    int cnt = e->cnt + 1;
    uint32_t pos = e->pos;
    float cx, cy, cz; int n;
    if (pos == POS_PENDING) { n = 0; cx = cy = cz = 0.0f; }
    else { float4 c = M.l0_cent[pos]; cx = c.x; cy = c.y; cz = c.z; n = __float_as_int(c.w); }
    auto add = [&](int j) {
      float4 p = pts[j];
      if (weighted) {
        int c2 = __float_as_int(p.w);
        if (n == 0) { cx = p.x; cy = p.y; cz = p.z; n = c2; }
        else {
          float n1 = (float)n, n2 = (float)c2, ns = n1 + n2;
          cx = (cx * n1 + p.x * n2) / ns; cy = (cy * n1 + p.y * n2) / ns; cz = (cz * n1 + p.z * n2) / ns;
          n += c2;
        }
      }
      else if (n == 0) { cx = p.x; cy = p.y; cz = p.z; n = 1; }
      else {
        float fn = (float)n, fn1 = (float)(n + 1);
        cx = (cx * fn + p.x) / fn1; cy = (cy * fn + p.y) / fn1; cz = (cz * fn + p.z) / fn1;
        ++n;
      }
    };
    if (cnt == 1) add(i);
    else if (cnt <= 32) {
      int idx[32]; int q = 0;
      for (int j = e->head; j >= 0 && q < 32; j = nxt[j]) idx[q++] = j;
      for (int a = 1; a < q; ++a) { int v = idx[a], b = a - 1; while (b >= 0 && idx[b] > v) { idx[b + 1] = idx[b]; --b; } idx[b + 1] = v; }
      for (int a = 0; a < q; ++a) add(idx[a]);
    } else {
      int last = -1;
      for (int a = 0; a < cnt; ++a) {  // selection by repeated list walks (pathological multiplicities only)
        int best = INT_MAX;
        for (int j = e->head; j >= 0; j = nxt[j]) if (j > last && j < best) best = j;
        add(best); last = best;
      }
    }
    float4 out = make_float4(cx, cy, cz, __int_as_float(n));
    if (pos == POS_PENDING) { newc[i] = out; isnew[i] = 1; }
    else { M.l0_cent[pos] = out; isnew[i] = 0; }
    e->first = 0xFFFFFFFFu; e->cnt = -1; e->head = -1;  // scratch back to idle
  }
} };
// one CTA: rank of every new voxel in first-seen order (4 points per thread); the rank is also left in the voxel's
// hash entry so that siblings can order themselves (k_ins_place)
struct k_ins_scan { static __device__ __forceinline__ void run(MapDev M, const int* __restrict__ d_m, const int* isnew, const int* pslot, int* newrank, int* us, int* newlist) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  __shared__ int sm[40];
  const int m = *d_m;
  if (!gate_v) return;
  int base = 0;
  for (int t0 = 0; t0 < m; t0 += 4 * blockDim.x) {
    const int i0 = t0 + 4 * threadIdx.x;
    int f[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { f[k] = (i0 + k < m) ? isnew[i0 + k] : 0; sum += f[k]; }
    int tot;
    int e = base + block_excl_scan(sum, &tot, sm);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < m) {
        newrank[i0 + k] = e;
        if (f[k]) { M.l0_tab[pslot[i0 + k]].rank = e; newlist[e] = i0 + k; ++e; }   // newlist: the points that created a voxel, compact, in creation order
      }
    }
    base += tot;
  }
  if (threadIdx.x == 0) {
    us[US_NNEW] = base;
    if ((long long)us[US_N0] + base > (long long)M.l0_cap) atomicOr(&us[US_ERR], ERR_CAP);
  }
} };
// Bulk inserts (more than BULK_UPD points, e.g. ApplyTransformAndRehash of a 10^7-voxel map): the same ranks from a three-step scan
// over INS_CHUNK-point chunks instead of one CTA walking the whole update.
constexpr int INS_CHUNK = 4096;
constexpr size_t BULK_UPD = 1u << 16;
__global__ void __launch_bounds__(1024) k_ins_scan_part(const int* __restrict__ d_m, const int* __restrict__ isnew, int* part) { TL_START();
  __shared__ int sm[40];
  const int m = *d_m;
  const int nchunks = (m + INS_CHUNK - 1) / INS_CHUNK;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int i0 = c * INS_CHUNK + 4 * threadIdx.x;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) sum += (i0 + k < m) ? isnew[i0 + k] : 0;
    int tot;
    block_excl_scan(sum, &tot, sm);
    if (threadIdx.x == 0) part[c] = tot;
  }
}
__global__ void __launch_bounds__(1024) k_ins_scan_top(MapDev M, const int* __restrict__ d_m, int* part, int* us) { TL_START();
  __shared__ int sm[40];
  const int m = *d_m;
  const int nchunks = (m + INS_CHUNK - 1) / INS_CHUNK;
  int base = 0;
  for (int t0 = 0; t0 < nchunks; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const int v = t < nchunks ? part[t] : 0;
    int tot;
    const int e = block_excl_scan(v, &tot, sm);
    if (t < nchunks) part[t] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) {
    us[US_NNEW] = base;
    if ((long long)us[US_N0] + base > (long long)M.l0_cap) atomicOr(&us[US_ERR], ERR_CAP);
  }
}
__global__ void __launch_bounds__(1024) k_ins_scan_apply(MapDev M, const int* __restrict__ d_m, const int* __restrict__ isnew, const int* __restrict__ pslot,
                                                         int* newrank, const int* __restrict__ part, int* newlist) { TL_START();
  __shared__ int sm[40];
  const int m = *d_m;
  const int nchunks = (m + INS_CHUNK - 1) / INS_CHUNK;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int i0 = c * INS_CHUNK + 4 * threadIdx.x;
    int f[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { f[k] = (i0 + k < m) ? isnew[i0 + k] : 0; sum += f[k]; }
    int tot;
    int e = part[c] + block_excl_scan(sum, &tot, sm);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < m) {
        newrank[i0 + k] = e;
        if (f[k]) { M.l0_tab[pslot[i0 + k]].rank = e; newlist[e] = i0 + k; ++e; }   // newlist: the points that created a voxel, compact, in creation order
      }
    }
  }
}
// creation ranks back to idle after a bulk insert (k_upd_close does it itself for keyframe-sized updates)
__global__ void __launch_bounds__(256) k_rank_clear(MapDev M, const int* __restrict__ d_m, const int* __restrict__ pslot, const int* __restrict__ isnew) { TL_START();
  const int m = *d_m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) if (isnew[i]) M.l0_tab[pslot[i]].rank = -1;
}
// one warp per new voxel: append it to the dense vector, find-or-create its parent and place it in the parent's child
// list in creation order (RegisterToParent, VoxelMap.cpp:77-80).  Lane c probes sibling cell c of the parent: the
// position is (#siblings that already existed) + (#new siblings created earlier), so nobody has to read nchild while
// it is being updated; the first new sibling writes the new count.
struct k_ins_place { static __device__ __forceinline__ void run(MapDev M, const int* __restrict__ d_m, int* us, const int* pslot, const int* isnew,
                                                   const int* newrank, const float4* newc, const int* __restrict__ newlist) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  const int m = *d_m;
  const int err_v = us[US_ERR];
  const int base = us[US_N0];
  if (!gate_v || (err_v & ERR_CAP)) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  // the compact list of the points that created a voxel (k_ins_scan): a warp never looks at the ~85 % of the points that did not
  const int nnew = us[US_NNEW];
  (void)m; (void)isnew;
  for (int r = warp; r < nnew; r += nwarps) {
    const int i = newlist[r];
    const int s0 = pslot[i];
    const int myrank = newrank[i];
    const int pos = base + myrank;
    const unsigned long long key = M.l0_tab[s0].key;
    int x, y, z;
    key_unpack(key, x, y, z);
    const int px = parent_coord(x, M.factor), py = parent_coord(y, M.factor), pz = parent_coord(z, M.factor);
    const int mycode = (x - px * M.factor) + 3 * (y - py * M.factor) + 9 * (z - pz * M.factor);
    int s1 = -1;
    if (lane == 0) {
      M.l0_cent[pos] = newc[i];
      M.l0_key[pos] = key;
      M.l0_slot[pos] = (uint32_t)s0;
      M.l0_tab[s0].pos = (uint32_t)pos;
      bool ins;
      s1 = l1_find_or_insert(M, key_pack(px, py, pz), &ins);  // a fresh slot has zeroed meta: no children, no surfel
      if (s1 < 0) atomicOr(&us[US_ERR], ERR_CAP);
      else if (ins) atomicAdd(&M.ctr[CT_N1], 1);
    }
    s1 = __shfl_sync(0xffffffffu, s1, 0);
    if (s1 < 0) continue;
    int is_old = 0, is_new = 0, earlier = 0;
    if (lane < 27 && lane != mycode) {
      int sc = l0_find(M, child_key(px, py, pz, M.factor, lane));
      if (sc >= 0) {
        int rk = M.l0_tab[sc].rank;
        if (rk < 0) is_old = 1; else { is_new = 1; earlier = rk < myrank; }
      }
    }
    const int n_old = __popc(__ballot_sync(0xffffffffu, is_old));
    const int n_new = __popc(__ballot_sync(0xffffffffu, is_new)) + 1;
    const int n_before = __popc(__ballot_sync(0xffffffffu, earlier));
    if (lane == 0) {
      L1Meta* mt = &M.l1_meta[s1];
      if (n_old + n_before < 27) mt->child[n_old + n_before] = (uint8_t)mycode;
      if (n_before == 0) mt->nchild = (uint8_t)(n_old + n_new);
    }
  }
} };
// ---- surfels ----------------------------------------------------------------------------------------------
// one warp per affected L1 (VoxelMap.cpp:187-261): lane c fetches child c of the child set; lane 0 sums in child-set
// order (f32, as the reference), runs the Jacobi SVD and applies the planarity gate.  Non-planar parents are queued
// for the purge.  The affected-set entry is cleared on the way out (the set is self-cleaning).
struct k_surfel { static __device__ __forceinline__ void run(MapDev M, FEntry* atab, const int* __restrict__ alist, int* us, int* plist, unsigned int* pfirst) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  const int naff = us[US_NAFF];
  const bool skip = (us[US_ERR] & ERR_CAP) || !M.compute_surfels;
  if (!gate_v) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int a = warp; a < naff; a += nwarps) {
    const int h = alist[a];
    const unsigned long long ak = atab[h].key;
    const unsigned int first = atab[h].first;
    __syncwarp();
    if (lane == 0) { atab[h].key = KEY_EMPTY; atab[h].first = 0xFFFFFFFFu; atab[h].cnt = -1; }
    if (skip) continue;
    int s1 = l1_find(M, ak);
    if (s1 < 0) continue;
    L1Meta* mt = &M.l1_meta[s1];
    const unsigned long long k1 = M.l1_tab[s1].key;
    const int N = mt->nchild;
    if (N < 5) { if (lane == 0) M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    if ((k1 & SURFEL_BIT) && mt->last_child_count == N) continue;  // incremental skip (VoxelMap.cpp:202-205)
    int px, py, pz;
    key_unpack(k1 & KEY_MASK, px, py, pz);
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    int have = 0;
    if (lane < N) {
      int s0 = l0_find(M, child_key(px, py, pz, M.factor, mt->child[lane]));
      if (s0 >= 0) { float4 c = M.l0_cent[M.l0_tab[s0].pos]; cx = c.x; cy = c.y; cz = c.z; have = 1; }
    }
    const unsigned hm = __ballot_sync(0xffffffffu, have);
    float cents[27 * 3]; int nc = 0;
    for (int c = 0; c < N; ++c) {  // uniform loop: every lane takes part in the shuffles
      float vx = __shfl_sync(0xffffffffu, cx, c), vy = __shfl_sync(0xffffffffu, cy, c), vz = __shfl_sync(0xffffffffu, cz, c);
      if ((hm >> c) & 1u) { cents[nc * 3] = vx; cents[nc * 3 + 1] = vy; cents[nc * 3 + 2] = vz; ++nc; }
    }
    if (lane != 0) continue;
    if (nc < 3) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    float mu[3], nrm[3], plan;
    fit_plane(cents, nc, mu, nrm, &plan);
    if (plan > M.planarity_thr) {  // not planar: the parent and all its children go (VoxelMap.cpp:244-253)
      int idx = atomicAdd(&us[US_NPURGE], 1);
      plist[idx] = s1; pfirst[idx] = first;
      continue;
    }
    L1Entry* e = &M.l1_tab[s1];
    e->n[0] = nrm[0]; e->n[1] = nrm[1]; e->n[2] = nrm[2];
    e->c[0] = mu[0]; e->c[1] = mu[1]; e->c[2] = mu[2];
    mt->planarity = plan; mt->last_child_count = N;
    __threadfence();
    e->key = k1 | SURFEL_BIT;
  }
} };

// The same refit in TWO kernels for lock-step batches: a warp gathers the child centroids of an affected L1 (lane c fetches child c,
// as above) into a scratch record, then ONE THREAD per affected L1 runs the sums, the Jacobi SVD and the planarity gate.  With S x ~900
// refits per step a warp per refit would leave 31 lanes waiting for lane 0's SVD (in-graph timeline at S = 128: 312 us per step in one
// kernel; a thread-per-L1 kernel that also walks the children serially: 231 us).  Same arithmetic in the same order: identical bits.
constexpr int SURFEL_REC = 88;   // floats per scratch record: s1, nc, first, N, then <= 27 centroids
// Four affected L1s per warp (eight lanes each: lane c of a group fetches children c, c + 8, ... - most voxels have <= 8 children, so a
// warp per voxel left three quarters of the lanes idle and, the chain being seven dependent memory round trips, a quarter of the
// possible requests in flight; in-graph timeline at S = 128: 245 -> see DESIGN.md).
struct k_surfel_gather { static __device__ __forceinline__ void run(MapDev M, FEntry* atab, const int* __restrict__ alist, int* us, float* rec) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  const int naff = us[US_NAFF];
  const bool skip = (us[US_ERR] & ERR_CAP) || !M.compute_surfels;
  if (!gate_v) return;
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int a0 = warp * 4; a0 < naff; a0 += nwarps * 4) {   // warp-uniform trip count
    const int a = a0 + g;
    bool go = a < naff;
    float* r = rec + (size_t)(go ? a : 0) * SURFEL_REC;
    unsigned long long ak = 0; unsigned int first = 0; int h = 0;
    if (go) { h = alist[a]; ak = atab[h].key; first = atab[h].first; }
    __syncwarp();
    if (go && gl == 0) { atab[h].key = KEY_EMPTY; atab[h].first = 0xFFFFFFFFu; atab[h].cnt = -1; r[0] = __int_as_float(-1); }
    go = go && !skip;
    int s1 = go ? l1_find(M, ak) : -1;
    go = go && s1 >= 0;
    unsigned long long k1 = 0; int N = 0; L1Meta* mt = nullptr;
    if (go) {
      mt = &M.l1_meta[s1];
      k1 = M.l1_tab[s1].key;
      N = mt->nchild;
      if (N < 5) { if (gl == 0) M.l1_tab[s1].key = k1 & ~SURFEL_BIT; go = false; }
      else if ((k1 & SURFEL_BIT) && mt->last_child_count == N) go = false;  // incremental skip (VoxelMap.cpp:202-205)
    }
    int px = 0, py = 0, pz = 0;
    if (go) key_unpack(k1 & KEY_MASK, px, py, pz);
    int q0 = 0;   // present children of this group's voxel so far
    for (int c0 = 0; c0 < 27; c0 += 8) {
      const int c = c0 + gl;
      const bool mine = go && c < N;
      if (!__any_sync(0xffffffffu, mine)) break;
      float cx = 0.0f, cy = 0.0f, cz = 0.0f;
      int have = 0;
      if (mine) {
        int s0 = l0_find(M, child_key(px, py, pz, M.factor, mt->child[c]));
        if (s0 >= 0) { float4 cc = M.l0_cent[M.l0_tab[s0].pos]; cx = cc.x; cy = cc.y; cz = cc.z; have = 1; }
      }
      const unsigned hm = (__ballot_sync(0xffffffffu, have) >> (g * 8)) & 0xffu;
      if (have) { const int q = q0 + __popc(hm & ((1u << gl) - 1u)); r[4 + 3 * q] = cx; r[5 + 3 * q] = cy; r[6 + 3 * q] = cz; }   // child-set order, present children only
      q0 += __popc(hm);
    }
    if (go && gl == 0) { r[1] = __int_as_float(q0); r[2] = __uint_as_float(first); r[3] = __int_as_float(N); r[0] = __int_as_float(s1); }
  }
} };
struct k_surfel_fit { static __device__ __forceinline__ void run(MapDev M, int* us, const float* __restrict__ rec, int* plist, unsigned int* pfirst) { TL_START();
  const int gate_v = M.gate ? *M.gate : 1;
  const int naff = us[US_NAFF];
  if (!gate_v) return;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < naff; a += gridDim.x * blockDim.x) {
    const float* r = rec + (size_t)a * SURFEL_REC;
    const int s1 = __float_as_int(r[0]);
    if (s1 < 0) continue;
    const int nc = __float_as_int(r[1]), N = __float_as_int(r[3]);
    const unsigned int first = __float_as_uint(r[2]);
    L1Meta* mt = &M.l1_meta[s1];
    const unsigned long long k1 = M.l1_tab[s1].key;
    if (nc < 3) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    float cents[27 * 3];
    for (int c = 0; c < nc * 3; ++c) cents[c] = r[4 + c];
    float mu[3], nrm[3], plan;
    fit_plane(cents, nc, mu, nrm, &plan);
    if (plan > M.planarity_thr) {  // not planar: the parent and all its children go (VoxelMap.cpp:244-253)
      int idx = atomicAdd(&us[US_NPURGE], 1);
      plist[idx] = s1; pfirst[idx] = first;
      continue;
    }
    L1Entry* e = &M.l1_tab[s1];
    e->n[0] = nrm[0]; e->n[1] = nrm[1]; e->n[2] = nrm[2];
    e->c[0] = mu[0]; e->c[1] = mu[1]; e->c[2] = mu[2];
    mt->planarity = plan; mt->last_child_count = N;
    __threadfence();
    e->key = k1 | SURFEL_BIT;
  }
} };

// RecomputeAllSurfels (VoxelMap.cpp:304-366): every L1; non-planar parents only lose the surfel (no purge)
__global__ void k_surfel_all(MapDev M) { TL_START();
  const int tcap = 1 << M.l1_log2cap;
  for (int s1 = blockIdx.x * blockDim.x + threadIdx.x; s1 < tcap; s1 += gridDim.x * blockDim.x) {
    unsigned long long k1 = M.l1_tab[s1].key;
    if (k1 == KEY_EMPTY || k1 == KEY_TOMB) continue;
    L1Meta* mt = &M.l1_meta[s1];
    int N = mt->nchild;
    if (N < 5) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    int px, py, pz;
    key_unpack(k1 & KEY_MASK, px, py, pz);
    float cents[27 * 3]; int nc = 0;
    for (int ci = 0; ci < N; ++ci) {
      int s0 = l0_find(M, child_key(px, py, pz, M.factor, mt->child[ci]));
      if (s0 < 0) continue;
      float4 c = M.l0_cent[M.l0_tab[s0].pos];
      cents[nc * 3] = c.x; cents[nc * 3 + 1] = c.y; cents[nc * 3 + 2] = c.z; ++nc;
    }
    if (nc < 5) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    float mu[3], nrm[3], plan;
    fit_plane(cents, nc, mu, nrm, &plan);
    if (plan > M.planarity_thr) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    L1Entry* e = &M.l1_tab[s1];
    e->n[0] = nrm[0]; e->n[1] = nrm[1]; e->n[2] = nrm[2];
    e->c[0] = mu[0]; e->c[1] = mu[1]; e->c[2] = mu[2];
    mt->planarity = plan; mt->last_child_count = N;
    e->key = k1 | SURFEL_BIT;
  }
}
// new_centroid = R * centroid + t (VoxelMap.cpp:273-276), point_count carried in w
struct Rt12 { float R[9]; float t[3]; };
__global__ void k_xform_l0(MapDev M, int n0, Rt12 T, float4* out, int* d_n) { TL_START();
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
    float4 c = M.l0_cent[pos];
    float v[3] = {c.x, c.y, c.z}, r[3];
    mat3_vec(T.R, v, r);
    out[pos] = make_float4(r[0] + T.t[0], r[1] + T.t[1], r[2] + T.t[2], c.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *d_n = n0;
}

// one CTA closes the update: order the purged parents by the position of their key in affected_L1 (= first touching
// point), lay out the erase sequence (children in child-set order, one warp per parent), replay the swap-erases on
// indices, apply the moves, publish the counters.
struct k_upd_close { static __device__ __forceinline__ void run(MapDev M, int* us, const int* plist, const unsigned int* pfirst, int* pord, int* poff, int* seq_pos,
                                                    int* aux, int purge_ran, const int* __restrict__ d_m, const int* __restrict__ pslot,
                                                    const int* __restrict__ isnew, int ranks_cleared) { TL_START();
  if (M.gate && !*M.gate) return;
  extern __shared__ int smem[];
  __shared__ int sm[40];
  __shared__ int s_k;
  if (!ranks_cleared) {  // creation ranks are only meaningful inside one update: back to idle
    const int m = *d_m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) if (isnew[i]) M.l0_tab[pslot[i]].rank = -1;
  }
  const int P = purge_ran ? us[US_NPURGE] : 0;
  const int n_all = us[US_N0] + ((us[US_ERR] & ERR_CAP) ? 0 : us[US_NNEW]);
  if (P == 0) {
    const int err = us[US_ERR];
    __syncthreads();   // everybody has read the update's scalars: the block goes back to all-zero for the next update (no memset node)
    if (threadIdx.x == 0) { M.ctr[CT_N0] = n_all; M.ctr[CT_ERR] = err; M.ctr[6] = 0; M.ctr[7] = 0; }
    if (threadIdx.x < US_COUNT) us[threadIdx.x] = 0;
    return;
  }
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    unsigned int f = pfirst[t]; int r = 0;
    for (int u = 0; u < P; ++u) r += (pfirst[u] < f);
    pord[r] = plist[t];
  }
  __syncthreads();
  int base = 0;
  for (int t0 = 0; t0 < P; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    int c = t < P ? (int)M.l1_meta[pord[t]].nchild : 0, tot;
    int e = block_excl_scan(c, &tot, sm);
    if (t < P) poff[t] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) s_k = base;
  __syncthreads();
  const int k = s_k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int t = warp; t < P; t += nwarps) {
    int s1 = pord[t];
    L1Meta* mt = &M.l1_meta[s1];
    int px, py, pz;
    key_unpack(M.l1_tab[s1].key & KEY_MASK, px, py, pz);
    int N = mt->nchild, o = poff[t];
    if (lane < N) {
      int s0 = l0_find(M, child_key(px, py, pz, M.factor, mt->child[lane]));
      int pos = -1;
      if (s0 >= 0) { pos = (int)M.l0_tab[s0].pos; M.l0_tab[s0].key = KEY_TOMB; }
      else atomicOr(&us[US_ERR], ERR_INTERNAL);
      seq_pos[o + lane] = pos;
    }
    __syncwarp();
    if (lane == 0) {
      M.l1_tab[s1].key = KEY_TOMB;
      mt->nchild = 0; mt->last_child_count = 0;
      atomicSub(&M.ctr[CT_N1], 1); atomicAdd(&M.ctr[CT_TOMB1], 1);
    }
  }
  __syncthreads();
  swap_erase_sim(k, n_all, seq_pos, aux, smem, SIM_SMEM_K);
  swap_erase_apply(M, k, n_all - k, seq_pos, aux);
  if (threadIdx.x == 0) {
    M.ctr[CT_N0] = n_all - k;
    atomicAdd(&M.ctr[CT_TOMB0], k);
    M.ctr[CT_ERR] = us[US_ERR];
    M.ctr[6] = k; M.ctr[7] = P;   // purge size of this update (debug / statistics)
  }
  __syncthreads();
  if (threadIdx.x < US_COUNT) us[threadIdx.x] = 0;   // self-cleaning state block (see the P == 0 exit)
} };

// ---- table maintenance --------------------------------------------------------------------------------------
__global__ void k_l0_reinsert(MapDev M, int n0) { TL_START();
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
    bool ins;
    int s0 = l0_find_or_insert(M, M.l0_key[pos], &ins);
    if (s0 >= 0) { M.l0_tab[s0].pos = (uint32_t)pos; M.l0_slot[pos] = (uint32_t)s0; }
  }
}
__global__ void k_l1_reinsert(MapDev Mnew, const L1Entry* oldtab, const L1Meta* oldmeta, int oldcap) { TL_START();
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < oldcap; s += gridDim.x * blockDim.x) {
    unsigned long long k = oldtab[s].key;
    if (k == KEY_EMPTY || k == KEY_TOMB) continue;
    bool ins;
    int ns = l1_find_or_insert(Mnew, k & KEY_MASK, &ins);
    if (ns < 0) continue;
    L1Entry e = oldtab[s];
    Mnew.l1_tab[ns].n[0] = e.n[0]; Mnew.l1_tab[ns].n[1] = e.n[1]; Mnew.l1_tab[ns].n[2] = e.n[2];
    Mnew.l1_tab[ns].c[0] = e.c[0]; Mnew.l1_tab[ns].c[1] = e.c[1]; Mnew.l1_tab[ns].c[2] = e.c[2];
    Mnew.l1_meta[ns] = oldmeta[s];
    Mnew.l1_tab[ns].key = k;
  }
}
__global__ void k_fill_int(int* p, size_t n, int v) { TL_START();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- exports ----------------------------------------------------------------------------------------------------
__global__ void k_export_l0(MapDev M, int n0, float* xyz, int* keys, int* counts) { TL_START();
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
    float4 c = M.l0_cent[pos];
    xyz[pos * 3] = c.x; xyz[pos * 3 + 1] = c.y; xyz[pos * 3 + 2] = c.z;
    if (keys) { int x, y, z; key_unpack(M.l0_key[pos], x, y, z); keys[pos * 3] = x; keys[pos * 3 + 1] = y; keys[pos * 3 + 2] = z; }
    if (counts) counts[pos] = __float_as_int(c.w);
  }
}
__global__ void k_export_l1(MapDev M, int* counter, int cap, int surfels_only, int* keys, int* nchild, int* children, int* has, float* normal,
                            float* centroid, float* planarity, int* last) { TL_START();
  const int tcap = 1 << M.l1_log2cap;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < tcap; s += gridDim.x * blockDim.x) {
    unsigned long long k = M.l1_tab[s].key;
    if (k == KEY_EMPTY || k == KEY_TOMB) continue;
    if (surfels_only && !(k & SURFEL_BIT)) continue;
    int o = atomicAdd(counter, 1);
    if (o >= cap) continue;
    const L1Meta* mt = &M.l1_meta[s];
    int px, py, pz;
    key_unpack(k & KEY_MASK, px, py, pz);
    if (keys) { keys[o * 3] = px; keys[o * 3 + 1] = py; keys[o * 3 + 2] = pz; }
    if (nchild) nchild[o] = mt->nchild;
    if (children) for (int ci = 0; ci < 27; ++ci) {
      int code = ci < mt->nchild ? mt->child[ci] : 0;
      children[(o * 27 + ci) * 3] = ci < mt->nchild ? px * M.factor + code % 3 : 0;
      children[(o * 27 + ci) * 3 + 1] = ci < mt->nchild ? py * M.factor + (code / 3) % 3 : 0;
      children[(o * 27 + ci) * 3 + 2] = ci < mt->nchild ? pz * M.factor + code / 9 : 0;
    }
    if (has) has[o] = (k & SURFEL_BIT) ? 1 : 0;
    if (normal) for (int a = 0; a < 3; ++a) normal[o * 3 + a] = M.l1_tab[s].n[a];
    if (centroid) for (int a = 0; a < 3; ++a) centroid[o * 3 + a] = M.l1_tab[s].c[a];
    if (planarity) planarity[o] = mt->planarity;
    if (last) last[o] = mt->last_child_count;
  }
}
__global__ void k_count_surfels(MapDev M, int* out) { TL_START();
  const int tcap = 1 << M.l1_log2cap;
  int c = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < tcap; s += gridDim.x * blockDim.x) {
    unsigned long long k = M.l1_tab[s].key;
    if (k != KEY_EMPTY && k != KEY_TOMB && (k & SURFEL_BIT)) ++c;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---- host side ------------------------------------------------------------------------------------------------------
static int fill_int(cudaStream_t st, int* p, size_t n, int v) {
  if (!n) return 0;
  int blocks = (int)((n + 1023) / 1024); if (blocks > 2368) blocks = 2368;
  k_fill_int<<<blocks, 1024, 0, st>>>(p, n, v);
  return 1;
}
template <class T> static int dmalloc(T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu B) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return B2LO_E_NOMEM; }
  return B2LO_OK;
}
static int ceil_log2(size_t v) { int l = 4; while ((1ull << l) < v) ++l; return l; }

static int alloc_l0_table(b2lo_map* m, int log2cap) {
  MapDev& d = m->d;
  int rc;
  size_t cap = 1ull << log2cap;
  // same size (a rebuild that only sheds tombstones): the table is refilled from the dense key vector, so it is cleared in place -
  // no cudaFree (a device-wide synchronisation that would stall every other sequence on the GPU) and no cudaMalloc
  if (!(d.l0_tab && log2cap == d.l0_log2cap && m->tcap0 == cap) && (rc = dmalloc(&d.l0_tab, cap))) return rc;
  d.l0_log2cap = log2cap; m->tcap0 = cap;
  cudaStream_t st = m->ctx->stream;
  B2_CUDA(cudaMemsetAsync(d.l0_tab, 0xFF, cap * sizeof(L0Entry), st));  // key EMPTY, pos PENDING, scratch idle
  return B2LO_OK;
}
static int alloc_l1_table(b2lo_map* m, int log2cap) {
  MapDev& d = m->d;
  int rc;
  size_t cap = 1ull << log2cap;
  if ((rc = dmalloc(&d.l1_tab, cap))) return rc;
  if ((rc = dmalloc(&d.l1_meta, cap))) return rc;
  d.l1_log2cap = log2cap; m->tcap1 = cap;
  cudaStream_t st = m->ctx->stream;
  B2_CUDA(cudaMemsetAsync(d.l1_tab, 0xFF, cap * sizeof(L1Entry), st));
  B2_CUDA(cudaMemsetAsync(d.l1_meta, 0, cap * sizeof(L1Meta), st));
  return B2LO_OK;
}

// take over counters the caller has already copied into ctx->h_counts[0..8) and synchronised
int map_absorb_counts(b2lo_map* m) {
  b2lo_ctx* ctx = m->ctx;
  m->n0 = (size_t)ctx->h_counts[CT_N0]; m->n1 = (size_t)ctx->h_counts[CT_N1];
  m->tomb0 = (size_t)ctx->h_counts[CT_TOMB0]; m->tomb1 = (size_t)ctx->h_counts[CT_TOMB1];
  int err = ctx->h_counts[CT_ERR];
  if (err & ERR_CAP) { set_error("voxel map capacity exceeded during update"); return B2LO_E_CAPACITY; }
  if (err & ERR_INTERNAL) { set_error("voxel map internal inconsistency (child without L0 voxel)"); return B2LO_E_CAPACITY; }
  if (err & ERR_RANGE) { set_error("point outside the 21-bit voxel key domain was skipped"); return B2LO_E_RANGE; }
  return B2LO_OK;
}

int map_refresh_counts(b2lo_map* m) {
  b2lo_ctx* ctx = m->ctx;
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts, m->d.ctr, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  m->n0 = (size_t)ctx->h_counts[CT_N0]; m->n1 = (size_t)ctx->h_counts[CT_N1];
  m->tomb0 = (size_t)ctx->h_counts[CT_TOMB0]; m->tomb1 = (size_t)ctx->h_counts[CT_TOMB1];
  return B2LO_OK;
}

// make room for `need_l0` dense voxels and an update of `need_upd` points; rebuild tables that are too full.
// Host counts (n0, n1, tombstones) must be current.
int map_reserve(b2lo_map* m, size_t need_l0, size_t need_upd) {
  b2lo_ctx* ctx = m->ctx;
  MapDev& d = m->d;
  cudaStream_t st = ctx->stream;
  int rc;
  if (need_l0 > d.l0_cap) {  // grow the dense vectors (copy live prefix)
    m->alloc_epoch++;
    size_t ncap = d.l0_cap ? d.l0_cap : 1024;
    while (ncap < need_l0) ncap *= 2;
    float4* nc = nullptr; unsigned long long* nk = nullptr; uint32_t* ns = nullptr;
    if ((rc = dmalloc(&nc, ncap)) || (rc = dmalloc(&nk, ncap)) || (rc = dmalloc(&ns, ncap))) return rc;
    if (m->n0) {
      B2_CUDA(cudaMemcpyAsync(nc, d.l0_cent, m->n0 * sizeof(float4), cudaMemcpyDeviceToDevice, st));
      B2_CUDA(cudaMemcpyAsync(nk, d.l0_key, m->n0 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
      B2_CUDA(cudaMemcpyAsync(ns, d.l0_slot, m->n0 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    B2_CUDA(cudaStreamSynchronize(st));
    cudaFree(d.l0_cent); cudaFree(d.l0_key); cudaFree(d.l0_slot);
    d.l0_cent = nc; d.l0_key = nk; d.l0_slot = ns; d.l0_cap = (uint32_t)ncap;
    if ((rc = dmalloc(&m->c_flag, ncap)) || (rc = dmalloc(&m->c_removed, ncap)) || (rc = dmalloc(&m->c_aux, 4 * ncap + 16)) ||
        (rc = dmalloc(&m->c_l1work, ncap)) || (rc = dmalloc(&m->c_blkcnt, ncap / 1024 + 2)) || (rc = dmalloc(&m->c_blkoff, ncap / 1024 + 2)))
      return rc;
  }
  // L0 hash: keep (live + tombstones + incoming) under half the table
  if ((m->n0 + m->tomb0 + need_upd) * 2 > m->tcap0) {
    int l2 = ceil_log2((m->n0 + need_upd) * 4);
    if (l2 < d.l0_log2cap) l2 = d.l0_log2cap;
    if (!(d.l0_tab && l2 == d.l0_log2cap)) {   // the table moves: the old one must be idle before it is freed (and graphs must be rebuilt)
      m->alloc_epoch++;
      B2_CUDA(cudaStreamSynchronize(st));
    }
    const bool moved = !(d.l0_tab && l2 == d.l0_log2cap);
    if ((rc = alloc_l0_table(m, l2))) return rc;
    if (m->n0) { int blocks = (int)((m->n0 + 255) / 256); if (blocks > 2368) blocks = 2368; k_l0_reinsert<<<blocks, 256, 0, st>>>(d, (int)m->n0); ctx->launches++; }
    B2_CUDA(cudaMemsetAsync(d.ctr + CT_TOMB0, 0, sizeof(int), st));
    // an in-place rebuild leaves the pointers (and alloc_epoch) alone, so nobody downstream knows that this stream carries work: a
    // lock-step batch runs the sequence's kernels on the BATCH's stream.  Wait here (tens of microseconds, once in dozens of scans).
    if (!moved) B2_CUDA(cudaStreamSynchronize(st));
    m->tomb0 = 0;
  }
  if ((m->n1 + m->tomb1 + need_upd) * 2 > m->tcap1) {
    m->alloc_epoch++;
    int l2 = ceil_log2((m->n1 + need_upd) * 4);
    if (l2 < d.l1_log2cap) l2 = d.l1_log2cap;
    L1Entry* oldtab = d.l1_tab; L1Meta* oldmeta = d.l1_meta; size_t oldcap = m->tcap1;
    // a rebuild that only sheds tombstones ping-pongs between two buffers of the same size: the table left behind by the previous
    // rebuild is the target of this one (no cudaMalloc / cudaFree - the latter synchronises the whole device - in steady state)
    const bool reuse = m->l1_spare_tab && m->l1_spare_cap == ((size_t)1 << l2);
    d.l1_tab = reuse ? m->l1_spare_tab : nullptr; d.l1_meta = reuse ? m->l1_spare_meta : nullptr;
    if (!reuse && m->l1_spare_tab) { cudaFree(m->l1_spare_tab); cudaFree(m->l1_spare_meta); }
    m->l1_spare_tab = nullptr; m->l1_spare_meta = nullptr; m->l1_spare_cap = 0;
    if (reuse) {
      d.l1_log2cap = l2; m->tcap1 = (size_t)1 << l2;
      B2_CUDA(cudaMemsetAsync(d.l1_tab, 0xFF, m->tcap1 * sizeof(L1Entry), st));
      B2_CUDA(cudaMemsetAsync(d.l1_meta, 0, m->tcap1 * sizeof(L1Meta), st));
    } else if ((rc = alloc_l1_table(m, l2))) return rc;
    if (oldtab && m->n1) { int blocks = (int)((oldcap + 255) / 256); if (blocks > 2368) blocks = 2368; k_l1_reinsert<<<blocks, 256, 0, st>>>(d, oldtab, oldmeta, (int)oldcap); ctx->launches++; }
    B2_CUDA(cudaMemsetAsync(d.ctr + CT_TOMB1, 0, sizeof(int), st));
    if (oldtab && oldcap == m->tcap1) { m->l1_spare_tab = oldtab; m->l1_spare_meta = oldmeta; m->l1_spare_cap = oldcap; }   // stream-ordered: reused only by a later rebuild
    else {
      B2_CUDA(cudaStreamSynchronize(st));
      if (oldtab) cudaFree(oldtab);
      if (oldmeta) cudaFree(oldmeta);
    }
    m->tomb1 = 0;
  }
  if (need_upd > m->upd_cap) {
    m->alloc_epoch++;
    size_t ncap = m->upd_cap ? m->upd_cap : 4096;
    while (ncap < need_upd) ncap *= 2;
    B2_CUDA(cudaStreamSynchronize(st));
    if ((rc = dmalloc(&m->u_pts, ncap)) || (rc = dmalloc(&m->u_pslot, ncap)) || (rc = dmalloc(&m->u_next, ncap)) || (rc = dmalloc(&m->u_isnew, ncap)) ||
        (rc = dmalloc(&m->u_newrank, 2 * ncap)) || (rc = dmalloc(&m->u_part, ncap / INS_CHUNK + 2)))
      return rc;
    m->a_log2cap = ceil_log2(ncap * 2);
    if ((rc = dmalloc(&m->a_tab, (size_t)1 << m->a_log2cap)) || (rc = dmalloc(&m->a_list, ncap * 4 + 16)) || (rc = dmalloc(&m->a_slots, ncap + 16))) return rc;
    B2_CUDA(cudaMemsetAsync(m->a_tab, 0xFF, sizeof(FEntry) << m->a_log2cap, st));  // the affected set cleans itself afterwards (k_surfel)
    // purge scratch: up to 27 children per affected parent
    if ((rc = dmalloc(&m->s_rec, ncap * (size_t)SURFEL_REC))) return rc;   // surfel refit records of the lock-step path (k_surfel_gather / k_surfel_fit)
    m->p_cap = ncap * 27;
    if ((rc = dmalloc(&m->p_seq, m->p_cap + 16)) || (rc = dmalloc(&m->p_aux, m->p_cap * 4 + 16))) return rc;
    m->upd_cap = ncap;
  }
  return B2LO_OK;
}

static int grid_for(size_t n, int threads) { size_t b = (n + threads - 1) / threads; if (b < 1) b = 1; if (b > 1184) b = 1184; return (int)b; }

// UpdateVoxelMap on a world-frame cloud already on the device (float4 stream).  n_cap = host-known upper
// bound of *d_n.  Ends with a counter read-back (one synchronisation).
int map_update_dev(b2lo_map* m, float4* d_world, const int* d_n, size_t n_cap, const float sensor[3], float radius_sq, int rehash,
                   const int* gate, const float* sensor_dev, const float4* local, const float* T16_dev) {
  b2lo_ctx* ctx = m->ctx;
  if (n_cap == 0) return B2LO_S_EMPTY;
  int rc = map_reserve(m, m->n0 + n_cap, n_cap);
  if (rc) return rc;
  m->d.gate = gate; m->d.sensor_dev = sensor_dev;   // passed by value with every launch below
  MapDev d = m->d;
  m->d.gate = nullptr; m->d.sensor_dev = nullptr;
  cudaStream_t st = ctx->stream;
  int* us = m->u_state;
  // (k_cull_fix / k_upd_close take SIM_SMEM_BYTES of dynamic shared memory: launch<>() raises their limit on first use)
  prof_begin(ctx, PS_MAP);
  // the per-update scalar block `us` is all-zero here: zeroed at creation / Clear and again by the closing kernel of every update
  const int n0 = m->graph_mode ? -1 : (int)m->n0;
  if ((n0 > 0 || m->graph_mode) && !rehash) {
    int tiles = (int)(((m->graph_mode ? (size_t)d.l0_cap : m->n0) + 1023) / 1024);
    int g = tiles > ctx->sm_count ? ctx->sm_count : tiles;
    int g4 = (tiles + 3) / 4; if (g4 > 2 * ctx->sm_count) g4 = 2 * ctx->sm_count;
    g = batch_grid(ctx, g); g4 = batch_grid(ctx, g4);
    prof_end(ctx);
    prof_begin(ctx, PS_CULL);
    launch<k_cull_mark, 1024, 1>(ctx, dim3((unsigned)(g4)), dim3((unsigned)(1024)), 0, st, d, n0, sensor[0], sensor[1], sensor[2], radius_sq, m->c_flag, m->c_blkcnt, m->c_blkoff, us);
    prof_end(ctx);
    prof_begin(ctx, PS_MAP);
    launch<k_cull_lists, 1024, 1>(ctx, dim3((unsigned)(g)), dim3((unsigned)(1024)), 0, st, d, n0, m->c_flag, m->c_blkoff, us, m->c_removed, m->c_l1work);
    launch<k_cull_fix, 1024, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(1024)), SIM_SMEM_BYTES, st, d, m->c_flag, us, m->c_l1work, m->c_removed, m->c_aux, n0);
    ctx->launches += 3;
  }
  int gm = batch_grid(ctx, grid_for(n_cap, 256));
  int gw = batch_grid(ctx, grid_for(n_cap * 32, 256));   // one warp per point
  int* plist = m->a_list; unsigned int* pfirst = reinterpret_cast<unsigned int*>(m->a_list + m->upd_cap);
  int* pord = m->a_list + 2 * m->upd_cap; int* poff = m->a_list + 3 * m->upd_cap;
  launch<k_ins_probe, 256, 1, 5>(ctx, dim3((unsigned)(gm)), dim3((unsigned)(256)), 0, st, d, d_world, d_n, us, m->u_pslot, m->u_next, m->a_tab, m->a_log2cap, m->a_slots, local, T16_dev);
  launch<k_ins_apply, 256, 1>(ctx, dim3((unsigned)(gm)), dim3((unsigned)(256)), 0, st, d, d_world, d_n, m->u_pslot, m->u_next, m->u_isnew, m->u_pts, rehash);
  const bool bulk = n_cap > BULK_UPD && !gate;
  if (bulk) {
    int gc = (int)((n_cap + INS_CHUNK - 1) / INS_CHUNK); if (gc > 4 * ctx->sm_count) gc = 4 * ctx->sm_count;
    k_ins_scan_part<<<gc, 1024, 0, st>>>(d_n, m->u_isnew, m->u_part);
    k_ins_scan_top<<<1, 1024, 0, st>>>(d, d_n, m->u_part, us);
    k_ins_scan_apply<<<gc, 1024, 0, st>>>(d, d_n, m->u_isnew, m->u_pslot, m->u_newrank, m->u_part, m->u_newrank + m->upd_cap);
    ctx->launches += 2;
  } else {
    launch<k_ins_scan, 1024, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(1024)), 0, st, d, d_n, m->u_isnew, m->u_pslot, m->u_newrank, us, m->u_newrank + m->upd_cap);
  }
  launch<k_ins_place, 256, 1, 5>(ctx, dim3((unsigned)(gw)), dim3((unsigned)(256)), 0, st, d, d_n, us, m->u_pslot, m->u_isnew, m->u_newrank, m->u_pts, m->u_newrank + m->upd_cap);
  ctx->launches += 4;
  int purge = 0;
  if (rehash) {
    // ApplyTransformAndRehash always ends in RecomputeAllSurfels (VoxelMap.cpp:301), whatever compute_surfels says;
    // the affected set is still drained (skip = !compute_surfels is overridden by passing through k_surfel with surfels off)
    MapDev dd = d; dd.compute_surfels = 0;
    launch<k_surfel, 256, 1>(ctx, dim3((unsigned)(gw)), dim3((unsigned)(256)), 0, st, dd, m->a_tab, m->a_slots, us, plist, pfirst);
    k_surfel_all<<<grid_for(m->tcap1, 128), 128, 0, st>>>(d);
    ctx->launches += 2;
  } else {
    purge = d.compute_surfels;
    if (ctx->batch_S > 0 && m->s_rec) {   // lock-step batches: gather by warps, fit by threads (see k_surfel_gather)
      launch<k_surfel_gather, 256, 1, 6>(ctx, dim3((unsigned)(gw)), dim3(256u), 0, st, d, m->a_tab, m->a_slots, us, m->s_rec);
      launch<k_surfel_fit, 128, 1>(ctx, dim3((unsigned)(batch_grid(ctx, grid_for(n_cap, 128)))), dim3(128u), 0, st, d, us, (const float*)m->s_rec, plist, pfirst);
      ctx->launches += 1;
    }
    else launch<k_surfel, 256, 1>(ctx, dim3((unsigned)(gw)), dim3((unsigned)(256)), 0, st, d, m->a_tab, m->a_slots, us, plist, pfirst);
    ctx->launches += 1;
  }
  if (bulk) { k_rank_clear<<<gm, 256, 0, st>>>(d, d_n, m->u_pslot, m->u_isnew); ctx->launches++; }
  launch<k_upd_close, 1024, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(1024)), SIM_SMEM_BYTES, st, d, us, plist, pfirst, pord, poff, m->p_seq, m->p_aux, purge, d_n, m->u_pslot, m->u_isnew, bulk ? 1 : 0);
  prof_end(ctx);
  ctx->launches += 1;
  B2_CUDA(cudaGetLastError());
  if (gate) return B2LO_OK;   // speculative: the caller reads the counters back with its own batch (map_absorb_counts)
  rc = map_refresh_counts(m);
  if (rc) return rc;
  int err = ctx->h_counts[CT_ERR];
  if (err & ERR_CAP) { set_error("voxel map capacity exceeded during update"); return B2LO_E_CAPACITY; }
  if (err & ERR_INTERNAL) { set_error("voxel map internal inconsistency (child without L0 voxel)"); return B2LO_E_CAPACITY; }
  if (err & ERR_RANGE) { set_error("point outside the 21-bit voxel key domain was skipped"); return B2LO_E_RANGE; }
  return B2LO_OK;
}

}  // namespace b2

// ======================================================================================================================
using namespace b2;

extern "C" int b2lo_map_create(b2lo_ctx* ctx, float voxel_size, int hierarchy_factor, float planarity_threshold, int compute_surfels,
                               size_t l0_capacity_hint, b2lo_map** out) {
  if (!ctx || !out) return B2LO_E_ARG;
  if (!(voxel_size > 0.0f)) { set_error("Voxel size must be positive"); return B2LO_E_ARG; }              // VoxelMap.cpp:28-30
  if (hierarchy_factor != 3) { set_error("Hierarchy factor must be 3 in this build (positive odd in the reference)"); return B2LO_E_ARG; }
  cudaSetDevice(ctx->device);
  b2lo_map* m = new b2lo_map();
  m->ctx = ctx;
  MapDev& d = m->d;
  d.voxel = voxel_size; d.factor = hierarchy_factor; d.scale1 = voxel_size * (float)hierarchy_factor;     // VoxelMap.cpp:51-52
  d.planarity_thr = planarity_threshold; d.compute_surfels = compute_surfels ? 1 : 0;
  if (cudaMalloc(&d.ctr, 16 * sizeof(int)) != cudaSuccess || cudaMalloc(&m->u_state, US_COUNT * sizeof(int)) != cudaSuccess) { delete m; return B2LO_E_NOMEM; }
  cudaMemsetAsync(d.ctr, 0, 16 * sizeof(int), ctx->stream);
  cudaMemsetAsync(m->u_state, 0, US_COUNT * sizeof(int), ctx->stream);
  {
    // CUDA loads a kernel when it is first used; the table-maintenance kernels first run dozens of scans into a sequence (when
    // tombstones have piled up), where a cold load showed up as a 15-190 ms stall of one UpdateVoxelMap call.  Load them now.
    static bool loaded = false;
    if (!loaded) {
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, k_l0_reinsert); cudaFuncGetAttributes(&fa, k_l1_reinsert); cudaFuncGetAttributes(&fa, k_fill_int);
      cudaGetLastError();
      loaded = true;
    }
  }
  size_t hint = l0_capacity_hint ? l0_capacity_hint : (1u << 16);
  int rc = alloc_l0_table(m, ceil_log2(hint * 4));
  // the L1 table starts at the size its first tombstone-shedding rebuild would pick for a map of `hint` voxels, and the second buffer
  // of that rebuild's ping-pong exists from the start: a sequence never waits for cudaMalloc in the middle of a run
  const int l1_l2 = ceil_log2(hint * 2);
  if (!rc) rc = alloc_l1_table(m, l1_l2);
  if (!rc) {
    if (cudaMalloc((void**)&m->l1_spare_tab, sizeof(L1Entry) << l1_l2) == cudaSuccess && cudaMalloc((void**)&m->l1_spare_meta, sizeof(L1Meta) << l1_l2) == cudaSuccess)
      m->l1_spare_cap = (size_t)1 << l1_l2;
    else { cudaGetLastError(); if (m->l1_spare_tab) cudaFree(m->l1_spare_tab); m->l1_spare_tab = nullptr; m->l1_spare_meta = nullptr; }   // optional: the rebuild allocates on demand
  }
  if (!rc) rc = map_reserve(m, hint, 4096);
  if (rc) { b2lo_map_destroy(m); return rc; }
  cudaStreamSynchronize(ctx->stream);
  *out = m;
  return B2LO_OK;
}

extern "C" int b2lo_map_destroy(b2lo_map* m) {
  if (!m) return B2LO_E_ARG;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  MapDev& d = m->d;
  void* ptrs[] = {d.l0_cent, d.l0_key, d.l0_slot, d.l0_tab, d.l1_tab, d.l1_meta, d.ctr,
                  m->u_pts, m->u_pslot, m->u_next, m->u_isnew, m->u_newrank, m->a_tab, m->a_list, m->a_slots, m->c_flag, m->c_blkcnt,
                  m->c_blkoff, m->c_removed, m->c_aux, m->c_l1work, m->p_seq, m->p_aux, m->u_state, m->u_part, m->r_tmp, m->r_n, m->s_rec};
  if (m->l1_spare_tab) cudaFree(m->l1_spare_tab);
  if (m->l1_spare_meta) cudaFree(m->l1_spare_meta);
  for (void* p : ptrs) if (p) cudaFree(p);
  delete m;
  return B2LO_OK;
}

extern "C" int b2lo_map_clear(b2lo_map* m) {
  if (!m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  cudaSetDevice(m->ctx->device);
  cudaStream_t st = m->ctx->stream;
  B2_CUDA(cudaMemsetAsync(m->d.l0_tab, 0xFF, m->tcap0 * sizeof(L0Entry), st));
  B2_CUDA(cudaMemsetAsync(m->d.l1_tab, 0xFF, m->tcap1 * sizeof(L1Entry), st));
  B2_CUDA(cudaMemsetAsync(m->d.l1_meta, 0, m->tcap1 * sizeof(L1Meta), st));
  B2_CUDA(cudaMemsetAsync(m->d.ctr, 0, 16 * sizeof(int), st));
  B2_CUDA(cudaMemsetAsync(m->u_state, 0, US_COUNT * sizeof(int), st));
  B2_CUDA(cudaStreamSynchronize(st));
  m->n0 = m->n1 = m->tomb0 = m->tomb1 = 0;
  return B2LO_OK;
}
extern "C" int b2lo_map_set_planarity_threshold(b2lo_map* m, float thr) { if (!m) return B2LO_E_ARG; m->d.planarity_thr = thr; return B2LO_OK; }
extern "C" int b2lo_map_set_compute_surfels(b2lo_map* m, int on) { if (!m) return B2LO_E_ARG; m->d.compute_surfels = on ? 1 : 0; return B2LO_OK; }

extern "C" int b2lo_map_update(b2lo_map* m, const float* world_xyz, size_t n, size_t stride_floats, const double sensor[3], double max_distance) {
  if (!m || !sensor) return B2LO_E_ARG;
  if (!world_xyz || n == 0) return B2LO_S_EMPTY;  // VoxelMap.cpp:134-136
  if (stride_floats < 3) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  b2lo_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  static const bool trace = getenv("B2LO_TRACE_UPDATE") != nullptr;   // debug aid: where the stand-alone call spends its wall time
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = trace ? now() : 0.0;
  int rc = ctx_reserve_points(ctx, n);
  if (rc) return rc;
  rc = ctx_stage_h2d(ctx, world_xyz, n, stride_floats, 1, ctx->d_world, ctx->d_nquery);
  if (rc) return rc;
  const double t1 = trace ? now() : 0.0;
  const unsigned long long ep0 = m->alloc_epoch;
  rc = map_reserve(m, m->n0 + n, n);
  if (rc) return rc;
  const double t2 = trace ? now() : 0.0;
  float sf[3] = {(float)sensor[0], (float)sensor[1], (float)sensor[2]};  // sensor_position.cast<float>()  (VoxelMap.cpp:143)
  float r2 = (float)(max_distance * max_distance);                        // (:144)
  rc = map_update_dev(m, ctx->d_world, ctx->d_nquery, n, sf, r2);
  if (trace) {
    const double t3 = now();
    std::fprintf(stderr, "[b2lo map_update] n %zu n0 %zu tomb0 %zu: stage %.1f us, reserve %.1f us (%s), update %.1f us\n", n, m->n0, m->tomb0, t1 - t0, t2 - t1,
                 m->alloc_epoch != ep0 ? "REBUILT" : "-", t3 - t2);
  }
  return rc;
}

extern "C" int b2lo_map_counts(b2lo_map* m, size_t* l0, size_t* l1, size_t* surfels) {
  if (!m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  b2lo_ctx* ctx = m->ctx;
  if (l0) *l0 = m->n0;
  if (l1) *l1 = m->n1;
  if (surfels) {
    cudaSetDevice(ctx->device);
    B2_CUDA(cudaMemsetAsync(m->d.ctr + CT_SURF, 0, sizeof(int), ctx->stream));
    k_count_surfels<<<grid_for(m->tcap1, 256), 256, 0, ctx->stream>>>(m->d, m->d.ctr + CT_SURF);
    ctx->launches++;
    B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 8, m->d.ctr + CT_SURF, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    *surfels = (size_t)ctx->h_counts[8];
  }
  return B2LO_OK;
}

extern "C" int b2lo_map_export_l0(b2lo_map* m, float* xyz, int* keys, int* counts, size_t cap, size_t* n) {
  if (!m || !n) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  b2lo_ctx* ctx = m->ctx;
  *n = m->n0;
  if (m->n0 == 0) return B2LO_OK;
  if (!xyz || cap < m->n0) { set_error("export_l0: buffer too small (%zu < %zu)", cap, m->n0); return B2LO_E_CAPACITY; }
  cudaSetDevice(ctx->device);
  size_t n0 = m->n0;
  float* dxyz = nullptr; int* dk = nullptr; int* dc = nullptr;
  B2_CUDA(cudaMalloc(&dxyz, n0 * 3 * sizeof(float)));
  if (keys) B2_CUDA(cudaMalloc(&dk, n0 * 3 * sizeof(int)));
  if (counts) B2_CUDA(cudaMalloc(&dc, n0 * sizeof(int)));
  k_export_l0<<<grid_for(n0, 256), 256, 0, ctx->stream>>>(m->d, (int)n0, dxyz, dk, dc);
  ctx->launches++;
  B2_CUDA(cudaMemcpyAsync(xyz, dxyz, n0 * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (keys) B2_CUDA(cudaMemcpyAsync(keys, dk, n0 * 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (counts) B2_CUDA(cudaMemcpyAsync(counts, dc, n0 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(dxyz); if (dk) cudaFree(dk); if (dc) cudaFree(dc);
  return B2LO_OK;
}

static int export_l1_common(b2lo_map* m, int surfels_only, int* keys, int* nchild, int* children, int* has, float* normal, float* centroid,
                            float* planarity, int* last, size_t cap, size_t* n) {
  b2lo_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  size_t n1 = m->n1;
  if (cap > n1) cap = n1;
  int *dk = nullptr, *dn = nullptr, *dch = nullptr, *dh = nullptr, *dl = nullptr, *dcount = nullptr;
  float *dnr = nullptr, *dce = nullptr, *dp = nullptr;
  size_t c1 = cap ? cap : 1;
  B2_CUDA(cudaMalloc(&dcount, sizeof(int)));
  B2_CUDA(cudaMemsetAsync(dcount, 0, sizeof(int), ctx->stream));
  if (keys) B2_CUDA(cudaMalloc(&dk, c1 * 3 * sizeof(int)));
  if (nchild) B2_CUDA(cudaMalloc(&dn, c1 * sizeof(int)));
  if (children) B2_CUDA(cudaMalloc(&dch, c1 * 81 * sizeof(int)));
  if (has) B2_CUDA(cudaMalloc(&dh, c1 * sizeof(int)));
  if (last) B2_CUDA(cudaMalloc(&dl, c1 * sizeof(int)));
  if (normal) B2_CUDA(cudaMalloc(&dnr, c1 * 3 * sizeof(float)));
  if (centroid) B2_CUDA(cudaMalloc(&dce, c1 * 3 * sizeof(float)));
  if (planarity) B2_CUDA(cudaMalloc(&dp, c1 * sizeof(float)));
  k_export_l1<<<grid_for(m->tcap1, 256), 256, 0, ctx->stream>>>(m->d, dcount, (int)cap, surfels_only, dk, dn, dch, dh, dnr, dce, dp, dl);
  ctx->launches++;
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 9, dcount, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  size_t got = (size_t)ctx->h_counts[9];
  *n = got;
  size_t cp = got < cap ? got : cap;
  if (cp) {
    if (keys) B2_CUDA(cudaMemcpy(keys, dk, cp * 3 * sizeof(int), cudaMemcpyDeviceToHost));
    if (nchild) B2_CUDA(cudaMemcpy(nchild, dn, cp * sizeof(int), cudaMemcpyDeviceToHost));
    if (children) B2_CUDA(cudaMemcpy(children, dch, cp * 81 * sizeof(int), cudaMemcpyDeviceToHost));
    if (has) B2_CUDA(cudaMemcpy(has, dh, cp * sizeof(int), cudaMemcpyDeviceToHost));
    if (last) B2_CUDA(cudaMemcpy(last, dl, cp * sizeof(int), cudaMemcpyDeviceToHost));
    if (normal) B2_CUDA(cudaMemcpy(normal, dnr, cp * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (centroid) B2_CUDA(cudaMemcpy(centroid, dce, cp * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (planarity) B2_CUDA(cudaMemcpy(planarity, dp, cp * sizeof(float), cudaMemcpyDeviceToHost));
  }
  void* ptrs[] = {dk, dn, dch, dh, dl, dcount, dnr, dce, dp};
  for (void* p : ptrs) if (p) cudaFree(p);
  return (got > cap) ? B2LO_E_CAPACITY : B2LO_OK;
}
extern "C" int b2lo_map_export_l1(b2lo_map* m, int* keys, int* nchild, int* children, int* has_surfel, float* normal, float* centroid,
                                  float* planarity, int* last_child_count, size_t cap, size_t* n) {
  if (!m || !n) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  return export_l1_common(m, 0, keys, nchild, children, has_surfel, normal, centroid, planarity, last_child_count, cap, n);
}
extern "C" int b2lo_map_export_surfels(b2lo_map* m, float* centroid, float* normal, float* planarity, int* l1keys, size_t cap, size_t* n) {
  if (!m || !n) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  return export_l1_common(m, 1, l1keys, nullptr, nullptr, nullptr, normal, centroid, planarity, nullptr, cap, n);
}

// RebuildKdTree (VoxelMap.cpp:420-438).  The engine needs no separate index: the L0 Z-order hash is the
// uniform grid the exact 5-NN search walks (b2lo_knn.cuh), and it is always current.
int b2::map_rebuild_knn_locked(b2lo_map* m) {
  m->knn_ready = m->n0 > 0;
  return B2LO_OK;
}
extern "C" int b2lo_map_rebuild_knn(b2lo_map* m) {
  if (!m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  return map_rebuild_knn_locked(m);
}
extern "C" int b2lo_map_has_knn(b2lo_map* m) { return (m && m->knn_ready) ? 1 : 0; }

// ApplyTransformAndRehash (VoxelMap.cpp:264-302): transform every L0 centroid, clear both levels, re-insert in dense
// order (collisions merged weighted by point_count), rebuild L1, RecomputeAllSurfels (:304-366)
extern "C" int b2lo_map_transform_rehash(b2lo_map* m, const float T16[16]) {
  if (!m || !T16) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  std::lock_guard<std::recursive_mutex> lkc(m->ctx->mu);   // the context's staging buffers / counters / stream (same order as b2lo_icp_optimize)
  b2lo_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t n0 = m->n0;
  if (n0 == 0) return B2LO_S_EMPTY;
  if (n0 > m->r_tmp_cap) {   // kept between calls: pose-graph corrections come in bursts
    B2_CUDA(cudaStreamSynchronize(st));
    if (m->r_tmp) cudaFree(m->r_tmp);
    m->r_tmp = nullptr; m->r_tmp_cap = 0;
    const size_t cap = n0 + n0 / 8 + 1024;
    if (cudaMalloc(&m->r_tmp, cap * sizeof(float4)) != cudaSuccess) { set_error("cudaMalloc(rehash scratch, %zu B) failed", cap * sizeof(float4)); return B2LO_E_NOMEM; }
    m->r_tmp_cap = cap;
  }
  if (!m->r_n) B2_CUDA(cudaMalloc(&m->r_n, sizeof(int)));
  float4* tmp = m->r_tmp; int* d_n = m->r_n;
  Rt12 T;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T.R[i * 3 + j] = T16[i * 4 + j]; T.t[i] = T16[i * 4 + 3]; }
  k_xform_l0<<<grid_for(n0, 256), 256, 0, st>>>(m->d, (int)n0, T, tmp, d_n);
  ctx->launches++;
  int rc = b2lo_map_clear(m);
  if (!rc) {
    float zero[3] = {0, 0, 0};
    rc = map_update_dev(m, tmp, d_n, n0, zero, 0.0f, 1);
  }
  cudaStreamSynchronize(st);
  return rc;
}

#ifdef B2LO_TIMELINE
namespace b2 { int tl_fetch_map(unsigned long long* out, int cap) { return tl_fetch(out, cap); } }
#endif
