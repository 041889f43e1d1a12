// b2lo_dev.cuh — device-side data layout of the voxel map and shared helpers.
//
// HBM layout (all arrays owned by b2lo_map, see b2lo_map.cu):
//   L0 ("leaf voxels", src/database/VoxelMap.h:302-309) — a DENSE SoA vector kept in exactly the order the
//   reference's ankerl::unordered_dense map iterates (insertion order, erase = swap-with-last):
//     l0_cent[pos]  float4  (cx, cy, cz, point_count as int bits)          16 B  <- cull scan / export stream
//     l0_key[pos]   u64     63-bit Z-order code of the voxel key             8 B
//     l0_slot[pos]  u32     index of the voxel's entry in the L0 hash        4 B
//   L0 hash: open addressing, linear probing, 32 B entries = one sector: {u64 key, u32 pos, per-update scratch
//            (first point, point count - 1, list head, creation rank)}; the scratch is self-cleaning (all 0xFF when idle).
//   L1 ("parent voxels with surfels", VoxelMap.h:312-324) — the hash entry IS the storage:
//     l1_tab[slot]  32 B = one sector: {u64 key | has_surfel<<63, float n[3], float c[3]}   <- K2 probe = 1 sector
//     l1_meta[slot] 40 B: child list (<=27 five-bit codes in the reference's child-set order), counts, planarity
// Keys are the reference's own VoxelKeyHash Z-order code (VoxelMap.h:166-183); they are unique as long as
// every voxel coordinate stays inside [-2^20, 2^20), which the engine checks (B2LO_E_RANGE).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "b2lo_math.cuh"

// ---- in-graph timeline (debug builds only: make TIMELINE=1 -> libb2lo_tl.so) ----------------------------------------------------
// A replayed CUDA graph is one opaque item to ncu, and its serialised, cold-cache launch list says nothing about the gaps between
// kernels.  With -DB2LO_TIMELINE the first thread of every kernel of the per-scan path files (file id, source line, %globaltimer) into a
// per-translation-unit buffer; b2lo_debug_timeline() merges them by time, which gives the start-to-start intervals of the scan's
// kernels as they really run inside the graph (tools/gpu_timeline.py).  The macro is empty in the product build.
#ifdef B2LO_TIMELINE
#ifndef B2LO_TL_FILE
#define B2LO_TL_FILE 0
#endif
namespace b2 {
constexpr int TL_CAP = 1 << 16;
static __device__ unsigned long long g_tl[2 * TL_CAP];
static __device__ unsigned int g_tl_n;
__device__ __forceinline__ void tl_mark(int line) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned int i = atomicAdd(&g_tl_n, 1u);
  if (i < TL_CAP) { g_tl[2 * i] = ((unsigned long long)B2LO_TL_FILE << 32) | (unsigned)line; g_tl[2 * i + 1] = t; }
}
// host side of this translation unit's buffer: copies the marks out and clears the counter
static inline int tl_fetch(unsigned long long* out, int cap) {
  unsigned int n = 0;
  cudaMemcpyFromSymbol(&n, g_tl_n, sizeof n);
  if ((int)n > TL_CAP) n = TL_CAP;
  if ((int)n > cap) n = cap;
  if (n) cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 2 * n);
  unsigned int zero = 0;
  cudaMemcpyToSymbol(g_tl_n, &zero, sizeof zero);
  return (int)n;
}
}  // namespace b2
#define TL_START() do { if (threadIdx.x == 0 && blockIdx.x == 0) b2::tl_mark(__LINE__); } while (0)
#define TL_HERE() b2::tl_mark(__LINE__)
#else
#define TL_START() do {} while (0)
#define TL_HERE() do {} while (0)
#endif

namespace b2 {

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t KEY_TOMB = 0xFFFFFFFFFFFFFFFEull;
constexpr uint64_t KEY_MASK = 0x7FFFFFFFFFFFFFFFull;
constexpr uint64_t SURFEL_BIT = 0x8000000000000000ull;
constexpr uint32_t POS_PENDING = 0xFFFFFFFFu;

struct __align__(32) L0Entry { unsigned long long key; uint32_t pos; unsigned int first; int cnt; int head; int rank; int pad; };  // 32 B
struct __align__(32) L1Entry { unsigned long long key; float n[3]; float c[3]; };    // 32 B
struct L1Meta { uint8_t child[27]; uint8_t nchild; float planarity; int last_child_count; int mark; };  // 40 B (mark: per-update dedupe flag)

__host__ __device__ __forceinline__ uint64_t expand21(uint64_t v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
__host__ __device__ __forceinline__ uint32_t compact21(uint64_t v) {
  v &= 0x1249249249249249ull;
  v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
  v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
  v = (v | (v >> 8)) & 0x1F0000FF0000FFull;
  v = (v | (v >> 16)) & 0x1F00000000FFFFull;
  v = (v | (v >> 32)) & 0x1FFFFFull;
  return (uint32_t)v;
}
// VoxelKeyHash: Z-order code of (v + 2^20) & 0x1fffff per axis, x in the lowest bit
__host__ __device__ __forceinline__ uint64_t key_morton(int x, int y, int z) {
  return expand21((uint64_t)(int64_t)(x + (1 << 20))) | (expand21((uint64_t)(int64_t)(y + (1 << 20))) << 1) |
         (expand21((uint64_t)(int64_t)(z + (1 << 20))) << 2);
}
__host__ __device__ __forceinline__ void morton_key(uint64_t m, int& x, int& y, int& z) {
  x = (int)compact21(m) - (1 << 20);
  y = (int)compact21(m >> 1) - (1 << 20);
  z = (int)compact21(m >> 2) - (1 << 20);
}
// Internal key of the two hash tables: the three biased 21-bit coordinates side by side (x lowest).  The reference's Morton
// code (VoxelKeyHash, VoxelMap.h:168-182; b2lo_voxel_key_hash) is only the hash FUNCTION of its map and is re-mixed by
// unordered_dense; any bijection of the 63 bits gives the same container semantics, and packing costs ~8 integer instructions
// per key where three bit interleaves cost ~75 (K2 and K5 build one key per query per iteration).
__host__ __device__ __forceinline__ uint64_t key_pack(int x, int y, int z) {
  return (uint64_t)((uint32_t)(x + (1 << 20)) & 0x1fffffu) | ((uint64_t)((uint32_t)(y + (1 << 20)) & 0x1fffffu) << 21) |
         ((uint64_t)((uint32_t)(z + (1 << 20)) & 0x1fffffu) << 42);
}
__host__ __device__ __forceinline__ void key_unpack(uint64_t k, int& x, int& y, int& z) {
  x = (int)((uint32_t)k & 0x1fffffu) - (1 << 20);
  y = (int)((uint32_t)(k >> 21) & 0x1fffffu) - (1 << 20);
  z = (int)((uint32_t)(k >> 42) & 0x1fffffu) - (1 << 20);
}
__host__ __device__ __forceinline__ bool key_in_range(int x, int y, int z) {
  return (unsigned)(x + (1 << 20)) < (1u << 21) && (unsigned)(y + (1 << 20)) < (1u << 21) && (unsigned)(z + (1 << 20)) < (1u << 21);
}
// floor-division parent (VoxelMap::GetParentKey, VoxelMap.cpp:60-67)
__host__ __device__ __forceinline__ int parent_coord(int k, int f) { return k >= 0 ? k / f : (k - (f - 1)) / f; }

__device__ __forceinline__ uint32_t hash_slot(uint64_t key, int log2cap) {
  return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
}

// float -> voxel coordinate exactly as (int)std::floor(p / scale) (VoxelMap.cpp:54-56); out-of-int-range
// values are flagged by the caller through key_in_range on the saturated result.
__device__ __forceinline__ int voxel_coord(float p, float scale) {
  float q = floorf(p / scale);
  if (!(q > -2147483000.0f)) return INT_MIN + 1;
  if (!(q < 2147483000.0f)) return INT_MAX - 1;
  return (int)q;
}

// EXPERIMENT, not used by K2 (see corr_issue): the same value without the division on the common path: q = p * (1/scale) is within 1.5 * 2^-23 |q| of the correctly rounded quotient,
// so floor(q) can differ from floor(p / scale) only when q sits within a few ulp of an integer - then (and for |q| >= 2^23, where every
// float is an integer) the true division decides.  Bit-identical to voxel_coord for every input (tests: voxel-boundary values through
// the correspondence taps); ~6 instructions instead of the ~20 of an IEEE f32 division, three times per query in K2.
__device__ __forceinline__ int voxel_coord_fast(float p, float scale, float inv) {
  const float qm = p * inv;
  float q = floorf(qm);
  const float d = qm - q, tol = fabsf(qm) * 4.0e-7f + 1.0e-30f;
  if (!(d >= tol && d <= 1.0f - tol)) q = floorf(p / scale);      // also NaN / inf / huge
  if (!(q > -2147483000.0f)) return INT_MIN + 1;
  if (!(q < 2147483000.0f)) return INT_MAX - 1;
  return (int)q;
}
// EXPERIMENT, not used: a locality-preserving slot function for the L1 table - the four cells of a 2 x 2 block in x, y land in ONE 128 B
// line (the block id is hashed, the low x and y bits pick the 32 B entry inside the line), so that the 64 B DRAM access of a probe also
// brings the neighbour's entry.  Same-box A/B on the 10^7-voxel map (2^20 queries): coherent sweep 25.6 -> 23.9 us, random probes
// 41.2 -> 45.9 us (blocks of four cluster under linear probing and lengthen the chains), lone KITTI sequence +3 %.  Kept for a bucketised
// table (probe a whole line, then hop); the L1 table uses hash_slot.
__device__ __forceinline__ uint32_t hash_slot_l1_block(uint64_t key, int log2cap) {
  const uint64_t blk = key & ~(1ull | (1ull << 21));
  const uint32_t h = (uint32_t)((blk * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
  return (h & ~3u) | ((uint32_t)key & 1u) | (((uint32_t)(key >> 21) & 1u) << 1);
}

struct MapDev {
  // parameters
  float voxel, scale1;   // scale1 = voxel * (float)factor, rounded to f32 first as the reference does
  int factor;
  float planarity_thr;
  int compute_surfels;
  // L0
  float4* l0_cent; unsigned long long* l0_key; uint32_t* l0_slot;
  L0Entry* l0_tab; int l0_log2cap; uint32_t l0_cap;   // dense capacity
  // L1
  L1Entry* l1_tab; L1Meta* l1_meta; int l1_log2cap;
  // optional device-side gate of a speculatively enqueued update (odometry: keyframe decided on the device): when non-null and
  // *gate == 0 every update kernel returns at once; sensor_dev (row-major 4x4 pose, translation = sensor position) then replaces the
  // host-passed sensor position
  const int* gate; const float* sensor_dev;
  // counters (device): [0]=n0, [1]=n1, [2]=l0 tombstones, [3]=l1 tombstones, [4]=error flags, [5]=surfel count
  int* ctr;
};

// ---- L0 hash -------------------------------------------------------------------------------------
__device__ __forceinline__ int l0_find(const MapDev& M, uint64_t key) {
  uint32_t mask = (1u << M.l0_log2cap) - 1u;
  uint32_t s = hash_slot(key, M.l0_log2cap);
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long k = M.l0_tab[s].key;
    if (k == key) return (int)s;
    if (k == KEY_EMPTY) return -1;
    s = (s + 1) & mask;
  }
  return -1;
}
// find-or-insert; *inserted tells whether this call created the entry (pos = POS_PENDING).
__device__ __forceinline__ int l0_find_or_insert(const MapDev& M, uint64_t key, bool* inserted) {
  uint32_t mask = (1u << M.l0_log2cap) - 1u;
  uint32_t s = hash_slot(key, M.l0_log2cap);
  *inserted = false;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long k = *((volatile unsigned long long*)&M.l0_tab[s].key);
    if (k == key) return (int)s;
    if (k == KEY_TOMB) {
      // tombstones are never recycled inside an update (another thread may be inserting the same key
      // further down the chain); they are dropped by the periodic rebuild.
    } else if (k == KEY_EMPTY) {
      unsigned long long old = atomicCAS(&M.l0_tab[s].key, KEY_EMPTY, (unsigned long long)key);
      if (old == KEY_EMPTY) { *inserted = true; return (int)s; }  // pos is POS_PENDING (0xFFFFFFFF) in a cleared entry
      if (old == key) return (int)s;
    }
    s = (s + 1) & mask;
  }
  return -1;
}

// ---- L1 hash -------------------------------------------------------------------------------------
__device__ __forceinline__ int l1_find(const MapDev& M, uint64_t key) {
  uint32_t mask = (1u << M.l1_log2cap) - 1u;
  uint32_t s = hash_slot(key, M.l1_log2cap);
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long k = M.l1_tab[s].key;
    if (k != KEY_EMPTY && k != KEY_TOMB && (k & KEY_MASK) == key) return (int)s;
    if (k == KEY_EMPTY) return -1;
    s = (s + 1) & mask;
  }
  return -1;
}
__device__ __forceinline__ int l1_find_or_insert(const MapDev& M, uint64_t key, bool* inserted) {
  uint32_t mask = (1u << M.l1_log2cap) - 1u;
  uint32_t s = hash_slot(key, M.l1_log2cap);
  *inserted = false;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long k = *((volatile unsigned long long*)&M.l1_tab[s].key);
    if (k != KEY_EMPTY && k != KEY_TOMB && (k & KEY_MASK) == key) return (int)s;
    if (k == KEY_EMPTY) {
      unsigned long long old = atomicCAS(&M.l1_tab[s].key, KEY_EMPTY, (unsigned long long)key);
      if (old == KEY_EMPTY) { *inserted = true; return (int)s; }
      if (old != KEY_TOMB && (old & KEY_MASK) == key) return (int)s;
    }
    s = (s + 1) & mask;
  }
  return -1;
}

// ---- block-level helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
  return v;
}
// exclusive scan of one int per thread across the block (blockDim.x <= 1024); returns exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* smem /*>=33 ints*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int x = lane < nw ? smem[lane] : 0;
    int xi = warp_incl_scan(x);
    smem[lane] = xi - x;
    if (lane == 31) smem[32] = xi;
  }
  __syncthreads();
  int res = smem[w] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

}  // namespace b2
