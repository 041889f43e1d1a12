// b2lo_filter.cu — K1: stride + Z-order voxel downsample on the device.
//
// Replaces map::FastVoxelFilter::filter (/root/reference/src/database/VoxelMap.h:73-104) with its
// computeMortonKey (:124-135) / expandBits (:114-122).  Contract reproduced bit-for-bit:
//   * every `stride`-th point, non-finite points skipped;
//   * key = Z-order of clamp(floor(x * (1/voxel)) + 2^20, 0, 2^21-1) per axis;
//   * per voxel the f32 sums run SEQUENTIALLY in input order, centroid = sum * (1/(float)count);
//   * output order = first-seen order of the voxels (unordered_dense iteration order).
// Device algorithm (no global sort, six small kernels on the context stream):
//   F1 insert   : sampled point -> key -> claim/find a slot in a scratch hash; count and min-index per voxel
//   F2 flags    : per point, packed (is-leader << 32 | points of its voxel) - the dependent table look-ups, spread over the grid
//   F3 scan     : one CTA, coalesced; voxel rank = prefix count of leaders (first point of each voxel) in input order,
//                 segment offsets = prefix sum of the per-voxel counts in that order (one packed 64-bit scan)
//   F4 fill     : every point drops its index into its voxel's segment (arbitrary slot)
//   F5 rank     : every point counts the smaller indices of its segment and drops its coordinates at that position
//   F6 reduce   : one thread per voxel streams its (now ordered, contiguous) points, adds them in input order, writes the centroid
// Algorithmic bytes: 16 B read per sampled point + 16 B written per voxel.
#include <climits>
#include <cstdlib>
#define B2LO_TL_FILE 1
#include "b2lo_internal.h"
#include "b2lo_launch.cuh"

namespace b2 {

__device__ __forceinline__ long long x86_f2ll(float f) {
  // static_cast<int64_t>(float) as cvttss2si does it: out-of-range -> INT64_MIN ("integer indefinite")
  if (!(f < 9223372036854775808.0f) || f < -9223372036854775808.0f) return LLONG_MIN;
  return (long long)f;
}
__device__ __forceinline__ unsigned long long filter_key(float x, float y, float z, float inv) {
  const long long OFF = 1ll << 20, HI = (1ll << 21) - 1;
  long long ix = x86_f2ll(floorf(x * inv)) + OFF;
  long long iy = x86_f2ll(floorf(y * inv)) + OFF;
  long long iz = x86_f2ll(floorf(z * inv)) + OFF;
  ix = ix < 0 ? 0 : (ix > HI ? HI : ix);
  iy = iy < 0 ? 0 : (iy > HI ? HI : iy);
  iz = iz < 0 ? 0 : (iz > HI ? HI : iz);
  return expand21((uint64_t)ix) | (expand21((uint64_t)iy) << 1) | (expand21((uint64_t)iz) << 2);
}

// memcpy(&f, p, 4) of the reference's PLY reader for any alignment: one aligned 32-bit load when possible, else four byte loads
__device__ __forceinline__ float load_f32_bytes(const unsigned char* p) {
  if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0) return *reinterpret_cast<const float*>(p);
  const unsigned int u = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
  return __uint_as_float(u);
}

struct k_flt_insert { static __device__ __forceinline__ void run(const ScanParams* __restrict__ sp, FEntry* tab, int log2cap,
                             float4* samp, int* slot_of) { TL_START();
  const float* __restrict__ src = sp->flt_src;
  const int n_samples = sp->flt_ns;
  const size_t sample_stride = (size_t)sp->flt_stride;
  const float inv = sp->flt_inv;
  const unsigned int rec = sp->flt_rec, ox = sp->flt_off[0], oy = sp->flt_off[1], oz = sp->flt_off[2];
  const unsigned int mode = sp->flt_mode;
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count: the lanes of a warp pool their table updates below
  for (int j0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); j0 < n_samples; j0 += gridDim.x * blockDim.x) {
    const int j = j0 + lane;
    const bool valid = j < n_samples;
    float x = 0.0f, y = 0.0f, z = 0.0f;
    if (valid) {
    if (rec) {   // byte records (PLY vertices, ply_player.cpp:330-337): three 4-byte copies at arbitrary, possibly unaligned offsets
      const unsigned char* b = reinterpret_cast<const unsigned char*>(src) + (size_t)j * sample_stride;
      x = load_f32_bytes(b + ox); y = load_f32_bytes(b + oy); z = load_f32_bytes(b + oz);
    } else {
      const float* p = src + (size_t)j * sample_stride;
      x = p[0]; y = p[1]; z = p[2];
    }
    samp[j] = make_float4(x, y, z, 0.0f);
    }
    int s = -1;
    bool take = valid && isfinite(x) && isfinite(y) && isfinite(z);
    unsigned long long key = 0ull;
    if (take) {
      if (mode == 0) key = filter_key(x, y, z, inv);
      else {
        // util::VoxelGrid::get_voxel_key (PointCloudUtils.h:547-552): floor(p / leaf) per axis (true division; `inv` carries the leaf);
        // x is packed highest so that ascending keys are std::map<VoxelKey>'s (x, y, z) order
        const int gx = voxel_coord(x, inv), gy = voxel_coord(y, inv), gz = voxel_coord(z, inv);
        take = key_in_range(gx, gy, gz);
        key = key_pack(gz, gy, gx);
      }
    }
    // consecutive returns of a ring fall into the same voxel: the lanes of a warp that hold the same key elect their lowest lane
    // (= lowest point index) to find / insert the cell and to file the whole group's count - a quarter of the atomics on hot cells
    const unsigned grp = __match_any_sync(0xffffffffu, take ? key : 0xFFFFFFFFFFFFFFFFull);
    const int leader = __ffs(grp) - 1;
    if (take && lane == leader) {
      uint32_t mask = (1u << log2cap) - 1u;
      uint32_t h = hash_slot(key, log2cap);
      for (;;) {
        unsigned long long k = *((volatile unsigned long long*)&tab[h].key);
        if (k == key) break;
        if (k == KEY_EMPTY) {
          unsigned long long old = atomicCAS(&tab[h].key, KEY_EMPTY, key);
          if (old == KEY_EMPTY || old == key) break;
        }
        h = (h + 1) & mask;
      }
      s = (int)h;
      atomicAdd(&tab[h].cnt, __popc(grp));       // starts at -1 (memset 0xFF): holds count-1
      atomicMin(&tab[h].first, (unsigned)j);     // starts at 0xFFFFFFFF
    }
    s = __shfl_sync(0xffffffffu, s, leader);
    if (valid) slot_of[j] = take ? s : -1;
  }
} };

// exclusive scan of one u64 per thread across the block; *total = block sum (smem: >= 33 u64)
__device__ __forceinline__ unsigned long long block_excl_scan64(unsigned long long v, unsigned long long* total, unsigned long long* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    unsigned long long x = lane < nw ? smem[lane] : 0ull, xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long n = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += n; }
    smem[lane] = xi - x;
    if (lane == 31) smem[32] = xi;
  }
  __syncthreads();
  unsigned long long res = smem[w] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

struct k_flt_flags { static __device__ __forceinline__ void run(const FEntry* __restrict__ tab, const int* __restrict__ slot_of, const ScanParams* __restrict__ sp,
                            unsigned long long* __restrict__ packed) { TL_START();
  const int n_samples = sp->flt_ns;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_samples; j += gridDim.x * blockDim.x) {
    const int s = slot_of[j];
    unsigned long long p = 0ull;
    if (s >= 0) { const FEntry e = tab[s]; if (e.first == (unsigned)j) p = (1ull << 32) | (unsigned long long)(unsigned)(e.cnt + 1); }
    packed[j] = p;
  }
} };
// one CTA of 1024 threads walks the packed flags in input order, IPT per thread; the voxel rank (count of earlier leaders) and
// the segment offset (sum of earlier leaders' point counts) ride one packed 64-bit scan: leaders << 32 | points
template <int IPT>
struct k_flt_scan { static __device__ __forceinline__ void run(const unsigned long long* __restrict__ packed, const ScanParams* __restrict__ sp,
                                                    int* vid_of_point, int* seg_start, int* seg_cnt, int* lead_of_vid, int* d_nvox) { TL_START();
  __shared__ unsigned long long sm[40];
  const int n_samples = sp->flt_ns;
  unsigned long long base = 0ull;
  for (int t0 = 0; t0 < n_samples; t0 += IPT * blockDim.x) {
    const int j0 = t0 + IPT * threadIdx.x;
    unsigned long long p[IPT], sum = 0ull;
    if (j0 + IPT <= n_samples) {   // whole 16 B vector loads
      const ulonglong2* v = reinterpret_cast<const ulonglong2*>(packed + j0);
#pragma unroll
      for (int k = 0; k < IPT / 2; ++k) { const ulonglong2 q = v[k]; p[2 * k] = q.x; p[2 * k + 1] = q.y; }
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) p[k] = (j0 + k < n_samples) ? packed[j0 + k] : 0ull;
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) sum += p[k];
    unsigned long long tot;
    unsigned long long ex = base + block_excl_scan64(sum, &tot, sm);
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      if (p[k]) {
        int v = (int)(ex >> 32), off = (int)(ex & 0xffffffffull), cnt = (int)(p[k] & 0xffffffffull);
        vid_of_point[j0 + k] = v;
        seg_start[v] = off;
        seg_cnt[v] = cnt;
        lead_of_vid[v] = j0 + k;
      }
      ex += p[k];
    }
    base += tot;
  }
  if (threadIdx.x == 0) { *d_nvox = (int)(base >> 32); seg_start[(int)(base >> 32)] = (int)(base & 0xffffffffull); }
} };

// vid_pt[j] = voxel id of sample j (-1: dropped), left for k_flt_rank so that it does not repeat the slot -> leader -> voxel chain
struct k_flt_fill { static __device__ __forceinline__ void run(FEntry* tab, const int* __restrict__ slot_of, const ScanParams* __restrict__ sp, const int* __restrict__ vid_of_point,
                           const int* __restrict__ seg_start, int* bucket, int* vid_pt) { TL_START();
  const int n_samples = sp->flt_ns;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_samples; j += gridDim.x * blockDim.x) {
    int s = slot_of[j];
    if (s < 0) { vid_pt[j] = -1; continue; }
    int t = atomicSub(&tab[s].cnt, 1);  // count-1, count-2, ..., 0 : a unique ticket in [0, count)  (independent of the chain below)
    int v = vid_of_point[tab[s].first];
    vid_pt[j] = v;
    bucket[seg_start[v] + t] = j;
  }
} };

// every point counts the smaller indices of its voxel's segment -> its position in input order; it drops its COORDINATES
// there, so that the reduction below streams contiguous float4s (the O(m^2) ordering work of a crowded voxel is spread
// over its m points' threads instead of serialising on one)
struct k_flt_rank { static __device__ __forceinline__ void run(const int* __restrict__ vid_pt, const ScanParams* __restrict__ sp, const int* __restrict__ seg_start,
                           const int* __restrict__ seg_cnt, const int* __restrict__ bucket, const float4* __restrict__ samp,
                           float4* __restrict__ sorted) { TL_START();
  const int n_samples = sp->flt_ns;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_samples; j += gridDim.x * blockDim.x) {
    const int v = vid_pt[j];
    const float4 p = samp[j];   // independent of the ranking chain
    if (v < 0) continue;
    const int b = seg_start[v], m = seg_cnt[v];
    int r = 0, q = 0;
    for (; q + 8 <= m; q += 8) {
      int t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = bucket[b + q + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) r += (t[u] < j);
    }
    for (; q < m; ++q) r += (bucket[b + q] < j);
    sorted[b + r] = p;
  }
} };

// one thread per voxel adds its points in input order (sequential f32, VoxelMap.h:88-91) and writes centroid and key.  A crowded voxel
// (a wall or the ground next to the sensor: 100-350 sampled points) would make its one thread the critical path of the whole kernel
// (a dependent chain of L2 round trips, 16 points each); such voxels are handed to their whole warp instead: 32 points per coalesced
// load, then the owner lane takes them through shuffles in input order - the same additions in the same order.  In-graph timeline,
// same box, 60 KITTI-shaped scans: 19.5 us per scan with a thread per voxel and neighbours in one warp, 13.5 with the spread mapping
// below, 10.2 with the warp hand-over on top.
constexpr int FLT_HEAVY = 48;
struct k_flt_reduce { static __device__ __forceinline__ void run(const int* __restrict__ d_nvox, const int* __restrict__ seg_start, const int* __restrict__ seg_cnt,
                             const float4* __restrict__ sorted, const int* __restrict__ lead_of_vid, const int* __restrict__ slot_of,
                             FEntry* tab, float4* out, unsigned long long* out_key, const ScanParams* __restrict__ sp, int heavy_min) { TL_START();
  // the scratch hash is self-cleaning: every entry in use belongs to exactly one voxel, whose thread puts it back to idle (all 0xFF)
  // after taking the key - no 0.5 MB memset node per scan
  const int nv = *d_nvox;
  const unsigned int mode = sp->flt_mode;
  const int lane = threadIdx.x & 31;
  // lane l of warp w takes voxel w + W l (W = warps in the grid): crowded voxels are neighbours in first-seen order, and this way they
  // land in DIFFERENT warps, each of which hands its crowded voxel to all 32 lanes (a warp walks its own crowded voxels one by one)
  const int W = (int)((gridDim.x * blockDim.x) >> 5), wg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  for (int v0 = 0; v0 < nv; v0 += 32 * W) {   // warp-uniform trip count
    const int v = v0 + wg + W * lane;
    const bool live = v < nv;
    int b = 0, m = 0;
    if (live) { b = seg_start[v]; m = seg_cnt[v]; }
    if (live && mode == 1) {
      // util::VoxelGrid::WeightedCentroid::add_point (PointCloudUtils.h:502-520): first point copied, then
      // c = (w / (w + 1)) * c + (1 / (w + 1)) * p in f32, points in input order
      float4 c = sorted[b];
      float wgt = 1.0f;
      for (int q = 1; q < m; ++q) {
        const float4 p = sorted[b + q];
        const float tot = wgt + 1.0f, ro = wgt / tot, rn = 1.0f / tot;
        c.x = ro * c.x + rn * p.x; c.y = ro * c.y + rn * p.y; c.z = ro * c.z + rn * p.z;
        wgt = tot;
      }
      out[v] = make_float4(c.x, c.y, c.z, wgt);
      FEntry* e = &tab[slot_of[lead_of_vid[v]]];
      out_key[v] = e->key;
      e->key = KEY_EMPTY; e->cnt = -1; e->first = 0xFFFFFFFFu;
    }
    if (mode == 1) continue;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    const bool heavy = live && m > heavy_min;
    if (live && !heavy) {
      int q = 0;
      for (; q + 8 <= m; q += 8) {  // eight independent loads in flight, adds stay in order
        float4 p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u] = sorted[b + q + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) { sx += p[u].x; sy += p[u].y; sz += p[u].z; }
      }
      for (; q < m; ++q) { float4 p = sorted[b + q]; sx += p.x; sy += p.y; sz += p.z; }
    }
    unsigned hm = __ballot_sync(0xffffffffu, heavy);
    while (hm) {
      const int h = __ffs(hm) - 1;
      hm &= hm - 1;
      const int hb = __shfl_sync(0xffffffffu, b, h), hn = __shfl_sync(0xffffffffu, m, h);
      float4 nx = (lane < hn) ? sorted[hb + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < hn; q += 32) {
        const float4 p = nx;
        if (q + 32 < hn) nx = (q + 32 + lane < hn) ? sorted[hb + q + 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);   // next chunk under way
        const int cnt = hn - q < 32 ? hn - q : 32;
        for (int l = 0; l < cnt; ++l) {
          const float vx = __shfl_sync(0xffffffffu, p.x, l), vy = __shfl_sync(0xffffffffu, p.y, l), vz = __shfl_sync(0xffffffffu, p.z, l);
          if (lane == h) { sx += vx; sy += vy; sz += vz; }
        }
      }
    }
    if (live) {
      const float ic = 1.0f / (float)(unsigned)m;
      out[v] = make_float4(sx * ic, sy * ic, sz * ic, 0.0f);
      FEntry* e = &tab[slot_of[lead_of_vid[v]]];
      out_key[v] = e->key;
      e->key = KEY_EMPTY; e->cnt = -1; e->first = 0xFFFFFFFFu;
    }
  }
} };

int filter_run(b2lo_ctx* ctx, const float* src_dev, size_t n_samples, size_t sample_stride_floats, float voxel, int set, cudaStream_t on,
               const b2lo_record_fmt* fmt, int mode) {
  cudaStream_t st = on ? on : ctx->stream;
  if (set == 0) ctx->feat_set = 0;
  ctx->feat_cap_hint_set[set] = n_samples;
  if (n_samples == 0) { if (set == 0) ctx->feat_cap_hint = 0; B2_CUDA(cudaMemsetAsync(ctx->nfeat(set), 0, sizeof(int), st)); return B2LO_OK; }
  if (n_samples > (size_t)INT_MAX / 4) { set_error("filter: too many samples"); return B2LO_E_CAPACITY; }
  int rc = ctx_reserve_points(ctx, n_samples);
  if (rc) return rc;
  // launch geometry and scratch size depend on the CAPACITY only, the actual count travels in the parameter block:
  // the launch sequence is identical from scan to scan (graph-replayable)
  int log2cap = 12;
  while ((1ull << log2cap) < 2 * n_samples) ++log2cap;   // constant while the sample count stays in one power-of-two bucket
  if (log2cap > ctx->f_log2cap) { set_error("filter: scratch hash too small"); return B2LO_E_CAPACITY; }
  ctx->f_log2_last = log2cap;
  if (set == 0) ctx->feat_cap_hint = n_samples;
  if (!ctx->sp_preloaded) {
    if ((rc = sp_begin_write(ctx))) return rc;
    ctx->h_sp->flt_src = src_dev; ctx->h_sp->flt_ns = (int)n_samples; ctx->h_sp->flt_stride = sample_stride_floats;
    ctx->h_sp->flt_inv = mode == 1 ? voxel : 1.0f / voxel;  // m_inv_voxel_size (VoxelMap.h:57); VoxelGrid divides by the leaf size itself
    ctx->h_sp->flt_mode = (unsigned)mode;
    ctx->h_sp->flt_rec = fmt ? fmt->record_bytes : 0u;
    for (int a = 0; a < 3; ++a) ctx->h_sp->flt_off[a] = fmt ? (&fmt->off_x)[a] : 0u;
    if ((rc = sp_upload(ctx, 0, offsetof(ScanParams, T_init)))) return rc;
  }
  // ctx->f_tab is all-idle (0xFF) here: memset when allocated, put back by k_flt_reduce after every run
  int blocks = (int)((ctx->pts_cap + 255) / 256); if (blocks > 1184) blocks = 1184;
  blocks = batch_grid(ctx, blocks);
  const ScanParams* sp = ctx->d_sp;
  const bool prof = st == ctx->stream;   // the per-stage events live on the context stream
  if (prof) prof_begin(ctx, PS_FILTER);
  launch<k_flt_insert, 256, 1>(ctx, dim3((unsigned)(blocks)), dim3((unsigned)(256)), 0, st, sp, ctx->f_tab, log2cap, ctx->f_samp, ctx->f_slot);
  // the launch geometry follows the capacity, not the count (graph-replayable): scan-sized buffers take the one-trip scan
  launch<k_flt_flags, 256, 1>(ctx, dim3((unsigned)(blocks)), dim3((unsigned)(256)), 0, st, ctx->f_tab, ctx->f_slot, sp, ctx->f_packed);
  // 4 flags per thread and trip: a one-trip variant with 16 per thread measured slower (13.3 vs 9.9 us per scan, same box)
  launch<k_flt_scan<4>, 1024, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(1024)), 0, st, ctx->f_packed, sp, ctx->f_vid, ctx->f_segstart, ctx->f_segcnt, ctx->f_lead, ctx->nfeat(set));
  // (the packed flags are dead after the scan: their buffer carries the per-sample voxel ids from here on)
  launch<k_flt_fill, 256, 1>(ctx, dim3((unsigned)(blocks)), dim3((unsigned)(256)), 0, st, ctx->f_tab, ctx->f_slot, sp, ctx->f_vid, ctx->f_segstart, ctx->f_bucket, reinterpret_cast<int*>(ctx->f_packed));
  launch<k_flt_rank, 256, 1>(ctx, dim3((unsigned)(blocks)), dim3((unsigned)(256)), 0, st, reinterpret_cast<const int*>(ctx->f_packed), sp, ctx->f_segstart, ctx->f_segcnt, ctx->f_bucket, ctx->f_samp, ctx->f_sorted);
  launch<k_flt_reduce, 256, 1, 4>(ctx, dim3((unsigned)(blocks)), dim3((unsigned)(256)), 0, st, ctx->nfeat(set), ctx->f_segstart, ctx->f_segcnt, ctx->f_sorted, ctx->f_lead, ctx->f_slot, ctx->f_tab,
                                       ctx->feat(set), ctx->feat_key(set), sp, FLT_HEAVY);
  if (prof) prof_end(ctx);
  ctx->launches += 6;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}

}  // namespace b2

#ifdef B2LO_TIMELINE
namespace b2 { int tl_fetch_filter(unsigned long long* out, int cap) { return tl_fetch(out, cap); } }
#endif
