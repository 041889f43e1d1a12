// b2lo_internal.h — host-side objects behind the opaque handles of include/b2lo.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/b2lo.h"
#include "b2lo_dev.cuh"

namespace b2 { struct Recorder; }
namespace b2 {

void set_error(const char* fmt, ...);
#define B2_CUDA(expr)                                                                               \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      b2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);    \
      return B2LO_E_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

// ---- device-side state blocks ------------------------------------------------------------------------
struct FEntry { unsigned long long key; int cnt; unsigned int first; };  // filter scratch hash, 16 B

struct PkoTables {                // built on the host once per context (b2lo_pko_host.cpp)
  double alpha[129], Z[129];      // alpha candidates and partition functions (AdaptiveMEstimator.cpp:218-241)
  double exp2_t1[32], exp2_t2[32];     // 2^(j/32), 2^(j/1024): tables of the EM's exp (k_icp_pko1)
  int n_alpha;                    // num_alpha_segments + 1
  int kmeans_seed[129][2];        // uniform_int_distribution(0, ns-1)(mt19937(42)) draws for sample size ns
  int head_r[3][128];             // first swaps r_i (i < 128) of std::shuffle per mode (even, odd, large n)
  int head_pos[3][128];           // per mode: where position j ends up after the first 127 swaps alone (valid for n >= 128)
  int hit_off[3][129];            // per mode, per head position j: [hit_off[j], hit_off[j+1]) into hits
  int hit_n;
  // config scalars
  double min_sf, max_sf, trunc;
  int sample_size, kernel_type;
};

struct IcpState {                 // lives in device memory, one per context
  float R[9], t[3];               // current pose
  float T_init[16];
  int iter, done, status;         // done: 1 converged / max reached, 2 failed
  int n_corr, n_blocks;
  double scale, delta;
  int em_iters, kmeans_iters;
  int scale_forced;               // the scale came from ScanParams::force_scale: iteration 0 must not recompute it
  unsigned int ticket;            // last-block election of the GN reduction
  unsigned int ticket_corr;       // last-block election of the correspondence kernel (fused PKO fit)
  int num_iterations; int converged; double initial_cost, final_cost;
  b2lo_iter_trace trace[B2LO_MAX_ITERS];
  long long dbg[32];                // clock64 stamps of the single-CTA phases (tools/gpu_phase_clocks.py)
};

// Per-call scalars that change from scan to scan live in ONE device block, refreshed by a single H2D copy from its pinned
// mirror: the kernels read them through a pointer, so the launch sequence of a scan has constant arguments and can be
// replayed as a CUDA graph (b2lo_odom.cu).
struct DecideArgs { float guess[16]; float last_kf[16]; int ran_icp; int n_keyframes; double kf_dist, kf_rot; };
struct ScanParams {
  const float* flt_src; unsigned long long flt_stride; int flt_ns; float flt_inv;   // K1
  // byte-record input (b2lo_record_fmt: PLY vertex records, any layout): flt_rec != 0 -> flt_src is a byte stream, flt_stride the
  // distance between SAMPLED records in bytes, flt_off the byte offsets of the three IEEE f32 coordinates inside a record
  unsigned int flt_rec; unsigned int flt_off[3];
  unsigned int flt_mode;   // 0: FastVoxelFilter (Z-order key of floor(x * inv), centroid = sum / count); 1: util::VoxelGrid (b2lo_export.cu)
  float T_init[16];                                                                   // ICP initial pose
  double force_scale;      // parity tap (b2lo_icp_iterate): > 0 -> residual normalisation scale given by the caller instead of iteration 0's
  DecideArgs decide;                                                                  // odometry tail
};

struct IcpParams {                // kernel-argument POD
  int max_iterations, min_corr, use_robust, loss_type, use_pko, use_surfel, ctile;
  int gn_vgrid = 0;               // K5 in a lock-step batch: the grid a LONE sequence would start (the partition of its partial sums); 0 = gridDim.x
  double tol_t, tol_r, max_dist, robust_delta;
};

}  // namespace b2

namespace b2 {
// optional per-kernel CUDA-event timing on the context stream (bench.py roofline leg); off by default
enum ProfSlot { PS_FILTER = 0, PS_CORR = 1, PS_PKO1 = 2, PS_PKO2 = 3, PS_GN = 4, PS_MAP = 5, PS_XFORM = 6, PS_KNN = 7, PS_CULL = 8, PS_COUNT = 9 };
struct Prof {
  bool on = false;
  static constexpr int POOL = 512;
  cudaEvent_t a[POOL], b[POOL]; int slot[POOL]; int used = 0; bool created = false;
  double ms[PS_COUNT] = {0}; long long n[PS_COUNT] = {0};
};
}  // namespace b2


struct b2lo_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  unsigned long long h2d_bytes = 0, d2h_bytes = 0;   // bytes moved over PCIe by this context (bench accounting)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_stage = nullptr;
  bool stage_busy = false; bool sim_attr_set = false;
  unsigned long long alloc_epoch = 1;   // bumped whenever a device buffer of the context is reallocated (invalidates captured graphs)
  int sm_count = 148;
  size_t shard_m = 0;              // query count of the current point-sharded optimize (b2lo_icp_shard_begin)
  size_t feat_cap_hint = 0;        // host-known upper bound of *d_nfeat (samples of the last filter run into set 0)
  size_t feat_cap_hint_set[2] = {0, 0};   // the same bound per feature set (b2lo_icp_optimize_features follows feat_set)
  cudaGraphExec_t icp_graph_exec = nullptr; unsigned long long icp_graph_sig = 0;
  b2::MapDev* d_mapdev = nullptr;
  // capacities (points)
  size_t pts_cap = 0;
  // pinned staging + device staging for host clouds
  float* h_stage = nullptr; size_t h_stage_floats = 0;
  float* d_stage = nullptr; size_t d_stage_floats = 0;
  float* d_raw = nullptr; size_t raw_floats = 0;     // raw scan DMA target when the caller's buffer is page-locked
  // feature cloud (output of the filter, input of ICP / transform)
  float4* d_feat = nullptr; unsigned long long* d_feat_key = nullptr; int* d_nfeat = nullptr;
  // second feature set: the odometry's look-ahead K1 fills it on stream2 while the current scan registers out of the other one
  float4* d_feat2 = nullptr; unsigned long long* d_feat_key2 = nullptr; int* d_nfeat2 = nullptr;
  cudaStream_t stream2 = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int feat_set = 0;                // which set holds the features of the scan processed last (b2lo_ctx_features reads it)
  float4* feat(int set) const { return set ? d_feat2 : d_feat; }
  unsigned long long* feat_key(int set) const { return set ? d_feat_key2 : d_feat_key; }
  int* nfeat(int set) const { return set ? d_nfeat2 : d_nfeat; }
  // ICP query cloud uploaded from the host
  float4* d_query = nullptr; int* d_nquery = nullptr;
  float4* d_world = nullptr;       // transformed cloud (map update input)
  // filter scratch
  b2::FEntry* f_tab = nullptr; int f_log2cap = 0; int f_log2_last = 0;
  float4* f_samp = nullptr; int* f_slot = nullptr; int* f_vid = nullptr; int* f_segstart = nullptr; int* f_segcnt = nullptr;
  int* f_lead = nullptr; int* f_bucket = nullptr; int* f_ordered = nullptr; unsigned long long* f_packed = nullptr; float4* f_sorted = nullptr;
  // ICP scratch
  double* i_res = nullptr; int* i_slot = nullptr; int* i_cidx = nullptr; int* i_blkcnt = nullptr; int* i_blkoff = nullptr;
  double* i_tilesum = nullptr;     // per compaction tile: sum r, sum r^2 of the accepted queries
  double* i_partial = nullptr;     // per-block 28 doubles
  int i_max_blocks = 0;
  b2::IcpState* d_icp = nullptr; b2::IcpState* h_icp = nullptr /*pinned*/;
  b2::PkoTables* d_pko = nullptr; int* d_pko_hits = nullptr; b2::PkoTables h_pko; std::vector<int> h_pko_hits;
  b2lo_icp_cfg pko_cfg_built{}; bool pko_built = false;
  // KDTree-mode scratch (b2lo_knn.cuh): 5 neighbour ids + found count per query, unresolved queue, fitted planes
  int* k_idx = nullptr; int* k_n = nullptr; int* k_unres = nullptr; int* k_nunres = nullptr; float4* k_plane = nullptr; size_t k_cap = 0;
  int* l_cnt = nullptr;            // loop-closure ICP: [0] size of the matched keyframe's cloud, [1] inlier count
  // parity taps scratch
  int* d_tap_state = nullptr; int* d_tap_key = nullptr; unsigned long long* d_tap_morton = nullptr; float* d_tap_n = nullptr; float* d_tap_c = nullptr;
  int* h_counts = nullptr;         // pinned small readback area (64 ints)
  b2::ScanParams* d_sp = nullptr; b2::ScanParams* h_sp = nullptr /*pinned*/; cudaEvent_t ev_sp = nullptr; bool sp_busy = false;
  bool sp_preloaded = false;       // the caller has already uploaded the whole parameter block for this launch sequence
  b2::Prof* prof = nullptr;
  int batch_S = 0;                 // > 0 while a lock-step batch of that many sequences records: grids of order-independent kernels shrink (batch_grid)
  b2::Recorder* rec = nullptr;     // non-null while a lock-step batch records this context's launch sequence (b2lo_launch.cuh)
  double force_scale = 0.0;        // next icp_run: ScanParams::force_scale (b2lo_icp_iterate), reset by icp_run
  double host_us[8] = {0};         // wall-clock split of the host side of b2lo_odom_process (debug aid): gather, enqueue, wait, ...
  std::recursive_mutex mu;   // taken AFTER a map's mutex by every entry point that touches the context's staging buffers, counters or stream
};

struct b2lo_map {
  b2lo_ctx* ctx = nullptr;
  b2::MapDev d{};                  // device pointers + params (passed by value to kernels)
  size_t tcap0 = 0, tcap1 = 0;
  b2::L1Entry* l1_spare_tab = nullptr; b2::L1Meta* l1_spare_meta = nullptr; size_t l1_spare_cap = 0;   // the L1 table of the previous same-size rebuild
  // update scratch sized by the number of new points
  size_t upd_cap = 0;
  float4* u_pts = nullptr; int* u_pslot = nullptr; int* u_next = nullptr; int* u_isnew = nullptr; int* u_newrank = nullptr;
  float* s_rec = nullptr;           // per affected L1: gathered child centroids (lock-step surfel refit)
  int* u_part = nullptr;            // per-chunk totals of the new-voxel scan for bulk inserts (k_ins_scan_part/top/apply)
  float4* r_tmp = nullptr; size_t r_tmp_cap = 0; int* r_n = nullptr;   // transformed centroids of ApplyTransformAndRehash (kept between calls)
  b2::FEntry* a_tab = nullptr; int a_log2cap = 0;   // affected-L1 set of the current update
  int* a_list = nullptr;                            // purge lists (4 x upd_cap)
  int* a_slots = nullptr;                           // compacted slots of the affected-L1 set
  // cull scratch (sized by dense capacity)
  uint8_t* c_flag = nullptr; int* c_blkcnt = nullptr; int* c_blkoff = nullptr; int* c_removed = nullptr; int* c_aux = nullptr;
  int* c_l1work = nullptr;
  // purge scratch
  int* p_seq = nullptr; int* p_aux = nullptr; size_t p_cap = 0;
  int* u_state = nullptr;          // small device int block of per-update scalars
  // host mirrors
  size_t n0 = 0, n1 = 0, tomb0 = 0, tomb1 = 0;
  bool knn_ready = false;
  bool graph_mode = false;              // update kernels read the live voxel count from the device, grids come from capacities
  unsigned long long alloc_epoch = 1;   // bumped whenever a device buffer of the map is reallocated
  std::recursive_mutex mu;
};

namespace b2 {
// Launch geometry inside a lock-step batch.  A lone sequence sizes its grids by capacity (up to 1184 CTAs for the warp-per-point kernels),
// which costs nothing when the other SMs idle; S sequences side by side would start S x 3900 mostly empty CTAs per step (measured: +19 us
// per sequence and step).  Kernels whose results do not depend on the grid (grid-stride loops over hash / per-voxel work) therefore get
// about two resident waves divided by S.  K2 and K5 keep their grids: their per-CTA partial sums are part of the bit-exact result.
inline int batch_grid(const b2lo_ctx* ctx, int g) {
  if (ctx->batch_S <= 0) return g;
  int lim = 2368 / ctx->batch_S;
  if (lim < 4) lim = 4;
  // test hook: B2LO_BATCH_GRID_LIMIT=<n> forces small per-sequence grids, so that a batch of a few sequences exercises what a batch of
  // hundreds does (CTAs that walk several tiles / virtual blocks); results must not depend on it (tests/test_gpu_more.py)
  const char* e = std::getenv("B2LO_BATCH_GRID_LIMIT");   // read while a batch records its launches (once per graph build), not per scan
  const int forced = e ? std::atoi(e) : 0;
  if (forced > 0) lim = forced;
  return g < lim ? g : lim;
}
void prof_begin(b2lo_ctx* ctx, int slot);
void prof_end(b2lo_ctx* ctx);
void prof_drain(b2lo_ctx* ctx);
int sp_begin_write(b2lo_ctx* ctx);                                   // host may now write ctx->h_sp
int sp_upload(b2lo_ctx* ctx, size_t offset, size_t bytes);           // enqueue the H2D copy of [offset, offset + bytes) of the block
// implemented across the .cu files
int ctx_reserve_points(b2lo_ctx* ctx, size_t n);
int ctx_stage_h2d(b2lo_ctx* ctx, const float* xyz, size_t n, size_t stride_floats, size_t take_every, float4* dst, int* d_count);
int ctx_transform(b2lo_ctx* ctx, const float4* src, const int* d_n, size_t n_cap, const float T16[16], float4* dst);
int ctx_transform_dev(b2lo_ctx* ctx, const float4* src, const int* d_n, size_t n_cap, const float* T16_dev, const int* gate, float4* dst);
int ctx_read_cloud(b2lo_ctx* ctx, const float4* src, const int* d_n, size_t n_cap, float* out_xyz, size_t out_cap, size_t* n_out);
// sample_stride: floats between sampled points, or BYTES between sampled records when fmt != nullptr (byte-record input)
int filter_run(b2lo_ctx* ctx, const float* src_dev, size_t n_samples, size_t sample_stride, float voxel, int set = 0, cudaStream_t on = nullptr,
               const b2lo_record_fmt* fmt = nullptr, int mode = 0);
int ctx_stage_records_h2d(b2lo_ctx* ctx, const void* bytes, size_t n_records, const b2lo_record_fmt* fmt, size_t take_every);
int icp_build_pko(b2lo_ctx* ctx, const b2lo_icp_cfg* cfg);
int icp_prepare(b2lo_ctx* ctx, const b2lo_icp_cfg* cfg);
int icp_run(b2lo_map* map, const float4* d_pts, const int* d_npts, size_t npts_cap, const float* T_init16, const b2lo_icp_cfg* cfg, bool init_pose_on_device,
            bool restore_on_failure = true);
// local / T16_dev (optional): d_world is an OUTPUT, filled inside the update with `local` moved by the device-resident pose T16_dev
int map_update_dev(b2lo_map* map, float4* d_world, const int* d_n, size_t n_cap, const float sensor_f[3], float radius_sq, int rehash = 0,
                   const int* gate = nullptr, const float* sensor_dev = nullptr, const float4* local = nullptr, const float* T16_dev = nullptr);
int map_absorb_counts(b2lo_map* map);
int map_reserve(b2lo_map* map, size_t need_l0, size_t need_upd);
int map_refresh_counts(b2lo_map* map);
int map_rebuild_knn_locked(b2lo_map* map);
void pko_build_host(const b2lo_icp_cfg* cfg, PkoTables* t, std::vector<int>* hits);
}  // namespace b2
