// b2lo_odom.cu — per-scan driver that keeps the scan on the device between the hot-path stages.
//
// Mirrors the in-scope part of processing::Estimator (/root/reference/src/processing/Estimator.cpp):
//   process_frame :116-233, initialize_first_frame :235-269, estimate_motion_dual_frame :271-320,
//   should_create_keyframe :349-368, create_keyframe (map part) :449-470, preprocess_frame :561-589.
// Per scan: K1 downsample -> K2..K5 ICP (device-resident Gauss-Newton loop) -> one small read-back
// (pose + counters) -> host keyframe decision -> K6 map update on keyframes.  The feature cloud never
// leaves HBM; only the 16-float pose and a few counters cross PCIe (plus the strided scan upload for
// b2lo_odom_process).  Loop closure / PGO / viewer stay with the host Estimator and are out of scope.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>
#define B2LO_TL_FILE 3
#include "b2lo_internal.h"
#include "b2lo_launch.cuh"

static inline double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

using namespace b2;

struct OdomDev;
struct b2lo_odom {
  b2lo_ctx* ctx = nullptr;
  b2lo_map* map = nullptr;
  b2lo_odom_cfg cfg{};
  Pose pose, prev_pose, velocity, last_kf_pose;
  bool initialized = false;
  int n_keyframes = 0;
  // replayable launch sequences of one steady-state scan, one per (where K1 runs, which feature set registers): index 2 * mode + set
  cudaGraphExec_t gexec[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned long long gsig[6][4] = {};
  // look-ahead (b2lo_odom_lookahead): the scan announced for the NEXT call, and the scan whose K1 already ran into set pre_set
  bool la_valid = false, pre_valid = false;
  const float* la_src = nullptr; size_t la_ns = 0, la_stride = 0;
  const float* pre_src = nullptr; size_t pre_ns = 0, pre_stride = 0; int pre_set = 0;
  long long lookahead_hits = 0;
  bool allow_graph = true;
  // a steady-state scan between its launch (steady_begin) and the read of its results (steady_finish)
  struct Pending { bool active = false; int mode = 0, set = 0; const float* nx_src = nullptr; size_t nx_ns = 0, nx_stride = 0; double t1 = 0.0;
                   bool synced = false;   // the caller has already waited for the work (a lock-step batch waits once for all its sequences)
  } pend;
  bool has_fmt = false; b2lo_record_fmt fmt{};   // b2lo_odom_set_record_fmt: scans arrive as byte-record streams (KITTI .bin / PLY vertices)
  long long graph_launches = 0, graph_builds = 0, launches_per_graph = 0;
  OdomDev* d_out = nullptr;        // device result block of k_odom_decide
  OdomDev* h_out = nullptr;        // pinned mirror
};

static Pose pose_identity() {
  Pose p;
  p.R = mat3_identity();
  p.t[0] = p.t[1] = p.t[2] = 0.0f;
  return p;
}
static Pose pose_reproject(const Pose& a) {  // SE3f(R.matrix, t): the SO3(Matrix3f) ctor re-projects (MathUtils.h:116-117)
  Pose r = a;
  r.R = so3_project(a.R);
  return r;
}

// Device-side tail of estimate_motion + should_create_keyframe (Estimator.cpp:300-302, 349-368), so that the keyframe
// update can be enqueued behind the ICP without a host round trip: one thread turns the ICP state into the scan's pose
// and decides whether the (already enqueued, gated) map update runs.
struct OdomDev { float pose[16]; int keyframe; int icp_status; int pad[2]; };
struct k_odom_decide { static __device__ __forceinline__ void run(const IcpState* st, const ScanParams* __restrict__ sp, const int* __restrict__ d_nfeat, OdomDev* out) { TL_START();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const DecideArgs a = sp->decide;
  Pose result = pose_from_T16(a.guess);
  int status = B2LO_S_EMPTY;
  if (a.ran_icp) {
    status = st->status;
    if (st->status == B2LO_OK) {
      Pose opt;
      for (int i = 0; i < 9; ++i) opt.R.m[i] = st->R[i];
      for (int i = 0; i < 3; ++i) opt.t[i] = st->t[i];
      opt.R = so3_project_near(opt.R);   // SE3f(optimized.RotationMatrix(), ...) re-projects (Estimator.cpp:300-302)
      result = opt;
    }
  }
  int kf = 1;
  if (a.n_keyframes > 0) {
    Pose last = pose_from_T16(a.last_kf);
    float d[3] = {result.t[0] - last.t[0], result.t[1] - last.t[1], result.t[2] - last.t[2]};
    double distance = (double)sqrtf(sqn3(d));
    Mat3 rd = so3_project_near(mat3_mul(mat3_t(last.R), result.R));
    float lg[3];
    so3_log(rd, lg);
    double angle = (double)sqrtf(sqn3(lg));
    kf = (distance > a.kf_dist || angle > a.kf_rot) ? 1 : 0;
  }
  if (*d_nfeat == 0) kf = 0;   // empty feature cloud: process_frame returns before touching anything (Estimator.cpp:131-134)
  pose_to_T16(result, out->pose);
  out->keyframe = kf;
  out->icp_status = status;
} };

// The scan's read-back in ONE launch: the ICP state header, the pose/keyframe block, the map counters and the feature count are written
// straight into the context's page-locked host mirrors (cudaMallocHost memory is device-addressable under UVA), instead of four D2H
// copy nodes at the end of the replayed graph (a small copy costs more than a small kernel there).  ~0.3 KB over PCIe, posted writes.
struct k_odom_readback { static __device__ __forceinline__ void run(const IcpState* __restrict__ st, int icp_words, const OdomDev* __restrict__ out, const int* __restrict__ ctr,
                                const int* __restrict__ d_nfeat, int* h_icp, int* h_out, int* h_counts) { TL_START();
  const int t = threadIdx.x;
  const int* a = reinterpret_cast<const int*>(st);
  for (int i = t; i < icp_words; i += blockDim.x) h_icp[i] = a[i];
  const int* b = reinterpret_cast<const int*>(out);
  for (int i = t; i < (int)(sizeof(OdomDev) / sizeof(int)); i += blockDim.x) h_out[i] = b[i];
  if (t < 8) h_counts[t] = ctr[t];
  if (t == 8) h_counts[32] = *d_nfeat;
  __threadfence_system();
} };

extern "C" void b2lo_default_odom_cfg(b2lo_odom_cfg* c, int mid360) {  // config/kitti.yaml / config/mid360.yaml
  if (!c) return;
  c->voxel_size = mid360 ? 0.4f : 0.5f;
  c->point_stride = mid360 ? 4 : 8;
  c->map_voxel_size = mid360 ? 0.4f : 0.5f;
  c->max_range = 100.0;
  c->surfel_planarity_threshold = 0.1f;
  c->keyframe_distance_threshold = 1.0;
  c->keyframe_rotation_threshold = 0.3;
  b2lo_default_icp_cfg(&c->icp);
  if (mid360) c->icp.use_surfel_correspondence = 0;
}

extern "C" int b2lo_odom_create(b2lo_ctx* ctx, const b2lo_odom_cfg* cfg, b2lo_odom** out) {
  if (!ctx || !cfg || !out) return B2LO_E_ARG;
  *out = nullptr;
  b2lo_odom* od = new b2lo_odom();
  od->ctx = ctx;
  od->cfg = *cfg;
  // Estimator.cpp:78-81: VoxelMap(map_voxel_size), hierarchy factor 3, planarity threshold, surfels iff surfel correspondence
  int rc = b2lo_map_create(ctx, cfg->map_voxel_size, 3, cfg->surfel_planarity_threshold, cfg->icp.use_surfel_correspondence, 1u << 17, &od->map);
  if (rc) { delete od; return rc; }
  od->pose = od->prev_pose = od->velocity = od->last_kf_pose = pose_identity();
  od->allow_graph = getenv("B2LO_NO_GRAPH") == nullptr;
  if (cudaMalloc((void**)&od->d_out, sizeof(OdomDev)) != cudaSuccess || cudaMallocHost((void**)&od->h_out, sizeof(OdomDev)) != cudaSuccess) {
    b2lo_odom_destroy(od);
    return B2LO_E_NOMEM;
  }
  *out = od;
  return B2LO_OK;
}
extern "C" int b2lo_odom_destroy(b2lo_odom* od) {
  if (!od) return B2LO_E_ARG;
  for (cudaGraphExec_t g : od->gexec) if (g) cudaGraphExecDestroy(g);
  if (od->d_out) cudaFree(od->d_out);
  if (od->h_out) cudaFreeHost(od->h_out);
  if (od->map) b2lo_map_destroy(od->map);
  delete od;
  return B2LO_OK;
}
extern "C" b2lo_map* b2lo_odom_map(b2lo_odom* od) { return od ? od->map : nullptr; }

static bool should_create_keyframe(const b2lo_odom* od, const Pose& cur) {  // Estimator.cpp:349-368
  if (od->n_keyframes == 0) return true;
  float d[3] = {cur.t[0] - od->last_kf_pose.t[0], cur.t[1] - od->last_kf_pose.t[1], cur.t[2] - od->last_kf_pose.t[2]};
  double distance = (double)sqrtf(sqn3(d));
  Mat3 rd = so3_project(mat3_mul(so3_project(mat3_t(od->last_kf_pose.R)), cur.R));
  float lg[3];
  so3_log(rd, lg);
  double angle = (double)sqrtf(sqn3(lg));
  return distance > od->cfg.keyframe_distance_threshold || angle > od->cfg.keyframe_rotation_threshold;
}

// create_keyframe, map part (Estimator.cpp:449-470): feature cloud at the optimised pose -> UpdateVoxelMap
static int create_keyframe(b2lo_odom* od, size_t n_cap) {
  b2lo_ctx* ctx = od->ctx;
  float T16[16];
  pose_to_T16(od->pose, T16);
  int rc = ctx_transform(ctx, ctx->d_feat, ctx->d_nfeat, n_cap, T16, ctx->d_world);
  if (rc) return rc;
  float sensor[3] = {od->pose.t[0], od->pose.t[1], od->pose.t[2]};
  double md = od->cfg.max_range * 1.2;  // Estimator.cpp:455
  rc = map_update_dev(od->map, ctx->d_world, ctx->d_nfeat, n_cap, sensor, (float)(md * md));
  if (rc < 0) return rc;
  if (!od->cfg.icp.use_surfel_correspondence) { rc = map_rebuild_knn_locked(od->map); if (rc < 0) return rc; }
  od->last_kf_pose = od->pose;
  od->n_keyframes++;
  return B2LO_OK;
}

// where K1 runs relative to the captured sequence
enum ScanMode { K1_SERIAL = 0 /* this scan's K1, in line */, K1_NEXT = 1 /* the announced next scan's K1, beside the registration */, K1_NONE = 2 };

// the launch sequence of one steady-state scan on the context stream: K1 -> ICP -> pose/keyframe decision -> gated K6 -> read-backs.
// Everything that changes from scan to scan sits in the parameter block, so the very same sequence is what the CUDA graph replays.
// In K1_NEXT mode the filter fields of the block describe the NEXT scan: its K1 runs on the side stream into the other feature set
// while this scan registers (mostly single-CTA, latency-bound kernels) out of set `set`; the two branches join before the read-backs end.
static int enqueue_scan(b2lo_odom* od, size_t flt_ns, size_t cap, bool in_graph, int mode, int set, bool side_stream) {
  b2lo_ctx* ctx = od->ctx;
  b2lo_map* map = od->map;
  cudaStream_t st = ctx->stream;
  int rc = B2LO_OK;
  if (in_graph) B2_CUDA(cudaMemcpyAsync(ctx->d_sp, ctx->h_sp, sizeof(ScanParams), cudaMemcpyHostToDevice, st));
  ctx->sp_preloaded = true;
  bool forked = false;
  if (mode == K1_SERIAL) rc = filter_run(ctx, ctx->h_sp->flt_src, flt_ns, (size_t)ctx->h_sp->flt_stride, od->cfg.voxel_size, set, st);
  else if (mode == K1_NEXT) {
    cudaStream_t fs = side_stream ? ctx->stream2 : st;
    if (fs != st) {
      cudaError_t e = cudaEventRecord(ctx->ev_fork, st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(fs, ctx->ev_fork, 0);
      if (e != cudaSuccess) { ctx->sp_preloaded = false; set_error("odometry: fork failed: %s", cudaGetErrorString(e)); return B2LO_E_CUDA; }
      forked = true;
    }
    rc = filter_run(ctx, ctx->h_sp->flt_src, flt_ns, (size_t)ctx->h_sp->flt_stride, od->cfg.voxel_size, set ^ 1, fs);
    if (forked && cudaEventRecord(ctx->ev_join, fs) != cudaSuccess) rc = rc ? rc : B2LO_E_CUDA;
  }
  float4* feat = ctx->feat(set);
  int* nfeat = ctx->nfeat(set);
  if (!rc) rc = icp_run(map, feat, nfeat, cap, ctx->h_sp->T_init, &od->cfg.icp, false, /*restore_on_failure=*/false);
  ctx->sp_preloaded = false;
  if (!rc) {
    launch<k_odom_decide, 32, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(32)), 0, st, ctx->d_icp, ctx->d_sp, nfeat, od->d_out);
    ctx->launches++;
  }
  if (!rc) {
    const double md = od->cfg.max_range * 1.2;  // Estimator.cpp:455
    const float zero3[3] = {0.0f, 0.0f, 0.0f};
    map->graph_mode = true;   // cull kernels take the live voxel count from the device counter, grids from the capacity
    // the feature cloud is moved to the final pose inside the update's first insert kernel (transform_point_cloud, Estimator.cpp:165-169)
    rc = map_update_dev(map, ctx->d_world, nfeat, cap, zero3, (float)(md * md), 0, &od->d_out->keyframe, od->d_out->pose, feat, od->d_out->pose);
    map->graph_mode = false;
    if (rc > 0) rc = B2LO_OK;
  }
  if (!rc) {
    static_assert(offsetof(IcpState, trace) % sizeof(int) == 0 && sizeof(OdomDev) % sizeof(int) == 0, "read-back copies whole words");
    launch<k_odom_readback, 128, 1>(ctx, dim3((unsigned)(1)), dim3((unsigned)(128)), 0, st, ctx->d_icp, (int)(offsetof(IcpState, trace) / sizeof(int)), od->d_out, map->d.ctr, nfeat,
                                       reinterpret_cast<int*>(ctx->h_icp), reinterpret_cast<int*>(od->h_out), ctx->h_counts);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("odometry: read-back failed: %s", cudaGetErrorString(e)); rc = B2LO_E_CUDA; }
  }
  // always re-join a forked branch, also on errors: a capture must not end with an unjoined stream
  if (forked && cudaStreamWaitEvent(st, ctx->ev_join, 0) != cudaSuccess && !rc) rc = B2LO_E_CUDA;
  return rc;
}

static void sp_set_fmt(ScanParams* sp, const b2lo_record_fmt* f) {
  sp->flt_rec = f ? f->record_bytes : 0u;
  for (int a = 0; a < 3; ++a) sp->flt_off[a] = f ? (&f->off_x)[a] : 0u;
}

// the scan parameter block of one steady-state scan in the context's page-locked mirror: K1 source, motion-model guess (Estimator.cpp:154),
// keyframe thresholds
static void fill_scan_params(b2lo_odom* od, const float* flt_src, size_t flt_ns, size_t flt_stride, const b2lo_record_fmt* fmt) {
  ScanParams* sp = od->ctx->h_sp;
  Pose guess = pose_mul(od->prev_pose, od->velocity);  // Estimator.cpp:154
  Pose init = pose_reproject(guess);
  sp->flt_src = flt_src; sp->flt_ns = (int)flt_ns;
  sp->flt_stride = flt_stride; sp->flt_inv = 1.0f / od->cfg.voxel_size;
  sp_set_fmt(sp, fmt);
  sp->flt_mode = 0;
  pose_to_T16(init, sp->T_init);
  sp->force_scale = 0.0;
  pose_to_T16(guess, sp->decide.guess);
  pose_to_T16(od->last_kf_pose, sp->decide.last_kf);
  sp->decide.ran_icp = 1; sp->decide.n_keyframes = od->n_keyframes;
  sp->decide.kf_dist = od->cfg.keyframe_distance_threshold; sp->decide.kf_rot = od->cfg.keyframe_rotation_threshold;
}

// eff: record format of THIS call's source (nullptr: float-stride cloud, e.g. the staged copy of a pageable image); an announced next
// scan is always read in place and therefore in the odometry's own format
// steady_begin enqueues the whole scan (graph replay or plain launches) and returns without waiting; steady_finish waits for it and
// absorbs the results.  b2lo_odom_process* call them back to back; b2lo_odom_process_batch_dev begins a scan on every sequence of a
// batch before it finishes the first one, so that independent sequences overlap on the GPU without a host thread per sequence.
static int steady_begin(b2lo_odom* od, const float* src_dev, size_t ns, size_t sample_stride_floats, double t0, const b2lo_record_fmt* eff) {
  b2lo_ctx* ctx = od->ctx;
  b2lo_map* map = od->map;
  cudaStream_t st = ctx->stream;
  int* hc = ctx->h_counts + 32;
  int rc;
  if (map->n0 == 0) {  // no local map yet (Estimator.cpp:279-285): the motion model alone; cannot happen after the first keyframe
    set_error("odometry: the map is empty after initialisation");
    return B2LO_E_ARG;
  }
  // look-ahead bookkeeping: was this scan's K1 already done by the previous call, and is a next scan announced?
  const bool have_pre = od->pre_valid && od->pre_src == src_dev && od->pre_ns == ns && od->pre_stride == sample_stride_floats;
  const int set = have_pre ? od->pre_set : 0;
  const bool want_next = od->la_valid && od->la_ns > 0;
  const float* nx_src = od->la_src; const size_t nx_ns = od->la_ns, nx_stride = od->la_stride;
  od->pre_valid = false; od->la_valid = false;
  if (have_pre) od->lookahead_hits++;
  // capacities first (may reallocate and synchronise): nothing below allocates
  if ((rc = ctx_reserve_points(ctx, want_next && nx_ns > ns ? nx_ns : ns))) return rc;
  const size_t cap = ctx->pts_cap;
  if ((rc = map_reserve(map, map->n0 + cap, cap))) return rc;
  if ((rc = icp_prepare(ctx, &od->cfg.icp))) return rc;
  const int mode = want_next ? K1_NEXT : (have_pre ? K1_NONE : K1_SERIAL);
  if (!have_pre && mode != K1_SERIAL) {   // first scan of a pipelined run: this scan's K1 in line, ahead of the replayed sequence
    if ((rc = filter_run(ctx, src_dev, ns, sample_stride_floats, od->cfg.voxel_size, 0, st, eff))) return rc;
  }
  // the parameter block of this scan
  if ((rc = sp_begin_write(ctx))) return rc;
  const size_t flt_ns = mode == K1_NEXT ? nx_ns : ns;
  fill_scan_params(od, mode == K1_NEXT ? nx_src : src_dev, flt_ns, mode == K1_NEXT ? nx_stride : sample_stride_floats,
                   mode == K1_NEXT ? (od->has_fmt ? &od->fmt : nullptr) : eff);
  int l2 = 12;
  while ((1ull << l2) < 2 * flt_ns) ++l2;
  if (mode == K1_NONE) l2 = 0;
  const unsigned long long sig[4] = {ctx->alloc_epoch, map->alloc_epoch, (unsigned long long)l2, (unsigned long long)cap};
  const bool profiling = ctx->prof && ctx->prof->on;
  bool use_graph = od->allow_graph && !profiling && od->n_keyframes >= 1;
  const int gi = 2 * mode + set;
  if (use_graph && (!od->gexec[gi] || std::memcmp(sig, od->gsig[gi], sizeof sig) != 0)) {
    if (od->gexec[gi]) { cudaGraphExecDestroy(od->gexec[gi]); od->gexec[gi] = nullptr; }
    cudaGraph_t g = nullptr;
    long long launches0 = ctx->launches;
    B2_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_scan(od, flt_ns, cap, true, mode, set, true);
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    const long long in_capture = ctx->launches - launches0;
    ctx->launches = launches0;
    if (rc || ce != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      od->allow_graph = false;  // fall back to plain stream launches for good
      use_graph = false;
    } else {
      ce = cudaGraphInstantiate(&od->gexec[gi], g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { od->gexec[gi] = nullptr; od->allow_graph = false; use_graph = false; cudaGetLastError(); }
      else { std::memcpy(od->gsig[gi], sig, sizeof sig); od->graph_builds++; od->launches_per_graph = in_capture; }
    }
  }
  if (use_graph) {
    B2_CUDA(cudaGraphLaunch(od->gexec[gi], st));
    ctx->launches += od->launches_per_graph;
    od->graph_launches++;
  } else {
    if ((rc = sp_upload(ctx, 0, sizeof(ScanParams)))) return rc;
    if ((rc = enqueue_scan(od, flt_ns, cap, false, mode, set, !profiling))) return rc;
  }
  ctx->feat_set = set;
  const double t1 = now_us();
  ctx->host_us[1] += t1 - t0;
  od->pend.active = true; od->pend.mode = mode; od->pend.set = set; od->pend.nx_src = nx_src; od->pend.nx_ns = nx_ns; od->pend.nx_stride = nx_stride;
  od->pend.t1 = t1;
  return B2LO_OK;
}

static int steady_finish(b2lo_odom* od, b2lo_odom_result* res) {
  b2lo_ctx* ctx = od->ctx;
  b2lo_map* map = od->map;
  cudaStream_t st = ctx->stream;
  int* hc = ctx->h_counts + 32;
  int rc;
  if (!od->pend.active) { set_error("odometry: no scan in flight"); return B2LO_E_ARG; }
  od->pend.active = false;
  const int mode = od->pend.mode, set = od->pend.set;
  const float* nx_src = od->pend.nx_src; const size_t nx_ns = od->pend.nx_ns, nx_stride = od->pend.nx_stride;
  if (!od->pend.synced) B2_CUDA(cudaStreamSynchronize(st));
  od->pend.synced = false;
  double t2 = now_us();
  ctx->host_us[2] += t2 - od->pend.t1;
  ctx->d2h_bytes += offsetof(IcpState, trace) + sizeof(int) + sizeof(OdomDev) + 8 * sizeof(int);
  if (mode == K1_NEXT) { od->pre_valid = true; od->pre_src = nx_src; od->pre_ns = nx_ns; od->pre_stride = nx_stride; od->pre_set = set ^ 1; }
  res->n_features = hc[0];
  if (hc[0] == 0) {   // nothing was changed: the gated update was switched off, the pose state is untouched
    res->icp_status = B2LO_S_EMPTY;
    pose_to_T16(od->pose, res->pose);
    res->l0 = map->n0; res->l1 = map->n1;
    return B2LO_S_EMPTY;
  }
  res->icp_status = od->h_out->icp_status;
  res->n_corr = ctx->h_icp->n_corr; res->n_iters = ctx->h_icp->num_iterations;
  od->pose = pose_from_T16(od->h_out->pose);
  od->velocity = pose_mul(pose_inv(od->prev_pose), od->pose);  // :177 (prev_pose = what m_previous_frame->get_pose() returns, see below)
  // Estimator.cpp:186-190 + LidarFrame.cpp:113-128: the next scan reads m_previous_frame->get_pose() - this frame's stored pose if it
  // becomes a keyframe, else last_keyframe.pose * (last_keyframe.pose^-1 * pose): the same pose up to the f32 rounding of two products,
  // which reaches the next motion-model guess and velocity (pinned against the real Estimator through the oracle pipeline)
  const Pose rel_to_kf = pose_mul(pose_inv(od->last_kf_pose), od->pose);
  const Pose kf_before = od->last_kf_pose;
  if (od->h_out->keyframe) {
    ctx->host_us[5] += ctx->h_counts[6]; ctx->host_us[6] += ctx->h_counts[7]; ctx->host_us[7] += 1.0;   // purge statistics
    rc = map_absorb_counts(map);
    if (rc < 0) return rc;
    if (!od->cfg.icp.use_surfel_correspondence) map_rebuild_knn_locked(map);
    od->last_kf_pose = od->pose;
    od->n_keyframes++;
    res->keyframe = 1;
  }
  ctx->host_us[3] += now_us() - t2;
  od->prev_pose = od->h_out->keyframe ? od->pose : pose_mul(kf_before, rel_to_kf);
  return B2LO_OK;
}

static int process_common(b2lo_odom* od, const float* src_dev, size_t ns, size_t sample_stride_floats, b2lo_odom_result* res,
                          const b2lo_record_fmt* eff, bool* ev1_recorded = nullptr) {
  b2lo_ctx* ctx = od->ctx;
  b2lo_map* map = od->map;
  cudaStream_t st = ctx->stream;
  std::memset(res, 0, sizeof *res);
  double t0 = now_us();
  int rc = B2LO_OK;
  int* hc = ctx->h_counts + 32;
  if (!od->initialized) {  // initialize_first_frame
    od->la_valid = false;   // the first frame never runs beside anything: an announcement made before it is dropped
    rc = filter_run(ctx, src_dev, ns, sample_stride_floats, od->cfg.voxel_size, 0, nullptr, eff);
    if (rc) return rc;
    B2_CUDA(cudaMemcpyAsync(hc, ctx->d_nfeat, sizeof(int), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    ctx->d2h_bytes += sizeof(int);
    res->n_features = hc[0];
    if (hc[0] == 0) { res->icp_status = B2LO_S_EMPTY; pose_to_T16(od->pose, res->pose); return B2LO_S_EMPTY; }  // Estimator.cpp:131-134
    od->pose = pose_identity();
    od->velocity = pose_identity();
    rc = create_keyframe(od, ns);
    if (rc) return rc;
    od->prev_pose = od->pose;
    od->initialized = true;
    res->keyframe = 1;
    res->icp_status = B2LO_S_EMPTY;
  } else {
    rc = steady_begin(od, src_dev, ns, sample_stride_floats, t0, eff);
    // the end-of-scan timing event goes in behind the launch, so that steady_finish's one stream synchronisation covers it too
    if (!rc && ev1_recorded && cudaEventRecord(ctx->ev1, st) == cudaSuccess) *ev1_recorded = true;
    if (!rc) rc = steady_finish(od, res);
    if (rc) return rc;
  }
  pose_to_T16(od->pose, res->pose);
  res->l0 = map->n0; res->l1 = map->n1;
  return B2LO_OK;
}

static int process_timed(b2lo_odom* od, const float* src_dev, size_t ns, size_t sstride, b2lo_odom_result* res, bool ev0_recorded,
                         const b2lo_record_fmt* eff = nullptr) {
  b2lo_ctx* ctx = od->ctx;
  if (!ev0_recorded) B2_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  bool ev1_done = false;
  int rc = process_common(od, src_dev, ns, sstride, res, eff, &ev1_done);
  if (rc < 0) return rc;
  if (!ev1_done) {   // first frame / paths that synchronise on their own
    B2_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    B2_CUDA(cudaEventSynchronize(ctx->ev1));
  }
  cudaEventElapsedTime(&res->device_ms, ctx->ev0, ctx->ev1);
  return rc;
}

extern "C" int b2lo_odom_process(b2lo_odom* od, const float* xyz, size_t n, size_t stride_floats, b2lo_odom_result* res) {
  if (!od || !res) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  if (!xyz || n == 0) { std::memset(res, 0, sizeof *res); return B2LO_S_EMPTY; }
  b2lo_ctx* ctx = od->ctx;
  std::lock_guard<std::recursive_mutex> lk(od->map->mu);
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  const size_t S = (size_t)(od->cfg.point_stride < 1 ? 1 : od->cfg.point_stride);
  const size_t ns = (n + S - 1) / S;
  const b2lo_record_fmt* F = od->has_fmt ? &od->fmt : nullptr;   // record streams: strides below are in bytes
  const size_t unit = F ? (size_t)F->record_bytes : stride_floats;
  B2_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  double tg = now_us();
  // Page-locked caller memory (cudaMallocHost / cudaHostRegister / torch pin_memory): DMA the raw scan as it is, asynchronously,
  // and let K1 read it strided on the device - no CPU pass over the scan at all.
  cudaPointerAttributes attr;
  bool pinned = (cudaPointerGetAttributes(&attr, xyz) == cudaSuccess) && attr.type == cudaMemoryTypeHost;
  if (!pinned) cudaGetLastError();
  if (pinned && attr.devicePointer && !getenv("B2LO_NO_ZERO_COPY")) {
    // zero-copy: K1 reads just the sampled points (one 32 B sector each) straight out of the caller's page-locked buffer over
    // PCIe; nothing else of the ~1.9 MB scan ever crosses the bus and no staging copy exists
    ctx->h2d_bytes += ns * 32;
    ctx->host_us[0] += now_us() - tg;
    return process_timed(od, static_cast<const float*>(attr.devicePointer), ns, unit * S, res, true, F);
  }
  if (pinned) {
    const size_t floats = F ? (n * unit + 3) / 4 : n * stride_floats;
    if (floats > ctx->raw_floats) {
      B2_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->d_raw) cudaFree(ctx->d_raw);
      ctx->d_raw = nullptr; ctx->raw_floats = 0;
      size_t cap = floats + floats / 4 + 1024;
      if (cudaMalloc((void**)&ctx->d_raw, cap * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc(raw scan, %zu B) failed", cap * sizeof(float)); return B2LO_E_NOMEM; }
      ctx->raw_floats = cap;
    }
    B2_CUDA(cudaMemcpyAsync(ctx->d_raw, xyz, floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += floats * sizeof(float);
    ctx->host_us[0] += now_us() - tg;
    return process_timed(od, ctx->d_raw, ns, unit * S, res, true, F);
  }
  // pageable memory: only every S-th point is ever read by the filter (VoxelMap.h:81): gather those into pinned staging, one H2D
  int rc = F ? ctx_stage_records_h2d(ctx, xyz, n, F, S) : ctx_stage_h2d(ctx, xyz, n, stride_floats, S, nullptr, nullptr);
  ctx->host_us[0] += now_us() - tg;
  if (rc) return rc;
  return process_timed(od, ctx->d_stage, ns, 3, res, true);
}

extern "C" int b2lo_odom_process_dev(b2lo_odom* od, const float* xyz_dev, size_t n, size_t stride_floats, b2lo_odom_result* res) {
  if (!od || !res) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  if (!xyz_dev || n == 0) { std::memset(res, 0, sizeof *res); return B2LO_S_EMPTY; }
  b2lo_ctx* ctx = od->ctx;
  std::lock_guard<std::recursive_mutex> lk(od->map->mu);
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  const size_t S = (size_t)(od->cfg.point_stride < 1 ? 1 : od->cfg.point_stride);
  const size_t ns = (n + S - 1) / S;
  const b2lo_record_fmt* F = od->has_fmt ? &od->fmt : nullptr;
  return process_timed(od, xyz_dev, ns, (F ? (size_t)F->record_bytes : stride_floats) * S, res, false, F);
}

// One scan on each of `count` independent sequences (own b2lo_odom, own context and stream each), all scans already in HBM: every
// sequence's launch sequence is enqueued before the first one is waited for.
extern "C" int b2lo_odom_process_batch_dev(b2lo_odom* const* ods, const float* const* xyz_dev, const size_t* n, const float* const* next_xyz_dev,
                                           const size_t* next_n, size_t stride_floats, int count, b2lo_odom_result* res) {
  if (!ods || !xyz_dev || !n || !res || count < 0) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  for (int a = 0; a < count; ++a) {
    if (!ods[a]) return B2LO_E_ARG;
    for (int b = 0; b < a; ++b) if (ods[a] == ods[b] || ods[a]->ctx == ods[b]->ctx) { set_error("batch: sequences must not share a handle or a context"); return B2LO_E_ARG; }
  }
  std::vector<int> state((size_t)count, 0);   // 0 nothing in flight, 1 begun
  int first_err = B2LO_OK, soft = B2LO_OK;
  static const bool dbg = getenv("B2LO_BATCH_DEBUG") != nullptr;
  static double acc_begin = 0.0, acc_finish = 0.0, acc_wait = 0.0; static long long acc_calls = 0;
  const double tb0 = now_us();
  // phase 1: enqueue
  for (int a = 0; a < count; ++a) {
    b2lo_odom* od = ods[a];
    b2lo_ctx* ctx = od->ctx;
    std::memset(&res[a], 0, sizeof res[a]);
    if (!xyz_dev[a] || n[a] == 0) { soft = B2LO_S_EMPTY; continue; }
    od->map->mu.lock(); ctx->mu.lock();
    cudaSetDevice(ctx->device);
    const size_t S = (size_t)(od->cfg.point_stride < 1 ? 1 : od->cfg.point_stride);
    const size_t ns = (n[a] + S - 1) / S;
    const b2lo_record_fmt* F = od->has_fmt ? &od->fmt : nullptr;
    const size_t sstride = (F ? (size_t)F->record_bytes : stride_floats) * S;
    int rc;
    if (!od->initialized) {   // the first frame of a sequence is synchronous (map build, no ICP)
      rc = process_timed(od, xyz_dev[a], ns, sstride, &res[a], false, F);
      ctx->mu.unlock(); od->map->mu.unlock();
      if (rc < 0 && !first_err) first_err = rc; else if (rc > 0) soft = rc;
      continue;
    }
    if (next_xyz_dev && next_n && next_xyz_dev[a] && next_n[a]) {   // b2lo_odom_lookahead(od, next, n, stride, on_device = 1)
      od->la_src = next_xyz_dev[a]; od->la_ns = (next_n[a] + S - 1) / S; od->la_stride = sstride; od->la_valid = true;
    }
    if (cudaEventRecord(ctx->ev0, ctx->stream) != cudaSuccess) rc = B2LO_E_CUDA;
    else rc = steady_begin(od, xyz_dev[a], ns, sstride, now_us(), F);
    if (!rc && cudaEventRecord(ctx->ev1, ctx->stream) != cudaSuccess) rc = B2LO_E_CUDA;
    if (rc) {
      od->pend.active = false;
      ctx->mu.unlock(); od->map->mu.unlock();
      if (rc < 0 && !first_err) first_err = rc;
      continue;
    }
    state[(size_t)a] = 1;
  }
  const double tb1 = now_us();
  double wait0 = 0.0;
  for (int a = 0; a < count; ++a) wait0 -= ods[a]->ctx->host_us[2];
  // phase 2: wait and absorb, in the same order
  for (int a = 0; a < count; ++a) {
    if (!state[(size_t)a]) continue;
    b2lo_odom* od = ods[a];
    b2lo_ctx* ctx = od->ctx;
    cudaSetDevice(ctx->device);
    int rc = steady_finish(od, &res[a]);
    if (rc >= 0) {
      if (cudaEventSynchronize(ctx->ev1) == cudaSuccess) cudaEventElapsedTime(&res[a].device_ms, ctx->ev0, ctx->ev1);
      pose_to_T16(od->pose, res[a].pose);
      res[a].l0 = od->map->n0; res[a].l1 = od->map->n1;
    }
    ctx->mu.unlock(); od->map->mu.unlock();
    if (rc < 0 && !first_err) first_err = rc; else if (rc > 0) soft = rc;
  }
  if (dbg) {
    for (int a = 0; a < count; ++a) wait0 += ods[a]->ctx->host_us[2];
    const double tb2 = now_us();
    acc_begin += tb1 - tb0; acc_finish += tb2 - tb1; acc_wait += wait0;
    if (++acc_calls % 50 == 0)
      std::fprintf(stderr, "[b2lo batch] %d sequences: enqueue %.1f us/call, finish %.1f us/call of which stream waits %.1f us\n", count, acc_begin / acc_calls,
                   acc_finish / acc_calls, acc_wait / acc_calls), acc_begin = acc_finish = acc_wait = 0.0, acc_calls = 0;
  }
  return first_err ? first_err : soft;
}

// ---- lock-step batches (b2lo_launch.cuh): S independent sequences advance by one scan per replay of ONE graph ------------------------
// Every kernel of the scan is started once per STEP with blockIdx.y = sequence; its per-sequence arguments (map descriptor, buffers, state
// blocks) sit in a device array recorded from the unchanged per-sequence host code.  Results per sequence are those of the sequence
// processed alone, bit for bit (same kernels, same grids in x, same arguments).
namespace b2 {
struct k_ls_params {   // the S scan parameter blocks, from the contexts' page-locked mirrors to their device blocks (zero-copy reads)
  static __device__ __forceinline__ void run(const int* const* h_sp, int* const* d_sp, int words) {
    const int* src = h_sp[blockIdx.x];
    int* dst = d_sp[blockIdx.x];
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  }
};
}  // namespace b2
struct b2lo_lockstep {
  std::vector<b2lo_odom*> ods;
  cudaStream_t st = nullptr;
  static constexpr int MAXB = 8;          // branches of one step's graph (see lockstep_build)
  cudaStream_t side[MAXB] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAXB] = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<unsigned long long> sig;
  std::vector<void*> d_packs;            // one device array of S argument packs per step of the recorded sequence
  const int** d_hsp = nullptr; int** d_dsp = nullptr;
  long long kernels_per_step = 0, replays = 0, builds = 0, fallbacks = 0;
  long long kernels_per_seq = 0;
  int branches = 1;
};
extern "C" int b2lo_lockstep_create(b2lo_odom* const* ods, int count, b2lo_lockstep** out) {
  if (!ods || !out || count < 1) return B2LO_E_ARG;
  *out = nullptr;
  for (int a = 0; a < count; ++a) {
    if (!ods[a]) return B2LO_E_ARG;
    if (ods[a]->ctx->device != ods[0]->ctx->device) { set_error("lock-step: all sequences must live on one device"); return B2LO_E_ARG; }
    if (!ods[a]->cfg.icp.use_surfel_correspondence) { set_error("lock-step: surfel correspondence mode only"); return B2LO_E_ARG; }
    for (int b = 0; b < a; ++b) if (ods[a] == ods[b] || ods[a]->ctx == ods[b]->ctx) { set_error("lock-step: sequences must not share a handle or a context"); return B2LO_E_ARG; }
  }
  cudaSetDevice(ods[0]->ctx->device);
  b2lo_lockstep* ls = new b2lo_lockstep();
  ls->ods.assign(ods, ods + count);
  if (cudaStreamCreateWithFlags(&ls->st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ls->ev0) != cudaSuccess || cudaEventCreate(&ls->ev1) != cudaSuccess ||
      cudaMalloc((void**)&ls->d_hsp, sizeof(void*) * count) != cudaSuccess || cudaMalloc((void**)&ls->d_dsp, sizeof(void*) * count) != cudaSuccess) {
    set_error("lock-step: allocation failed");
    delete ls;
    return B2LO_E_NOMEM;
  }
  bool ok = cudaEventCreateWithFlags(&ls->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int g = 1; g < b2lo_lockstep::MAXB && ok; ++g)
    ok = cudaStreamCreateWithFlags(&ls->side[g], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&ls->ev_join[g], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { set_error("lock-step: stream / event creation failed"); b2lo_lockstep_destroy(ls); return B2LO_E_NOMEM; }
  std::vector<const int*> h((size_t)count); std::vector<int*> d((size_t)count);
  for (int a = 0; a < count; ++a) { h[(size_t)a] = reinterpret_cast<const int*>(ods[a]->ctx->h_sp); d[(size_t)a] = reinterpret_cast<int*>(ods[a]->ctx->d_sp); }
  cudaMemcpy(ls->d_hsp, h.data(), sizeof(void*) * count, cudaMemcpyHostToDevice);
  cudaMemcpy(ls->d_dsp, d.data(), sizeof(void*) * count, cudaMemcpyHostToDevice);
  *out = ls;
  return B2LO_OK;
}
static void lockstep_drop_graph(b2lo_lockstep* ls) {
  if (ls->exec) { cudaGraphExecDestroy(ls->exec); ls->exec = nullptr; }
  for (void* p : ls->d_packs) if (p) cudaFree(p);
  ls->d_packs.clear();
}
extern "C" int b2lo_lockstep_destroy(b2lo_lockstep* ls) {
  if (!ls) return B2LO_E_ARG;
  cudaSetDevice(ls->ods[0]->ctx->device);
  if (ls->st) cudaStreamSynchronize(ls->st);
  lockstep_drop_graph(ls);
  if (ls->d_hsp) cudaFree(ls->d_hsp);
  if (ls->d_dsp) cudaFree(ls->d_dsp);
  if (ls->ev0) cudaEventDestroy(ls->ev0);
  if (ls->ev1) cudaEventDestroy(ls->ev1);
  if (ls->ev_fork) cudaEventDestroy(ls->ev_fork);
  for (int g = 1; g < b2lo_lockstep::MAXB; ++g) { if (ls->ev_join[g]) cudaEventDestroy(ls->ev_join[g]); if (ls->side[g]) cudaStreamDestroy(ls->side[g]); }
  if (ls->st) cudaStreamDestroy(ls->st);
  delete ls;
  return B2LO_OK;
}
extern "C" int b2lo_lockstep_stats(b2lo_lockstep* ls, long long* kernels_per_step, long long* replays, long long* builds, long long* fallbacks) {
  if (!ls) return B2LO_E_ARG;
  if (kernels_per_step) *kernels_per_step = ls->kernels_per_step;
  if (replays) *replays = ls->replays;
  if (builds) *builds = ls->builds;
  if (fallbacks) *fallbacks = ls->fallbacks;
  return B2LO_OK;
}
// record the launch sequence of one scan for every sequence, zip the lists and capture one batched launch per step
static int lockstep_build(b2lo_lockstep* ls, const std::vector<size_t>& flt_ns, const std::vector<size_t>& caps) {
  const int S = (int)ls->ods.size();
  lockstep_drop_graph(ls);
  std::vector<Recorder> recs((size_t)S);
  int rc = B2LO_OK;
  for (int a = 0; a < S && !rc; ++a) {
    b2lo_odom* od = ls->ods[a];
    od->ctx->rec = &recs[(size_t)a];
    od->ctx->batch_S = S;
    rc = enqueue_scan(od, flt_ns[(size_t)a], caps[(size_t)a], /*in_graph=*/false, K1_SERIAL, 0, false);
    od->ctx->rec = nullptr;
    od->ctx->batch_S = 0;
  }
  if (rc) return rc;
  const size_t steps = recs[0].recs.size();
  for (int a = 1; a < S; ++a) {
    if (recs[(size_t)a].recs.size() != steps) { set_error("lock-step: sequences record different launch sequences (%zu vs %zu kernels)", recs[(size_t)a].recs.size(), steps); return B2LO_E_ARG; }
    for (size_t k = 0; k < steps; ++k) {
      const LaunchRec& x = recs[0].recs[k]; const LaunchRec& y = recs[(size_t)a].recs[k];
      if (x.many != y.many || x.block.x != y.block.x || x.smem != y.smem || x.args.size() != y.args.size()) {
        set_error("lock-step: step %zu differs between sequences (kernel / block / shared memory)", k);
        return B2LO_E_ARG;
      }
    }
  }
  ls->d_packs.assign(steps, nullptr);
  std::vector<unsigned char> host;
  for (size_t k = 0; k < steps; ++k) {
    const size_t bytes = recs[0].recs[k].args.size();
    host.resize(bytes * (size_t)S);
    for (int a = 0; a < S; ++a) std::memcpy(host.data() + bytes * (size_t)a, recs[(size_t)a].recs[k].args.data(), bytes);
    B2_CUDA(cudaMalloc(&ls->d_packs[k], host.size()));
    B2_CUDA(cudaMemcpy(ls->d_packs[k], host.data(), host.size(), cudaMemcpyHostToDevice));
  }
  for (size_t k = 0; k < steps; ++k) recs[0].recs[k].prep(recs[0].recs[k].smem);
  cudaGraph_t g = nullptr;
  B2_CUDA(cudaStreamBeginCapture(ls->st, cudaStreamCaptureModeThreadLocal));
  launch<k_ls_params, 64, 1>(ls->ods[0]->ctx, dim3((unsigned)S), dim3(64), 0, ls->st, (const int* const*)ls->d_hsp, (int* const*)ls->d_dsp, (int)(sizeof(ScanParams) / sizeof(int)));
  // Large batches run as B independent BRANCHES of the one graph (sequences [a0, a1) each, on forked streams): every Gauss-Newton
  // iteration of a branch ends in the slowest of its single-SM PKO fits, and while one branch sits in that tail the kernels of the
  // others fill the GPU - what several batches driven from several host threads achieve, inside one call.  Measured on one B200
  // (tools/lockstep_branches.py): 128 sequences 90 k scans/s unbranched, 101 k / 110 k / 115 k with 2 / 4 / 8 branches; 256 sequences
  // 110 k -> 134 k / 138 k with 4 / 8; 384 sequences 144 k.  One branch per 16 sequences, at most eight.
  int B = S / 16;
  if (B < 1) B = 1;
  if (B > b2lo_lockstep::MAXB) B = b2lo_lockstep::MAXB;
  if (const char* e = std::getenv("B2LO_LOCKSTEP_BRANCHES")) { const int v = std::atoi(e); if (v >= 1 && v <= b2lo_lockstep::MAXB && v <= S) B = v; }   // test hook
  if (B > 1) {
    cudaEventRecord(ls->ev_fork, ls->st);
    for (int b = 1; b < B; ++b) cudaStreamWaitEvent(ls->side[b], ls->ev_fork, 0);
  }
  for (int b = 0; b < B; ++b) {
    const int a0 = (int)((long long)S * b / B), a1 = (int)((long long)S * (b + 1) / B);
    cudaStream_t sb = b == 0 ? ls->st : ls->side[b];
    for (size_t k = 0; k < steps; ++k) {
      dim3 grid = recs[0].recs[k].grid;
      for (int a = 1; a < S; ++a) if (recs[(size_t)a].recs[k].grid.x > grid.x) grid.x = recs[(size_t)a].recs[k].grid.x;   // grid-stride kernels: the widest sequence sets x
      const size_t bytes = recs[0].recs[k].args.size();
      recs[0].recs[k].many(static_cast<const unsigned char*>(ls->d_packs[k]) + bytes * (size_t)a0, a1 - a0, grid, recs[0].recs[k].block, recs[0].recs[k].smem, sb);
    }
  }
  for (int b = 1; b < B; ++b) { cudaEventRecord(ls->ev_join[b], ls->side[b]); cudaStreamWaitEvent(ls->st, ls->ev_join[b], 0); }
  ls->branches = B;
  cudaError_t ce = cudaStreamEndCapture(ls->st, &g);
  if (ce != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); set_error("lock-step: capture failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return B2LO_E_CUDA; }
  ce = cudaGraphInstantiate(&ls->exec, g, 0);
  cudaGraphDestroy(g);
  if (ce != cudaSuccess) { ls->exec = nullptr; set_error("lock-step: graph instantiation failed: %s", cudaGetErrorString(ce)); return B2LO_E_CUDA; }
  ls->kernels_per_step = (long long)steps * ls->branches + 1;
  ls->kernels_per_seq = (long long)steps + 1;
  ls->builds++;
  return B2LO_OK;
}
// One scan on every sequence of the batch, all scans resident in HBM.  device_ms (nullable): CUDA-event time of the whole step.
extern "C" int b2lo_lockstep_process_dev(b2lo_lockstep* ls, const float* const* xyz_dev, const size_t* n, size_t stride_floats, b2lo_odom_result* res,
                                         float* device_ms) {
  if (!ls || !xyz_dev || !n || !res) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  const int S = (int)ls->ods.size();
  if (device_ms) *device_ms = 0.0f;
  cudaSetDevice(ls->ods[0]->ctx->device);
  bool steady = true;
  for (int a = 0; a < S; ++a) {
    b2lo_odom* od = ls->ods[a];
    if (!xyz_dev[a] || n[a] == 0 || !od->initialized || od->n_keyframes < 1 || od->map->n0 == 0 || od->has_fmt || (od->ctx->prof && od->ctx->prof->on)) steady = false;
  }
  if (!steady) {   // first frames (map build), empty scans, record streams, profiling: sequence by sequence through the ordinary path
    ls->fallbacks++;
    int first_err = B2LO_OK, soft = B2LO_OK;
    for (int a = 0; a < S; ++a) {
      std::memset(&res[a], 0, sizeof res[a]);
      if (!xyz_dev[a] || n[a] == 0) { soft = B2LO_S_EMPTY; continue; }
      int rc = b2lo_odom_process_dev(ls->ods[a], xyz_dev[a], n[a], stride_floats, &res[a]);
      if (rc < 0 && !first_err) first_err = rc; else if (rc > 0) soft = rc;
    }
    return first_err ? first_err : soft;
  }
  struct Unlock { b2lo_lockstep* l; int k = 0; ~Unlock() { for (int a = k - 1; a >= 0; --a) { l->ods[a]->ctx->mu.unlock(); l->ods[a]->map->mu.unlock(); } } } guard{ls};
  std::vector<size_t> flt_ns((size_t)S), caps((size_t)S);
  std::vector<unsigned long long> sig;
  int rc = B2LO_OK;
  for (int a = 0; a < S; ++a) {
    b2lo_odom* od = ls->ods[a];
    b2lo_ctx* ctx = od->ctx;
    od->map->mu.lock(); ctx->mu.lock(); guard.k = a + 1;
    std::memset(&res[a], 0, sizeof res[a]);
    od->la_valid = od->pre_valid = false;   // no look-ahead inside a lock-step batch: the S sequences fill the GPU already
    const size_t St = (size_t)(od->cfg.point_stride < 1 ? 1 : od->cfg.point_stride);
    const size_t ns = (n[a] + St - 1) / St;
    if ((rc = ctx_reserve_points(ctx, ns))) return rc;
    const size_t cap = ctx->pts_cap;
    if ((rc = map_reserve(od->map, od->map->n0 + cap, cap))) return rc;
    if ((rc = icp_prepare(ctx, &od->cfg.icp))) return rc;
    if ((rc = sp_begin_write(ctx))) return rc;
    fill_scan_params(od, xyz_dev[a], ns, stride_floats * St, nullptr);
    flt_ns[(size_t)a] = ns; caps[(size_t)a] = cap;
    int l2 = 12;
    while ((1ull << l2) < 2 * ns) ++l2;
    sig.push_back(ctx->alloc_epoch); sig.push_back(od->map->alloc_epoch); sig.push_back((unsigned long long)l2); sig.push_back((unsigned long long)cap);
  }
  if (!ls->exec || sig != ls->sig) {
    for (int a = 0; a < S; ++a) B2_CUDA(cudaStreamSynchronize(ls->ods[a]->ctx->stream));   // whatever the sequences did on their own streams is done
    B2_CUDA(cudaStreamSynchronize(ls->st));
    if ((rc = lockstep_build(ls, flt_ns, caps))) return rc;
    ls->sig = sig;
  }
  B2_CUDA(cudaEventRecord(ls->ev0, ls->st));
  B2_CUDA(cudaGraphLaunch(ls->exec, ls->st));
  B2_CUDA(cudaEventRecord(ls->ev1, ls->st));
  B2_CUDA(cudaStreamSynchronize(ls->st));
  ls->replays++;
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, ls->ev0, ls->ev1);
  if (device_ms) *device_ms = ms;
  int first_err = B2LO_OK, soft = B2LO_OK;
  for (int a = 0; a < S; ++a) {
    b2lo_odom* od = ls->ods[a];
    od->ctx->launches += ls->kernels_per_seq;
    od->ctx->feat_set = 0;
    od->pend.active = true; od->pend.mode = K1_SERIAL; od->pend.set = 0; od->pend.nx_src = nullptr; od->pend.nx_ns = 0; od->pend.nx_stride = 0; od->pend.t1 = now_us();
    od->pend.synced = true;   // the step's stream was synchronised above; the sequence's own stream carried nothing (384 empty waits cost ~0.5 ms per step)
    int r = steady_finish(od, &res[a]);
    if (r >= 0) { pose_to_T16(od->pose, res[a].pose); res[a].l0 = od->map->n0; res[a].l1 = od->map->n1; res[a].device_ms = ms; }
    if (r < 0 && !first_err) first_err = r; else if (r > 0) soft = r;
  }
  return first_err ? first_err : soft;
}

extern "C" int b2lo_odom_lookahead(b2lo_odom* od, const float* xyz_next, size_t n, size_t stride_floats, int on_device) {
  if (!od) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  b2lo_ctx* ctx = od->ctx;
  std::lock_guard<std::recursive_mutex> lk(od->map->mu);
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  od->la_valid = false;
  if (!xyz_next || n == 0) return B2LO_S_EMPTY;
  cudaSetDevice(ctx->device);
  const float* src = xyz_next;
  if (!on_device) {
    cudaPointerAttributes attr;
    const bool pinned = (cudaPointerGetAttributes(&attr, xyz_next) == cudaSuccess) && attr.type == cudaMemoryTypeHost && attr.devicePointer;
    if (!pinned || getenv("B2LO_NO_ZERO_COPY")) { cudaGetLastError(); return B2LO_S_EMPTY; }
    src = static_cast<const float*>(attr.devicePointer);
  }
  const size_t S = (size_t)(od->cfg.point_stride < 1 ? 1 : od->cfg.point_stride);
  od->la_src = src; od->la_ns = (n + S - 1) / S; od->la_stride = (od->has_fmt ? (size_t)od->fmt.record_bytes : stride_floats) * S;
  od->la_valid = true;
  return B2LO_OK;
}

// b2lo_odom_lookahead(next) followed by b2lo_odom_process(xyz) in one call (one boundary crossing per scan for bindings whose calls are
// expensive, e.g. ctypes): next_xyz = nullptr skips the announcement
extern "C" int b2lo_odom_process_la(b2lo_odom* od, const float* xyz, size_t n, size_t stride_floats, const float* next_xyz, size_t next_n,
                                    size_t next_stride_floats, b2lo_odom_result* res) {
  if (!od || !res) return B2LO_E_ARG;
  if (next_xyz && next_n) {
    int rc = b2lo_odom_lookahead(od, next_xyz, next_n, next_stride_floats, 0);
    if (rc < 0) return rc;
  }
  return b2lo_odom_process(od, xyz, n, stride_floats, res);
}

extern "C" int b2lo_odom_set_record_fmt(b2lo_odom* od, const b2lo_record_fmt* fmt) {
  if (!od) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(od->map->mu);
  std::lock_guard<std::recursive_mutex> lk2(od->ctx->mu);
  if (fmt) {
    if (fmt->record_bytes == 0) { set_error("records: no format"); return B2LO_E_ARG; }
    for (int a = 0; a < 3; ++a)
      if ((size_t)(&fmt->off_x)[a] + 4 > fmt->record_bytes) { set_error("records: coordinate offset outside the record"); return B2LO_E_ARG; }
    od->fmt = *fmt;
  }
  od->has_fmt = fmt != nullptr;
  od->la_valid = od->pre_valid = false;   // announcements made under the previous format are dropped
  return B2LO_OK;
}

extern "C" int b2lo_odom_graph_stats(b2lo_odom* od, long long* replays, long long* builds, long long* kernels_per_replay) {
  if (!od) return B2LO_E_ARG;
  if (replays) *replays = od->graph_launches;
  if (builds) *builds = od->graph_builds;
  if (kernels_per_replay) *kernels_per_replay = od->launches_per_graph;
  return B2LO_OK;
}

extern "C" int b2lo_odom_reset(b2lo_odom* od) {
  if (!od) return B2LO_E_ARG;
  int rc = b2lo_map_clear(od->map);
  od->pose = od->prev_pose = od->velocity = od->last_kf_pose = pose_identity();
  od->initialized = false;
  od->n_keyframes = 0;
  od->la_valid = od->pre_valid = false;
  return rc;
}

#ifdef B2LO_TIMELINE
namespace b2 { int tl_fetch_odom(unsigned long long* out, int cap) { return tl_fetch(out, cap); } }
#endif
