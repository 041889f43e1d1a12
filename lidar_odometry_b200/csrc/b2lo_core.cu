// b2lo_core.cu — context, staging, and the C entry points of FastVoxelFilter / ICP / pose algebra.
//
// Reference members replaced by the entry points in this file:
//   FastVoxelFilter::filter                        /root/reference/src/database/VoxelMap.h:73-104
//   IterativeClosestPointOptimizer::optimize       src/optimization/IterativeClosestPointOptimizer.cpp:255-463
//   find_correspondences (parity tap)              :587-645
//   VoxelMap::GetSurfelAtPoint                     src/database/VoxelMap.cpp:368-386
//   util::SE3 / SO3 algebra                        src/util/MathUtils.h:57-168, MathUtils.cpp:23-181
// All device work of a context is issued on ctx->stream.  No CPU fallback exists: without a CUDA device
// b2lo_ctx_create fails with B2LO_E_CUDA and nothing else can be called.
#include <climits>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include "b2lo_internal.h"
#include "b2lo_launch.cuh"

namespace b2 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

void prof_drain(b2lo_ctx* ctx) {
  Prof* P = ctx->prof;
  if (!P || !P->used) return;
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < P->used; ++i) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, P->a[i], P->b[i]) == cudaSuccess) { P->ms[P->slot[i]] += ms; P->n[P->slot[i]] += 1; }
  }
  P->used = 0;
}
void prof_begin(b2lo_ctx* ctx, int slot) {
  Prof* P = ctx->prof;
  if (!P || !P->on) return;
  if (P->used == Prof::POOL) prof_drain(ctx);
  P->slot[P->used] = slot;
  cudaEventRecord(P->a[P->used], ctx->stream);
}
void prof_end(b2lo_ctx* ctx) {
  Prof* P = ctx->prof;
  if (!P || !P->on) return;
  cudaEventRecord(P->b[P->used], ctx->stream);
  P->used++;
}

template <class T> static int dev_alloc(T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu B) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return B2LO_E_NOMEM; }
  return B2LO_OK;
}

// packed host xyz (3 floats) -> float4 stream; also publishes the point count on the device
__global__ void k_pack4(const float* __restrict__ src, int n, float4* dst, int* d_count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst[i] = make_float4(src[i * 3], src[i * 3 + 1], src[i * 3 + 2], 0.0f);
  if (d_count && blockIdx.x == 0 && threadIdx.x == 0) *d_count = n;
}
__global__ void k_unpack3(const float4* __restrict__ src, const int* __restrict__ d_n, int cap, float* dst) {
  int n = d_n ? *d_n : cap;
  if (n > cap) n = cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = src[i];
    dst[i * 3] = p.x; dst[i * 3 + 1] = p.y; dst[i * 3 + 2] = p.z;
  }
}
// util::transform_point_cloud (src/util/PointCloudUtils.cpp:102-125): world = T * (x,y,z,1), f32, no FMA
struct Pose16 { float m[16]; };
__global__ void k_transform(const float4* __restrict__ src, const int* __restrict__ d_n, Pose16 T, float4* dst) {
  const int n = *d_n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = src[i];
    float x = ((T.m[0] * p.x + T.m[1] * p.y) + T.m[2] * p.z) + T.m[3] * 1.0f;
    float y = ((T.m[4] * p.x + T.m[5] * p.y) + T.m[6] * p.z) + T.m[7] * 1.0f;
    float z = ((T.m[8] * p.x + T.m[9] * p.y) + T.m[10] * p.z) + T.m[11] * 1.0f;
    dst[i] = make_float4(x, y, z, 0.0f);
  }
}

__global__ void k_transform_dev(const float4* __restrict__ src, const int* __restrict__ d_n, const float* __restrict__ T, const int* __restrict__ gate,
                                float4* dst) {
  if (gate && !*gate) return;
  const int n = *d_n;
  float m[12];
  for (int i = 0; i < 12; ++i) m[i] = T[i];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = src[i];
    float x = ((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3] * 1.0f;
    float y = ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7] * 1.0f;
    float z = ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11] * 1.0f;
    dst[i] = make_float4(x, y, z, 0.0f);
  }
}

static int grid_of(size_t n, int threads) { size_t b = (n + threads - 1) / threads; if (b < 1) b = 1; if (b > 1184) b = 1184; return (int)b; }

int ctx_reserve_points(b2lo_ctx* ctx, size_t n) {
  if (n <= ctx->pts_cap) return B2LO_OK;
  size_t cap = 16384;
  while (cap < n) cap *= 2;
  if (cap > (size_t)1 << 27) { set_error("point cloud too large (%zu points)", n); return B2LO_E_CAPACITY; }
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  int rc;
  // the feature clouds survive a grow (odometry keeps them across calls; a prefetched scan may sit in the second set)
  if (ctx->stream2) B2_CUDA(cudaStreamSynchronize(ctx->stream2));
  const size_t old_cap = ctx->pts_cap;
  {
    float4** fv[2] = {&ctx->d_feat, &ctx->d_feat2};
    unsigned long long** kv[2] = {&ctx->d_feat_key, &ctx->d_feat_key2};
    for (int s = 0; s < 2; ++s) {
      float4* old_feat = *fv[s]; unsigned long long* old_key = *kv[s];
      *fv[s] = nullptr; *kv[s] = nullptr;
      if ((rc = dev_alloc(fv[s], cap)) || (rc = dev_alloc(kv[s], cap))) return rc;
      if (old_feat) {
        B2_CUDA(cudaMemcpy(*fv[s], old_feat, old_cap * sizeof(float4), cudaMemcpyDeviceToDevice));
        B2_CUDA(cudaMemcpy(*kv[s], old_key, old_cap * sizeof(unsigned long long), cudaMemcpyDeviceToDevice));
        cudaFree(old_feat); cudaFree(old_key);
      }
    }
  }
  if ((rc = dev_alloc(&ctx->d_query, cap)) || (rc = dev_alloc(&ctx->d_world, cap))) return rc;
  int l2 = 4;
  while ((1ull << l2) < 2 * cap) ++l2;
  ctx->f_log2cap = l2;
  if ((rc = dev_alloc(&ctx->f_tab, (size_t)1 << l2)) || (rc = dev_alloc(&ctx->f_samp, cap)) || (rc = dev_alloc(&ctx->f_slot, cap)) ||
      (rc = dev_alloc(&ctx->f_vid, cap)) || (rc = dev_alloc(&ctx->f_segstart, cap + 1)) || (rc = dev_alloc(&ctx->f_segcnt, cap)) ||
      (rc = dev_alloc(&ctx->f_lead, cap)) || (rc = dev_alloc(&ctx->f_bucket, cap)) || (rc = dev_alloc(&ctx->f_ordered, cap)) ||
      (rc = dev_alloc(&ctx->f_packed, cap)) || (rc = dev_alloc(&ctx->f_sorted, cap)))
    return rc;
  // the filter's scratch hash cleans itself after every run (k_flt_reduce); it only needs the idle pattern once
  B2_CUDA(cudaMemsetAsync(ctx->f_tab, 0xFF, sizeof(FEntry) << l2, ctx->stream));
  if ((rc = dev_alloc(&ctx->i_res, cap)) || (rc = dev_alloc(&ctx->i_slot, cap)) || (rc = dev_alloc(&ctx->i_cidx, cap + 1024)) ||
      (rc = dev_alloc(&ctx->i_blkcnt, cap / 256 + 8)) || (rc = dev_alloc(&ctx->i_blkoff, cap / 256 + 8)) ||
      (rc = dev_alloc(&ctx->i_tilesum, 2 * (cap / 256 + 8))))
    return rc;
  if (ctx->h_stage) { cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; }
  ctx->h_stage_floats = cap * 6 + 64;
  B2_CUDA(cudaMallocHost((void**)&ctx->h_stage, ctx->h_stage_floats * sizeof(float)));
  ctx->d_stage_floats = cap * 4 + 64;
  if ((rc = dev_alloc(&ctx->d_stage, ctx->d_stage_floats))) return rc;
  ctx->pts_cap = cap;
  ctx->alloc_epoch++;
  return B2LO_OK;
}

// Host AoS cloud -> pinned staging (every take_every-th point, xyz only) -> device float4 stream.
// The H2D copy is asynchronous on the context stream; the staging buffer is reused only after a sync.
int ctx_stage_h2d(b2lo_ctx* ctx, const float* xyz, size_t n, size_t stride_floats, size_t take_every, float4* dst, int* d_count) {
  if (take_every < 1) take_every = 1;
  size_t nt = (n + take_every - 1) / take_every;
  int rc = ctx_reserve_points(ctx, nt);
  if (rc) return rc;
  if (ctx->stage_busy) { B2_CUDA(cudaEventSynchronize(ctx->ev_stage)); ctx->stage_busy = false; }
  float* h = ctx->h_stage;
  const size_t step = stride_floats * take_every;
  if (stride_floats == 3 && take_every == 1) std::memcpy(h, xyz, nt * 3 * sizeof(float));
  else for (size_t j = 0; j < nt; ++j) { const float* p = xyz + j * step; h[j * 3] = p[0]; h[j * 3 + 1] = p[1]; h[j * 3 + 2] = p[2]; }
  if (nt) B2_CUDA(cudaMemcpyAsync(ctx->d_stage, h, nt * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaEventRecord(ctx->ev_stage, ctx->stream));
  ctx->stage_busy = true;
  ctx->h2d_bytes += nt * 3 * sizeof(float);
  if (dst) {
    k_pack4<<<grid_of(nt, 256), 256, 0, ctx->stream>>>(ctx->d_stage, (int)nt, dst, d_count);
    ctx->launches++;
  }
  return B2LO_OK;
}

int sp_begin_write(b2lo_ctx* ctx) {
  if (ctx->sp_busy) { B2_CUDA(cudaEventSynchronize(ctx->ev_sp)); ctx->sp_busy = false; }
  return B2LO_OK;
}
int sp_upload(b2lo_ctx* ctx, size_t offset, size_t bytes) {
  B2_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(ctx->d_sp) + offset, reinterpret_cast<const char*>(ctx->h_sp) + offset, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaEventRecord(ctx->ev_sp, ctx->stream));
  ctx->sp_busy = true;
  return B2LO_OK;
}

int ctx_transform(b2lo_ctx* ctx, const float4* src, const int* d_n, size_t n_cap, const float T16[16], float4* dst) {
  Pose16 T;
  std::memcpy(T.m, T16, sizeof T.m);
  prof_begin(ctx, PS_XFORM);
  k_transform<<<grid_of(n_cap, 256), 256, 0, ctx->stream>>>(src, d_n, T, dst);
  prof_end(ctx);
  ctx->launches++;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}

int ctx_transform_dev(b2lo_ctx* ctx, const float4* src, const int* d_n, size_t n_cap, const float* T16_dev, const int* gate, float4* dst) {
  prof_begin(ctx, PS_XFORM);
  k_transform_dev<<<grid_of(n_cap, 256), 256, 0, ctx->stream>>>(src, d_n, T16_dev, gate, dst);
  prof_end(ctx);
  ctx->launches++;
  B2_CUDA(cudaGetLastError());
  return B2LO_OK;
}

// copy the float4 feature stream [0, *d_n) to the host as packed xyz; one synchronisation
int ctx_read_cloud(b2lo_ctx* ctx, const float4* src, const int* d_n, size_t n_cap, float* out_xyz, size_t out_cap, size_t* n_out) {
  if (n_cap > ctx->pts_cap) n_cap = ctx->pts_cap;
  if (ctx->stage_busy) { B2_CUDA(cudaEventSynchronize(ctx->ev_stage)); ctx->stage_busy = false; }
  k_unpack3<<<grid_of(n_cap ? n_cap : 1, 256), 256, 0, ctx->stream>>>(src, d_n, (int)n_cap, ctx->d_stage);
  ctx->launches++;
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 16, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (n_cap && out_xyz) B2_CUDA(cudaMemcpyAsync(ctx->h_stage, ctx->d_stage, n_cap * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  size_t m = (size_t)ctx->h_counts[16];
  if (m > n_cap) m = n_cap;
  *n_out = m;
  ctx->d2h_bytes += sizeof(int) + (out_xyz ? n_cap * 3 * sizeof(float) : 0);
  if (out_xyz) {
    if (m > out_cap) { set_error("output buffer too small (%zu < %zu points)", out_cap, m); return B2LO_E_CAPACITY; }
    std::memcpy(out_xyz, ctx->h_stage, m * 3 * sizeof(float));
  }
  return B2LO_OK;
}

}  // namespace b2

using namespace b2;

// Independent sequences overlap on the device only if their streams map to different hardware work queues; the driver's default is 8
// connections, which aliases the streams of a batch (measured: 32 sequences 11.4 k -> 30.9 k scans/s with 32).  Takes effect when this
// library is loaded before the process initialises CUDA; an explicit setting in the environment wins.
// Opt-in (bench.py and the batch tools call it before the first CUDA call): a library must not edit its host's environment behind
// its back.
extern "C" int b2lo_process_env_for_batches(int connections) {
  if (connections < 1 || connections > 32) return B2LO_E_ARG;
  char buf[16];
  snprintf(buf, sizeof buf, "%d", connections);
  return setenv("CUDA_DEVICE_MAX_CONNECTIONS", buf, 0) == 0 ? B2LO_OK : B2LO_E_ARG;
}

namespace b2 { Recorder* ctx_recorder(b2lo_ctx* ctx) { return ctx->rec; } }

extern "C" const char* b2lo_version(void) { return "b2lo 0.1 (sm_100a)"; }
extern "C" const char* b2lo_last_error(void) { return g_err; }
extern "C" void b2lo_struct_sizes(size_t out[5]) {
  out[0] = sizeof(b2lo_icp_cfg); out[1] = sizeof(b2lo_iter_trace); out[2] = sizeof(b2lo_icp_stats); out[3] = sizeof(b2lo_odom_cfg);
  out[4] = sizeof(b2lo_odom_result);
}

extern "C" void b2lo_default_icp_cfg(b2lo_icp_cfg* c) {  // Estimator.cpp:49-70 + config/kitti.yaml
  if (!c) return;
  c->max_iterations = 4; c->translation_tolerance = 0.005; c->rotation_tolerance = 0.005; c->max_correspondence_distance = 1.0;
  c->min_correspondence_points = 10; c->use_robust_loss = 1; c->robust_loss_delta = 0.1; c->use_surfel_correspondence = 1;
  c->use_adaptive_m_estimator = 1; c->loss_type = 0; c->min_scale_factor = 0.1; c->max_scale_factor = 10.0; c->num_alpha_segments = 100;
  c->truncated_threshold = 10.0; c->gmm_components = 3; c->gmm_sample_size = 100; c->pko_kernel_type = 0;
}

extern "C" int b2lo_ctx_create(int device, b2lo_ctx** out) {
  if (!out) return B2LO_E_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s) - this engine has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return B2LO_E_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range [0,%d)", device, ndev); return B2LO_E_ARG; }
  B2_CUDA(cudaSetDevice(device));
  b2lo_ctx* ctx = new b2lo_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
      cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    set_error("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return B2LO_E_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->i_max_blocks = prop.multiProcessorCount * 4;
  int rc = B2LO_OK;
  if (cudaMalloc((void**)&ctx->d_nfeat, sizeof(int)) != cudaSuccess || cudaMalloc((void**)&ctx->d_nfeat2, sizeof(int)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_nquery, sizeof(int)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_icp, sizeof(IcpState)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->i_partial, ((size_t)ctx->i_max_blocks * 28 + 320) * sizeof(double)) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_icp, sizeof(IcpState)) != cudaSuccess || cudaMalloc((void**)&ctx->d_sp, sizeof(ScanParams)) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_sp, sizeof(ScanParams)) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_sp, cudaEventDisableTiming) != cudaSuccess || cudaMallocHost((void**)&ctx->h_counts, 64 * sizeof(int)) != cudaSuccess)
    rc = B2LO_E_NOMEM;
  if (!rc) {
    std::memset(ctx->h_sp, 0, sizeof(ScanParams));
    cudaMemsetAsync(ctx->d_sp, 0, sizeof(ScanParams), ctx->stream);
    cudaMemsetAsync(ctx->d_nfeat, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(ctx->d_nfeat2, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(ctx->d_nquery, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(ctx->d_icp, 0, sizeof(IcpState), ctx->stream);
    cudaMemsetAsync(ctx->i_partial, 0, ((size_t)ctx->i_max_blocks * 28 + 320) * sizeof(double), ctx->stream);
    rc = ctx_reserve_points(ctx, 16384);
  }
  if (rc) { set_error("context allocation failed"); b2lo_ctx_destroy(ctx); return rc; }
  cudaStreamSynchronize(ctx->stream);
  *out = ctx;
  return B2LO_OK;
}

extern "C" int b2lo_ctx_destroy(b2lo_ctx* ctx) {
  if (!ctx) return B2LO_E_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->icp_graph_exec) cudaGraphExecDestroy(ctx->icp_graph_exec);
  void* dptrs[] = {ctx->d_raw, ctx->d_stage, ctx->d_feat, ctx->d_feat_key, ctx->d_nfeat, ctx->d_feat2, ctx->d_feat_key2, ctx->d_nfeat2, ctx->d_query, ctx->d_nquery, ctx->d_world, ctx->f_tab, ctx->f_samp,
                   ctx->f_slot, ctx->f_vid, ctx->f_segstart, ctx->f_segcnt, ctx->f_lead, ctx->f_bucket, ctx->f_ordered, ctx->f_packed, ctx->f_sorted, ctx->i_res, ctx->i_slot,
                   ctx->i_cidx, ctx->i_blkcnt, ctx->i_blkoff, ctx->i_tilesum, ctx->i_partial, ctx->d_icp, ctx->d_pko, ctx->d_pko_hits, ctx->d_tap_state,
                   ctx->d_tap_key, ctx->d_tap_morton, ctx->d_tap_n, ctx->d_tap_c, ctx->d_mapdev, ctx->k_idx, ctx->k_n, ctx->k_unres, ctx->k_nunres,
                   ctx->k_plane, ctx->l_cnt};
  for (void* p : dptrs) if (p) cudaFree(p);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_icp) cudaFreeHost(ctx->h_icp);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->h_sp) cudaFreeHost(ctx->h_sp);
  if (ctx->d_sp) cudaFree(ctx->d_sp);
  if (ctx->ev_sp) cudaEventDestroy(ctx->ev_sp);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->prof) { if (ctx->prof->created) for (int i = 0; i < Prof::POOL; ++i) { cudaEventDestroy(ctx->prof->a[i]); cudaEventDestroy(ctx->prof->b[i]); } delete ctx->prof; }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B2LO_OK;
}
extern "C" int b2lo_ctx_sync(b2lo_ctx* ctx) {
  if (!ctx) return B2LO_E_ARG;
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2LO_OK;
}
extern "C" void* b2lo_ctx_stream(b2lo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" long long b2lo_ctx_launch_count(b2lo_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" int b2lo_ctx_profile(b2lo_ctx* ctx, int enable) {
  if (!ctx) return B2LO_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->prof) ctx->prof = new Prof();
  Prof* P = ctx->prof;
  if (!P->created) { for (int i = 0; i < Prof::POOL; ++i) { cudaEventCreate(&P->a[i]); cudaEventCreate(&P->b[i]); } P->created = true; }
  prof_drain(ctx);
  P->on = enable != 0;
  if (enable) for (int i = 0; i < PS_COUNT; ++i) { P->ms[i] = 0.0; P->n[i] = 0; }
  return B2LO_OK;
}
extern "C" int b2lo_ctx_profile_read(b2lo_ctx* ctx, int slot, double* total_ms, long long* launches) {
  if (!ctx || !ctx->prof || slot < 0 || slot >= PS_COUNT) return B2LO_E_ARG;
  prof_drain(ctx);
  if (total_ms) *total_ms = ctx->prof->ms[slot];
  if (launches) *launches = ctx->prof->n[slot];
  return B2LO_OK;
}
extern "C" int b2lo_ctx_debug_clocks(b2lo_ctx* ctx, long long out[32]) {  // clock64 phase stamps of the last optimize (debug aid)
  if (!ctx || !out) return B2LO_E_ARG;
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  B2_CUDA(cudaMemcpy(out, ctx->d_icp->dbg, 32 * sizeof(long long), cudaMemcpyDeviceToHost));
  return B2LO_OK;
}
extern "C" int b2lo_ctx_host_us(b2lo_ctx* ctx, double out[8], int reset) {  // host-side wall-clock split of b2lo_odom_process (debug aid)
  if (!ctx || !out) return B2LO_E_ARG;
  for (int i = 0; i < 8; ++i) { out[i] = ctx->host_us[i]; if (reset) ctx->host_us[i] = 0.0; }
  return B2LO_OK;
}
extern "C" int b2lo_ctx_io_bytes(b2lo_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h) {
  if (!ctx) return B2LO_E_ARG;
  if (h2d) *h2d = ctx->h2d_bytes;
  if (d2h) *d2h = ctx->d2h_bytes;
  return B2LO_OK;
}

// ---- FastVoxelFilter ------------------------------------------------------------------------------------
extern "C" int b2lo_filter(b2lo_ctx* ctx, const float* xyz, size_t n, size_t stride_floats, int stride, float voxel_size, float* out_xyz,
                           uint64_t* out_keys, size_t* m) {
  if (!ctx || !m) return B2LO_E_ARG;
  *m = 0;
  if (stride < 1 || stride_floats < 3 || !(voxel_size > 0.0f)) { set_error("filter: bad stride / voxel size"); return B2LO_E_ARG; }
  if (!xyz || n == 0) return B2LO_S_EMPTY;  // input.empty() (VoxelMap.h:75)
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  size_t ns = (n + (size_t)stride - 1) / (size_t)stride;
  int rc = ctx_stage_h2d(ctx, xyz, n, stride_floats, (size_t)stride, nullptr, nullptr);
  if (rc) return rc;
  rc = filter_run(ctx, ctx->d_stage, ns, 3, voxel_size);
  if (rc) return rc;
  size_t got = 0;
  const size_t koff = (ns * 3 + 3) & ~(size_t)1;  // 8-byte aligned slot behind the xyz block of the pinned staging area
  if (ctx->stage_busy) { B2_CUDA(cudaEventSynchronize(ctx->ev_stage)); ctx->stage_busy = false; }
  if (out_keys) B2_CUDA(cudaMemcpyAsync(ctx->h_stage + koff, ctx->d_feat_key, ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  rc = ctx_read_cloud(ctx, ctx->d_feat, ctx->d_nfeat, ns, out_xyz, ns, &got);
  if (rc) return rc;
  if (out_keys) { std::memcpy(out_keys, ctx->h_stage + koff, got * sizeof(unsigned long long)); ctx->d2h_bytes += ns * sizeof(unsigned long long); }
  *m = got;
  return B2LO_OK;
}
extern "C" int b2lo_filter_dev(b2lo_ctx* ctx, const float* xyz_dev, size_t n, size_t stride_floats, int stride, float voxel_size) {
  if (!ctx) return B2LO_E_ARG;
  if (stride < 1 || stride_floats < 3 || !(voxel_size > 0.0f)) { set_error("filter: bad stride / voxel size"); return B2LO_E_ARG; }
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  if (!xyz_dev || n == 0) { B2_CUDA(cudaMemsetAsync(ctx->d_nfeat, 0, sizeof(int), ctx->stream)); return B2LO_S_EMPTY; }
  size_t ns = (n + (size_t)stride - 1) / (size_t)stride;
  return filter_run(ctx, xyz_dev, ns, stride_floats * (size_t)stride, voxel_size);
}
extern "C" int b2lo_ctx_features(b2lo_ctx* ctx, float* out_xyz, size_t cap, size_t* m) {
  if (!ctx || !m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  size_t ncap = cap < ctx->pts_cap ? cap : ctx->pts_cap;
  if (!out_xyz) ncap = 0;
  // first the count, then at most `cap` points
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 17, ctx->nfeat(ctx->feat_set), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  size_t have = (size_t)ctx->h_counts[17];
  if (!out_xyz) { *m = have; return B2LO_OK; }
  if (have > cap) { *m = have; set_error("features: buffer too small (%zu < %zu)", cap, have); return B2LO_E_CAPACITY; }
  return ctx_read_cloud(ctx, ctx->feat(ctx->feat_set), ctx->nfeat(ctx->feat_set), have, out_xyz, cap, m);
}

// ---- ICP ----------------------------------------------------------------------------------------------------
static void fill_stats(b2lo_ctx* ctx, float T_out[16], b2lo_icp_stats* stats, float ms) {
  const IcpState* h = ctx->h_icp;
  Pose p;
  for (int i = 0; i < 9; ++i) p.R.m[i] = h->R[i];
  for (int i = 0; i < 3; ++i) p.t[i] = h->t[i];
  if (T_out) pose_to_T16(p, T_out);
  if (stats) {
    stats->status = h->status; stats->num_iterations = h->num_iterations; stats->num_correspondences = h->n_corr;
    stats->converged = h->converged; stats->initial_cost = h->initial_cost; stats->final_cost = h->final_cost; stats->device_ms = ms;
    std::memcpy(stats->it, h->trace, sizeof(stats->it));
  }
}

static int icp_optimize_common(b2lo_map* map, const float4* d_pts, const int* d_n, size_t n_cap, const float T_init[16], const b2lo_icp_cfg* cfg,
                               float T_out[16], b2lo_icp_stats* stats) {
  b2lo_ctx* ctx = map->ctx;
  B2_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = icp_run(map, d_pts, d_n, n_cap, T_init, cfg, false);
  if (rc) return rc;
  B2_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(ctx->h_icp, ctx->d_icp, sizeof(IcpState), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->d2h_bytes += sizeof(IcpState);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  fill_stats(ctx, T_out, stats, ms);
  if (ctx->h_icp->status == B2LO_E_CAPACITY) { set_error("more than 2^22 correspondences in one optimize: outside the PKO sample tables"); return B2LO_E_CAPACITY; }
  return ctx->h_icp->status == B2LO_S_INSUFFICIENT ? B2LO_S_INSUFFICIENT : B2LO_OK;
}

extern "C" int b2lo_icp_optimize(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T_init[16],
                                 const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats) {
  if (!map || !T_init || !cfg || !T_out) return B2LO_E_ARG;
  if (stride_floats < 3) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  // empty map / empty cloud: find_correspondences returns 0 (ICP.cpp:593-603) -> below the minimum -> false, output = initial
  if (!local_xyz || m == 0 || map->n0 == 0) {
    std::memcpy(T_out, T_init, 16 * sizeof(float));
    if (stats) { std::memset(stats, 0, sizeof *stats); stats->status = B2LO_S_INSUFFICIENT; }
    return B2LO_S_INSUFFICIENT;
  }
  { int rr = ctx_reserve_points(ctx, m); if (rr) return rr; }  // may reallocate d_query: reserve before taking the pointer
  int rc = ctx_stage_h2d(ctx, local_xyz, m, stride_floats, 1, ctx->d_query, ctx->d_nquery);
  if (rc) return rc;
  return icp_optimize_common(map, ctx->d_query, ctx->d_nquery, m, T_init, cfg, T_out, stats);
}

// parity tap: ONE Gauss-Newton iteration (the loop body ICP.cpp:280-448) entered at T_in with the residual normalisation scale an earlier
// iteration fixed (scale <= 0: computed from this pose's residuals, as iteration 0 does)
extern "C" int b2lo_icp_iterate(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T_in[16], double scale,
                                const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats) {
  if (!map || !cfg) return B2LO_E_ARG;
  b2lo_icp_cfg one = *cfg;
  one.max_iterations = 1;
  {
    std::lock_guard<std::recursive_mutex> lk(map->ctx->mu);
    map->ctx->force_scale = scale > 0.0 ? scale : 0.0;
  }
  int rc = b2lo_icp_optimize(map, local_xyz, m, stride_floats, T_in, &one, T_out, stats);
  map->ctx->force_scale = 0.0;
  return rc;
}

extern "C" int b2lo_icp_optimize_features(b2lo_map* map, const float T_init[16], const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats) {
  if (!map || !T_init || !cfg || !T_out) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(map->mu);
  b2lo_ctx* ctx = map->ctx;
  std::lock_guard<std::recursive_mutex> lk2(ctx->mu);
  cudaSetDevice(ctx->device);
  // the features of the scan filtered LAST: after a look-ahead run they may sit in the second set (what b2lo_ctx_features reads)
  const int set = ctx->feat_set;
  const size_t cap = ctx->feat_cap_hint_set[set];
  if (map->n0 == 0 || cap == 0) {
    std::memcpy(T_out, T_init, 16 * sizeof(float));
    if (stats) { std::memset(stats, 0, sizeof *stats); stats->status = B2LO_S_INSUFFICIENT; }
    return B2LO_S_INSUFFICIENT;
  }
  return icp_optimize_common(map, ctx->feat(set), ctx->nfeat(set), cap, T_init, cfg, T_out, stats);
}

// ---- pose algebra (host; the same inline functions the device uses) ----------------------------------------------
extern "C" void b2lo_se3_mul(const float A16[16], const float B16[16], float C16[16]) {
  Pose a = pose_from_T16(A16), b = pose_from_T16(B16);
  pose_to_T16(pose_mul(a, b), C16);
}
extern "C" void b2lo_se3_inv(const float A16[16], float C16[16]) { pose_to_T16(pose_inv(pose_from_T16(A16)), C16); }
extern "C" void b2lo_se3_from_rt(const float T16_in[16], float T16_out[16]) {
  Pose p = pose_from_T16(T16_in);
  p.R = so3_project(p.R);
  pose_to_T16(p, T16_out);
}
extern "C" void b2lo_so3_log(const float T16[16], float w[3]) { Pose p = pose_from_T16(T16); so3_log(p.R, w); }
extern "C" void b2lo_so3_exp(const float w[3], float R9[9]) { Mat3 R = so3_exp(w); std::memcpy(R9, R.m, sizeof R.m); }
extern "C" void b2lo_svd3(const float A9[9], float U9[9], float S3[3], float V9[9]) {
  Mat3 A, U, V;
  std::memcpy(A.m, A9, sizeof A.m);
  svd3(A, U, S3, V);
  std::memcpy(U9, U.m, sizeof U.m); std::memcpy(V9, V.m, sizeof V.m);
}
extern "C" void b2lo_ldlt6_solve(const float H36[36], const float b6[6], float x6[6]) { ldlt6_solve(H36, b6, x6); }
extern "C" void b2lo_fit_plane(const float* pts, int n, float mu[3], float normal[3], float* planarity) { fit_plane(pts, n, mu, normal, planarity); }
extern "C" uint64_t b2lo_voxel_key_hash(int x, int y, int z) { return key_morton(x, y, z); }

// ---- in-graph timeline (debug builds only, see b2lo_dev.cuh) ---------------------------------------------------------------------
#ifdef B2LO_TIMELINE
namespace b2 { int tl_fetch_filter(unsigned long long*, int); int tl_fetch_icp(unsigned long long*, int); int tl_fetch_odom(unsigned long long*, int);
               int tl_fetch_map(unsigned long long*, int); }
// out: pairs (file id << 32 | source line, %globaltimer ns), unsorted; returns the number of marks and clears the buffers
extern "C" int b2lo_debug_timeline(unsigned long long* out, int cap) {
  cudaDeviceSynchronize();
  int n = 0;
  n += b2::tl_fetch_filter(out + 2 * n, cap - n);
  n += b2::tl_fetch_icp(out + 2 * n, cap - n);
  n += b2::tl_fetch_odom(out + 2 * n, cap - n);
  n += b2::tl_fetch_map(out + 2 * n, cap - n);
  return n;
}
#endif
