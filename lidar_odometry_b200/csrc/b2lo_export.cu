// b2lo_export.cu — final-map export: util::VoxelGrid on the device.
//
// Replaces util::VoxelGrid::filter (/root/reference/src/util/PointCloudUtils.h:462-557) as Estimator::save_map_to_ply uses it
// (/root/reference/src/processing/Estimator.cpp:1248-1305): the accumulated keyframe clouds (millions of world points at the end of a
// sequence) are reduced to one running-average centroid per leaf-sized voxel and emitted in std::map<VoxelKey> order (x, then y, then z).
// Device algorithm: the K1 machinery (b2lo_filter.cu) in mode 1 - hash-insert on the packed grid key, per-voxel segments, every point
// ranks itself inside its segment, one thread per voxel replays WeightedCentroid::add_point over its points in input order - followed
// by a radix sort of the (at most n) voxel keys (cub::DeviceRadixSort, a plain library sort outside the per-scan path) and a gather.
// Algorithmic bytes: 12 B read per input point + 12 B written per voxel.
// Deviation from the reference, documented: points with a non-finite coordinate or beyond +-2^20 leaves are dropped (the reference's
// static_cast<int>(floor(NaN)) is undefined behaviour; clouds reaching this call went through K1, which already drops non-finite points).
#include <cub/device/device_radix_sort.cuh>
#include <cstring>
#include "b2lo_internal.h"

using namespace b2;

namespace {
__global__ void k_iota(int* v, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}
__global__ void k_gather_xyz(const float4* __restrict__ src, const int* __restrict__ order, int n, float* __restrict__ out3) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 c = src[order[i]];
    out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
  }
}
struct Scratch {   // freed on every exit path
  unsigned long long* keys_out = nullptr; int* vals_in = nullptr; int* vals_out = nullptr; void* temp = nullptr; float* out3 = nullptr;
  ~Scratch() { cudaFree(keys_out); cudaFree(vals_in); cudaFree(vals_out); cudaFree(temp); cudaFree(out3); }
};
}  // namespace

extern "C" int b2lo_voxel_grid_filter(b2lo_ctx* ctx, const float* xyz, size_t n, size_t stride_floats, float leaf_size, float* out_xyz, size_t cap,
                                      size_t* m) {
  if (!ctx || !m) return B2LO_E_ARG;
  *m = 0;
  if (stride_floats < 3) { set_error("voxel grid: bad stride"); return B2LO_E_ARG; }
  if (!xyz || n == 0 || !(leaf_size > 0.0f)) return B2LO_S_EMPTY;   // output.clear() (PointCloudUtils.h:471-474)
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  int rc = ctx_stage_h2d(ctx, xyz, n, stride_floats, 1, nullptr, nullptr);
  if (rc) return rc;
  rc = filter_run(ctx, ctx->d_stage, n, 3, leaf_size, 0, nullptr, nullptr, 1);
  if (rc) return rc;
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 18, ctx->d_nfeat, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const size_t M = (size_t)ctx->h_counts[18];
  ctx->d2h_bytes += sizeof(int);
  ctx->feat_cap_hint = 0; ctx->feat_cap_hint_set[0] = 0; ctx->feat_set = 0;   // the feature buffer now holds map voxels, not scan features
  if (M == 0) return B2LO_OK;
  if (M > cap || !out_xyz) { *m = M; set_error("voxel grid: output buffer too small (%zu < %zu)", cap, M); return B2LO_E_CAPACITY; }
  Scratch s;
  size_t temp_bytes = 0;
  B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, ctx->d_feat_key, s.keys_out, s.vals_in, s.vals_out, (int)M, 0, 63, st));
  B2_CUDA(cudaMalloc((void**)&s.keys_out, M * sizeof(unsigned long long)));
  B2_CUDA(cudaMalloc((void**)&s.vals_in, M * sizeof(int)));
  B2_CUDA(cudaMalloc((void**)&s.vals_out, M * sizeof(int)));
  B2_CUDA(cudaMalloc(&s.temp, temp_bytes ? temp_bytes : 16));
  B2_CUDA(cudaMalloc((void**)&s.out3, M * 3 * sizeof(float)));
  int blocks = (int)((M + 255) / 256); if (blocks > 2368) blocks = 2368;
  k_iota<<<blocks, 256, 0, st>>>(s.vals_in, (int)M);
  B2_CUDA(cub::DeviceRadixSort::SortPairs(s.temp, temp_bytes, ctx->d_feat_key, s.keys_out, s.vals_in, s.vals_out, (int)M, 0, 63, st));
  k_gather_xyz<<<blocks, 256, 0, st>>>(ctx->d_feat, s.vals_out, (int)M, s.out3);
  ctx->launches += 2;
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaMemcpyAsync(out_xyz, s.out3, M * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  ctx->d2h_bytes += M * 3 * sizeof(float);
  *m = M;
  return B2LO_OK;
}
