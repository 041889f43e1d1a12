// Debug aid (GPU box): ceiling of random 32 B-sector gathers on this GPU, the access pattern of the K2 surfel probe on a hash table
// larger than L2.  Each thread issues U independent 32 B loads (two float4 halves of one sector) at hashed positions of a 128 MiB
// table, then folds them into a checksum.  Prints sectors/s and the equivalent "48 B per query" algorithmic GB/s for comparison
// with large_map_stress.k2_probe in bench.py.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_probe tools/gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ tab, uint32_t mask, int n, float* out, int coherent) {
  float acc = 0.f;
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t i = (uint32_t)(i0 + u * stride);
      uint32_t h = coherent ? (i / 9u) * 2654435761u : i * 2654435761u;   // coherent: 9 consecutive queries share a sector
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const float4* e = tab + 2 * (size_t)(h & mask);
      a[u] = __ldg(e); b[u] = __ldg(e + 1);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += a[u].x + a[u].w + b[u].y + b[u].z;
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const size_t sectors = 1u << 22;   // 4 Mi sectors x 32 B = 128 MiB
  float4* tab; float* out; char* flush;
  cudaMalloc(&tab, sectors * 32); cudaMemset(tab, 0, sectors * 32);
  cudaMalloc(&out, 4);
  cudaMalloc(&flush, 256u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int n = 1 << 20;
  for (int coherent = 0; coherent < 2; ++coherent)
    for (int blocks : {148, 296, 592, 1184, 2368, 4096}) {
      auto run = [&](auto tag, const char* name) {
        constexpr int U = decltype(tag)::value;
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          cudaMemset(flush, rep, 256u << 20);
          cudaEventRecord(e0);
          k_gather<U><<<blocks, 256>>>(tab, (uint32_t)(sectors - 1), n, out, coherent);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("%s blocks=%4d U=%s: %7.2f us  %6.2f Gsector/s  32B-only %7.1f GB/s  (as 48 B/query: %7.1f GB/s)\n", coherent ? "coherent" : "random  ", blocks, name,
               best * 1e3, n / (best * 1e-3) / 1e9, 32.0 * n / (best * 1e-3) / 1e9, 48.0 * n / (best * 1e-3) / 1e9);
      };
      run(std::integral_constant<int, 1>{}, "1");
      run(std::integral_constant<int, 4>{}, "4");
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
