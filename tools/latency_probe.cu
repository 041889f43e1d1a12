// Micro-probe (debug aid): dependent-chain latencies on the target GPU, in SM cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void probe(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sm[64];
  double x = a + threadIdx.x * 1e-9, y = b;
  float xf = (float)a, yf = (float)b;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, 1e-3);                                   // DFMA chain
    if (OP == 1) x = x + y;                                             // DADD chain
    if (OP == 2) xf = fmaf(xf, yf, 1e-3f);                              // FFMA chain
    if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9;         // shuffle (64-bit) + DADD
    if (OP == 4) x = 1.0 / x + 1.5;                                     // f64 division
    if (OP == 5) x = __drcp_rn(x) + 1.5;                                // f64 reciprocal
    if (OP == 6) x = rsqrt(x) + 1.5;                                    // f64 rsqrt
    if (OP == 7) x = exp(-x) + 0.5;                                     // libdevice exp
    if (OP == 8) { sm[threadIdx.x & 63] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 63] + 1e-9; __syncwarp(); }  // smem round trip
    if (OP == 9) { asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x)); x += 1e-9; }                          // named barrier
    if (OP == 10) x = sqrt(x) + 1.5;
    if (OP == 11) xf = 1.0f / xf + 1.5f;
    if (OP == 12) xf = sqrtf(xf) + 1.5f;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + xf;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA", "DADD", "FFMA", "SHFL64+DADD", "f64 div", "f64 rcp", "f64 rsqrt", "f64 exp", "smem rt", "bar.sync", "f64 sqrt", "f32 div", "f32 sqrt"};
  const int n = 2000;
  for (int threads : {32, 128, 320}) {
    printf("threads=%d\n", threads);
#define RUN(OP) { probe<OP><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n); cudaDeviceSynchronize(); probe<OP><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n); cudaDeviceSynchronize(); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("  %-12s %8.1f cycles/op\n", names[OP], (double)c / n); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
