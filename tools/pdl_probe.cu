// Probe: cost of a kernel boundary inside a CUDA graph on this GPU, with and without programmatic dependent launch.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pdl_probe tools/pdl_probe.cu && /tmp/pdl_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_small(int* p, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
  if (threadIdx.x == 0) p[blockIdx.x] += 1;
}
__global__ void k_work(int* p, int pdl, int spin) {   // ~single-CTA latency-bound kernel of a few us
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (threadIdx.x == 0) p[blockIdx.x] += 1;
}
static void launch(cudaStream_t s, int* d, int grid, int pdl, int spin) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  if (spin) cudaLaunchKernelEx(&cfg, k_work, d, pdl, spin); else cudaLaunchKernelEx(&cfg, k_small, d, pdl);
}
int main() {
  int* d; cudaMalloc(&d, 4096 * 4); cudaMemset(d, 0, 4096 * 4);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int spin : {0, 4000, 20000})
  for (int grid : {1, 148})
  for (int pdl = 0; pdl < 2; ++pdl) {
    const int NK = 36;
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < NK; ++i) launch(s, d, grid, pdl, spin);
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return 1; }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return 1; }
    for (int w = 0; w < 5; ++w) cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    const int R = 50;
    for (int r = 0; r < R; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("spin=%5d cycles grid=%3d pdl=%d : %.3f us per kernel (graph of %d)\n", spin, grid, pdl, 1e3 * ms / (R * NK), NK);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  int h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("check %d (err %s)\n", h[0], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
