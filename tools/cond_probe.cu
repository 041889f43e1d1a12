#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* ctr, cudaGraphConditionalHandle h) {
  int v = atomicAdd(ctr, 1) + 1;
  cudaGraphSetConditional(h, v < 5 ? 1u : 0u);
}
__global__ void pre(int* ctr) { *ctr = 0; }
int main() {
  int* d; cudaMalloc(&d, 4);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaGraph_t g; 
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  pre<<<1,1,0,s>>>(d);
  // add a WHILE node behind the captured work
  cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t cg; const cudaGraphNode_t* deps; size_t nd;
  cudaError_t e = cudaStreamGetCaptureInfo_v2(s, &st, &id, &cg, &deps, &nd);
  printf("info %s nd=%zu\n", cudaGetErrorString(e), nd);
  cudaGraphConditionalHandle h;
  e = cudaGraphConditionalHandleCreate(&h, cg, 1, cudaGraphCondAssignDefault);
  printf("handle %s\n", cudaGetErrorString(e));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  e = cudaGraphAddNode(&node, cg, deps, nd, &p);
  printf("addnode %s\n", cudaGetErrorString(e));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  cudaStream_t s2; cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  e = cudaStreamBeginCaptureToGraph(s2, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  printf("begin body %s\n", cudaGetErrorString(e));
  body<<<1,1,0,s2>>>(d, h);
  body<<<1,1,0,s2>>>(d, h);
  cudaGraph_t tmp; e = cudaStreamEndCapture(s2, &tmp);
  printf("end body %s\n", cudaGetErrorString(e));
  e = cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
  printf("update deps %s\n", cudaGetErrorString(e));
  e = cudaStreamEndCapture(s, &g);
  printf("end %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0);
  printf("inst %s\n", cudaGetErrorString(e));
  for (int r = 0; r < 3; ++r) { cudaGraphLaunch(ex, s); cudaStreamSynchronize(s); int h_; cudaMemcpy(&h_, d, 4, cudaMemcpyDeviceToHost); printf("ctr=%d\n", h_); }
  // timing
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, s); for (int r = 0; r < 200; ++r) cudaGraphLaunch(ex, s); cudaEventRecord(b, s); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); printf("per launch %.2f us (3 loop trips x 2 kernels + pre)\n", ms * 1000 / 200);
  printf("last %s\n", cudaGetErrorString(cudaGetLastError()));
}
