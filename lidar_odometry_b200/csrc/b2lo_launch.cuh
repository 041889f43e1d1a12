// b2lo_launch.cuh — one launch path for a lone sequence and for lock-step batches of sequences.
//
// A kernel of the per-scan path is written once, as a struct with a static __device__ run(args...).  launch<K>(ctx, grid, block, ...)
// either starts it for ONE sequence (k_one: run(args...), grid.y = 1) or, while the context is RECORDING, files the launch - kernel,
// geometry and the argument values as plain bytes - instead of starting it.  The lock-step driver (b2lo_odom.cu) records the launch
// sequence of one scan for each of S independent sequences with the unchanged host code, zips the S lists and starts every step as ONE
// kernel, k_many, whose blockIdx.y picks the sequence: its argument pack is read from a device array.  The single-CTA latency chains of a
// scan (PKO fit, Gauss-Newton finish, update close) then run for S sequences at once on S SMs, and the GPU's front end dispatches 29
// kernels per STEP instead of 29 per scan (the stream-per-sequence throughput mode is capped by that dispatch rate, ~1.2 M kernels/s).
// Kernel bodies must not read blockIdx.y / gridDim.y: a sequence sees a 1-D grid either way.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include <vector>

struct b2lo_ctx;

namespace b2 {

// plain-old-data argument pack (std::tuple is not guaranteed to be trivially copyable)
template <class... A> struct Pack;
template <> struct Pack<> {};
template <class H, class... T> struct Pack<H, T...> { H h; Pack<T...> t; };
inline void pack_fill(Pack<>&) {}
template <class H, class... T> inline void pack_fill(Pack<H, T...>& p, H h, T... t) { p.h = h; pack_fill(p.t, t...); }

template <class K, class... U> __device__ __forceinline__ void pack_call(const Pack<>&, U... u) { K::run(u...); }
template <class K, class H, class... T, class... U> __device__ __forceinline__ void pack_call(const Pack<H, T...>& p, U... u) { pack_call<K>(p.t, u..., p.h); }

template <class K, int MAXT, int MINB, class... A> __global__ void __launch_bounds__(MAXT, MINB) k_one(A... a) { K::run(a...); }
template <class K, int MAXT, int MINB, class... A> __global__ void __launch_bounds__(MAXT, MINB) k_many(const Pack<A...>* __restrict__ packs) {
  const Pack<A...> p = packs[blockIdx.y];   // this sequence's arguments (one small global read ahead of the body)
  pack_call<K>(p);
}

struct LaunchRec {
  void (*many)(const void* packs, int S, dim3 grid, dim3 block, size_t smem, cudaStream_t st) = nullptr;   // also the identity of the kernel
  void (*prep)(size_t smem) = nullptr;   // function attributes of the batched instantiation, to be set OUTSIDE a stream capture
  dim3 grid, block;
  size_t smem = 0;
  std::vector<unsigned char> args;   // the Pack<A...> of this sequence
};
struct Recorder { std::vector<LaunchRec> recs; };
Recorder* ctx_recorder(b2lo_ctx* ctx);   // b2lo_core.cu: the context's recorder while it is recording, else nullptr

template <class K, int MAXT, int MINB, class... A>
void launch_many_prep(size_t smem) {
  static size_t have = 48 * 1024;
  if (smem > have) { cudaFuncSetAttribute(k_many<K, MAXT, MINB, A...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); have = smem; }
}
template <class K, int MAXT, int MINB, class... A>
void launch_many(const void* packs, int S, dim3 g, dim3 b, size_t smem, cudaStream_t st) {
  k_many<K, MAXT, MINB, A...><<<dim3(g.x, (unsigned)S), b, smem, st>>>(static_cast<const Pack<A...>*>(packs));
}

// MINBM: min CTAs per SM of the BATCHED instantiation.  A lone sequence's gather kernels are latency chains of a few CTAs and take
// the registers they like; the same body running for 128 sequences is bound by the number of warps in flight, so its register budget
// is capped separately (a few spilled pointers cost less than half the occupancy).
template <class K, int MAXT = 1024, int MINB = 1, int MINBM = MINB, class... A>
inline void launch(b2lo_ctx* ctx, dim3 g, dim3 b, size_t smem, cudaStream_t st, A... a) {
  if (Recorder* r = ctx_recorder(ctx)) {
    LaunchRec rec;
    rec.many = &launch_many<K, MAXT, MINBM, A...>;
    rec.prep = &launch_many_prep<K, MAXT, MINBM, A...>;
    rec.grid = g; rec.block = b; rec.smem = smem;
    Pack<A...> p;
    std::memset(&p, 0, sizeof p);      // padding bytes compare equal between recordings
    pack_fill(p, a...);
    rec.args.resize(sizeof p);
    std::memcpy(rec.args.data(), &p, sizeof p);
    r->recs.push_back(std::move(rec));
    return;
  }
  static bool attr = false;
  if (smem > 48 * 1024 && !attr) { cudaFuncSetAttribute(k_one<K, MAXT, MINB, A...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  k_one<K, MAXT, MINB, A...><<<g, b, smem, st>>>(a...);
}

}  // namespace b2
