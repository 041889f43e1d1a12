// b2lo_math.cuh — small fixed-size numerics of the registration engine, host + device.
//
// Everything here is IEEE f32/f64 WITHOUT fused multiply-add (the library is compiled with
// -fmad=false / -ffp-contract=off) and in the association order the reference's Eigen expressions
// evaluate to, because voxel keys, planarity gates and the PKO arg-min are discrete functions of
// these bits.  Reference call sites:
//   SO3(Matrix3f) re-projection, Exp, Log, SE3 compose/inverse   src/util/MathUtils.cpp:23-99, MathUtils.h:78-158
//   JacobiSVD<Matrix3f> for surfels                              src/database/VoxelMap.cpp:239,348
//   Matrix<float,6,6>::ldlt().solve                              src/optimization/IterativeClosestPointOptimizer.cpp:418
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2 {

// size-3 reductions: floats reduce as a + (b + c), doubles as (a + b) + c (packet of two, then the tail)
B2_HD float add3(float a, float b, float c) { return a + (b + c); }
B2_HD double add3(double a, double b, double c) { return (a + b) + c; }
B2_HD float dot3(const float* a, const float* b) { return add3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
B2_HD float sqn3(const float* a) { return add3(a[0] * a[0], a[1] * a[1], a[2] * a[2]); }

struct Mat3 { float m[9]; };  // row-major

B2_HD Mat3 mat3_identity() { Mat3 r; for (int i = 0; i < 9; ++i) r.m[i] = (i % 4 == 0) ? 1.0f : 0.0f; return r; }
B2_HD void mat3_vec(const float* M, const float* v, float* o) {
  float a = add3(M[0] * v[0], M[1] * v[1], M[2] * v[2]);
  float b = add3(M[3] * v[0], M[4] * v[1], M[5] * v[2]);
  float c = add3(M[6] * v[0], M[7] * v[1], M[8] * v[2]);
  o[0] = a; o[1] = b; o[2] = c;
}
B2_HD Mat3 mat3_mul(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i * 3 + j] = add3(A.m[i * 3] * B.m[j], A.m[i * 3 + 1] * B.m[3 + j], A.m[i * 3 + 2] * B.m[6 + j]);
  return C;
}
B2_HD Mat3 mat3_t(const Mat3& A) {
  Mat3 T;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.m[j * 3 + i] = A.m[i * 3 + j];
  return T;
}
B2_HD float mat3_det(const Mat3& A) {
  const float* m = A.m;
  float h0 = m[0] * (m[4] * m[8] - m[5] * m[7]);
  float h1 = m[1] * (m[3] * m[8] - m[5] * m[6]);
  float h2 = m[2] * (m[3] * m[7] - m[4] * m[6]);
  return h0 - h1 + h2;
}

// ---- two-sided Jacobi SVD of a 3x3 f32 matrix (Eigen 3.4 JacobiSVD semantics) ----------------------
// All indices are compile-time constants after unrolling, so W/U/V stay in registers on the device
// (the Gauss-Newton finish and the surfel refit run this on a single thread: latency matters).
struct Rot2 { float c, s; };
B2_HD Rot2 sym_jacobi(float x, float y, float z) {
  Rot2 j;
  float deno = 2.0f * fabsf(y);
  if (deno < FLT_MIN) { j.c = 1.0f; j.s = 0.0f; return j; }
  float tau = (x - z) / deno;
  float w = sqrtf(tau * tau + 1.0f);
  float t = (tau > 0.0f) ? 1.0f / (tau + w) : 1.0f / (tau - w);
  float sign_t = t > 0.0f ? 1.0f : -1.0f;
  float n = 1.0f / sqrtf(t * t + 1.0f);
  j.s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
  j.c = n;
  return j;
}
// real_2x2_jacobi_svd on the (p,q) sub-block [[a b],[c d]] = [[Wpp Wpq],[Wqp Wqq]]
B2_HD void svd2x2(float a, float b, float c, float d, Rot2& jl, Rot2& jr) {
  Rot2 r1;
  float t = a + d, dd = c - b;
  if (fabsf(dd) < FLT_MIN) { r1.s = 0.0f; r1.c = 1.0f; }
  else { float u = t / dd; float tmp = sqrtf(1.0f + u * u); r1.s = 1.0f / tmp; r1.c = u / tmp; }
  if (!(r1.c == 1.0f && r1.s == 0.0f)) {
    float na = r1.c * a + r1.s * c, nb = r1.c * b + r1.s * d;
    float nd = r1.c * d - r1.s * b;
    a = na; b = nb; d = nd;
  }
  jr = sym_jacobi(a, b, d);
  Rot2 jrt{jr.c, -jr.s};
  jl.c = r1.c * jrt.c - r1.s * jrt.s;
  jl.s = r1.c * jrt.s + r1.s * jrt.c;
}
// apply_rotation_in_the_plane on (x, y): x' = c x + s y ; y' = c y - s x
#define B2_ROT(X, Y, J)                                   \
  do {                                                    \
    float _x = (X), _y = (Y);                             \
    (X) = (J).c * _x + (J).s * _y;                        \
    (Y) = (J).c * _y - (J).s * _x;                        \
  } while (0)
// one (p,q) step of the sweep; P, Q are literal indices
#define B2_SVD_PAIR(P, Q)                                                                           \
  do {                                                                                              \
    float thr = fmaxf(FLT_MIN, precision * maxd);                                                   \
    if (fabsf(W[P * 3 + Q]) > thr || fabsf(W[Q * 3 + P]) > thr) {                                   \
      done = false;                                                                                 \
      Rot2 jl, jr;                                                                                  \
      svd2x2(W[P * 3 + P], W[P * 3 + Q], W[Q * 3 + P], W[Q * 3 + Q], jl, jr);                       \
      if (!(jl.c == 1.0f && jl.s == 0.0f)) {                                                        \
        B2_ROT(W[P * 3 + 0], W[Q * 3 + 0], jl); B2_ROT(W[P * 3 + 1], W[Q * 3 + 1], jl); B2_ROT(W[P * 3 + 2], W[Q * 3 + 2], jl); \
        B2_ROT(Um[0 * 3 + P], Um[0 * 3 + Q], jl); B2_ROT(Um[1 * 3 + P], Um[1 * 3 + Q], jl); B2_ROT(Um[2 * 3 + P], Um[2 * 3 + Q], jl); \
      }                                                                                             \
      Rot2 jrt{jr.c, -jr.s};                                                                        \
      if (!(jrt.c == 1.0f && jrt.s == 0.0f)) {                                                      \
        B2_ROT(W[0 * 3 + P], W[0 * 3 + Q], jrt); B2_ROT(W[1 * 3 + P], W[1 * 3 + Q], jrt); B2_ROT(W[2 * 3 + P], W[2 * 3 + Q], jrt); \
        B2_ROT(Vm[0 * 3 + P], Vm[0 * 3 + Q], jrt); B2_ROT(Vm[1 * 3 + P], Vm[1 * 3 + Q], jrt); B2_ROT(Vm[2 * 3 + P], Vm[2 * 3 + Q], jrt); \
      }                                                                                             \
      maxd = fmaxf(maxd, fmaxf(fabsf(W[P * 3 + P]), fabsf(W[Q * 3 + Q])));                          \
    }                                                                                               \
  } while (0)
#define B2_SWAP_COL(A, I, J)                                                                        \
  do {                                                                                              \
    float _t;                                                                                       \
    _t = A[0 * 3 + I]; A[0 * 3 + I] = A[0 * 3 + J]; A[0 * 3 + J] = _t;                              \
    _t = A[1 * 3 + I]; A[1 * 3 + I] = A[1 * 3 + J]; A[1 * 3 + J] = _t;                              \
    _t = A[2 * 3 + I]; A[2 * 3 + I] = A[2 * 3 + J]; A[2 * 3 + J] = _t;                              \
  } while (0)
// A = U diag(S) V^T, S descending.
B2_HD void svd3(const Mat3& A, Mat3& U, float* S, Mat3& V) {
  const float precision = 2.0f * FLT_EPSILON;
  float scale = 0.0f;
  bool bad = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 9; ++i) { float a = fabsf(A.m[i]); if (!(a == a) || a > FLT_MAX) bad = true; if (a > scale) scale = a; }
  float Um[9] = {1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
  float Vm[9] = {1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
  if (!bad) {
    if (scale == 0.0f) scale = 1.0f;
    float W[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 9; ++i) W[i] = A.m[i] / scale;
    float maxd = fmaxf(fabsf(W[0]), fmaxf(fabsf(W[4]), fabsf(W[8])));
    bool done = false;
    while (!done) {
      done = true;
      B2_SVD_PAIR(1, 0);
      B2_SVD_PAIR(2, 0);
      B2_SVD_PAIR(2, 1);
    }
    s0 = fabsf(W[0]); s1 = fabsf(W[4]); s2 = fabsf(W[8]);
    if (W[0] < 0.0f) { Um[0] = -Um[0]; Um[3] = -Um[3]; Um[6] = -Um[6]; }
    if (W[4] < 0.0f) { Um[1] = -Um[1]; Um[4] = -Um[4]; Um[7] = -Um[7]; }
    if (W[8] < 0.0f) { Um[2] = -Um[2]; Um[5] = -Um[5]; Um[8] = -Um[8]; }
    s0 *= scale; s1 *= scale; s2 *= scale;
    // selection sort, descending, first maximum wins ties, stop at an all-zero tail (Eigen)
    bool stop = false;
    {
      int pos = 0; float mx = s0;
      if (s1 > mx) { mx = s1; pos = 1; }
      if (s2 > mx) { mx = s2; pos = 2; }
      if (mx == 0.0f) stop = true;
      else if (pos == 1) { float t = s0; s0 = s1; s1 = t; B2_SWAP_COL(Um, 0, 1); B2_SWAP_COL(Vm, 0, 1); }
      else if (pos == 2) { float t = s0; s0 = s2; s2 = t; B2_SWAP_COL(Um, 0, 2); B2_SWAP_COL(Vm, 0, 2); }
    }
    if (!stop) {
      if (s2 > s1) { float t = s1; s1 = s2; s2 = t; B2_SWAP_COL(Um, 1, 2); B2_SWAP_COL(Vm, 1, 2); }
    }
  }
  S[0] = s0; S[1] = s1; S[2] = s2;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 9; ++i) { U.m[i] = Um[i]; V.m[i] = Vm[i]; }
}

// nearest rotation: SO3::SO3(const Matrix3f&)
B2_HD Mat3 so3_project(const Mat3& R) {
  Mat3 U, V; float S[3];
  svd3(R, U, S, V);
  Mat3 Vt = mat3_t(V);
  Mat3 M = mat3_mul(U, Vt);
  if (mat3_det(M) < 0.0f) {
    U.m[2] *= -1.0f; U.m[5] *= -1.0f; U.m[8] *= -1.0f;
    M = mat3_mul(U, Vt);
  }
  return M;
}
// Nearest rotation of a matrix that is ALREADY orthonormal up to f32 rounding (a product / exponential of rotations):
// one Newton-Schulz step X <- X (3I - X^T X) / 2 takes an orthogonality error e to 1.5 e^2, i.e. from ~1e-7 to below f32
// resolution, so the result equals the SVD projection U V^T up to the last bit or two - at a fraction of the ~8k-cycle
// dependent chain of the Jacobi sweeps.  Used only on the single-thread finish of the device Gauss-Newton iteration;
// inputs that are not near-orthonormal (|X^T X - I| > 1e-3) fall back to the exact SVD path.
B2_HD Mat3 so3_project_near(const Mat3& X) {
  Mat3 Xt = mat3_t(X);
  Mat3 G = mat3_mul(Xt, X);
  float dev = 0.0f;
  for (int i = 0; i < 9; ++i) dev = fmaxf(dev, fabsf(G.m[i] - ((i % 4 == 0) ? 1.0f : 0.0f)));
  if (!(dev < 1e-3f)) return so3_project(X);
  Mat3 H;
  for (int i = 0; i < 9; ++i) H.m[i] = 0.5f * (((i % 4 == 0) ? 3.0f : 0.0f) - G.m[i]);
  return mat3_mul(X, H);
}
B2_HD Mat3 hat(const float* w) {
  Mat3 K;
  K.m[0] = 0.0f; K.m[1] = -w[2]; K.m[2] = w[1];
  K.m[3] = w[2]; K.m[4] = 0.0f; K.m[5] = -w[0];
  K.m[6] = -w[1]; K.m[7] = w[0]; K.m[8] = 0.0f;
  return K;
}
// sinf / cosf as glibc computes them (correctly rounded in practice): on the device the f64 value rounded to f32.
// Gauss-Newton steps are small angles, so |x| < 0.5 takes a Taylor series in f64 (error < 2e-17, then one rounding)
// instead of libdevice's full-range routines, whose ~40-deep f64 chains sit on the single-thread finish path.
B2_HD float sin_f32(float x) {
#if defined(__CUDA_ARCH__)
  const double xd = (double)x;
  if (fabs(xd) < 0.5) {
    const double z = xd * xd;
    double p = -1.0 / 1307674368000.0;            // x^15
    p = p * z + 1.0 / 6227020800.0;               // x^13
    p = p * z - 1.0 / 39916800.0;
    p = p * z + 1.0 / 362880.0;
    p = p * z - 1.0 / 5040.0;
    p = p * z + 1.0 / 120.0;
    p = p * z - 1.0 / 6.0;
    return (float)(xd + xd * (z * p));
  }
  return (float)sin(xd);
#else
  return sinf(x);
#endif
}
B2_HD float cos_f32(float x) {
#if defined(__CUDA_ARCH__)
  const double xd = (double)x;
  if (fabs(xd) < 0.5) {
    const double z = xd * xd;
    double p = 1.0 / 20922789888000.0;            // x^16
    p = p * z - 1.0 / 87178291200.0;              // x^14
    p = p * z + 1.0 / 479001600.0;
    p = p * z - 1.0 / 3628800.0;
    p = p * z + 1.0 / 40320.0;
    p = p * z - 1.0 / 720.0;
    p = p * z + 1.0 / 24.0;
    p = p * z - 0.5;
    return (float)(1.0 + z * p);
  }
  return (float)cos(xd);
#else
  return cosf(x);
#endif
}
B2_HD Mat3 so3_exp(const float* w) {
  float theta = sqrtf(sqn3(w));
  Mat3 I = mat3_identity(), M;
  if (theta < 1e-6f) {
    Mat3 K = hat(w);
    for (int i = 0; i < 9; ++i) M.m[i] = I.m[i] + K.m[i];
    return so3_project(M);
  }
  float ti = 1.0f / theta;
  float k[3] = {w[0] * ti, w[1] * ti, w[2] * ti};
  Mat3 K = hat(k);
  float s = sin_f32(theta), omc = 1.0f - cos_f32(theta);
  Mat3 oK;
  for (int i = 0; i < 9; ++i) oK.m[i] = omc * K.m[i];
  Mat3 KK = mat3_mul(oK, K);
  for (int i = 0; i < 9; ++i) M.m[i] = (I.m[i] + s * K.m[i]) + KK.m[i];
  return so3_project(M);
}
B2_HD void so3_log(const Mat3& R, float* out) {
  const float* m = R.m;
  float trace = add3(m[0], m[4], m[8]);
  float ct = (trace - 1.0f) * 0.5f;
  float cc = fmaxf(-1.0f, fminf(1.0f, ct));
  float theta = acosf(cc);
  if (theta < 1e-6f) { out[0] = m[7]; out[1] = m[2]; out[2] = m[3]; return; }
  float st = sin_f32(theta);
  if (fabsf(st) < 1e-6f) {
    float ax[3] = {0.0f, 0.0f, 0.0f};
    int mi = 0;
    if (m[4] > m[0]) mi = 1;
    if (m[8] > m[mi * 4]) mi = 2;
    ax[mi] = sqrtf((m[mi * 4] + 1.0f) * 0.5f);
    for (int i = 0; i < 3; ++i) if (i != mi) ax[i] = m[mi * 3 + i] / (2.0f * ax[mi]);
    float sk[3] = {(m[7] - m[5]) * 0.5f, (m[2] - m[6]) * 0.5f, (m[3] - m[1]) * 0.5f};
    if (dot3(ax, sk) < 0.0f) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; }
    out[0] = ax[0] * theta; out[1] = ax[1] * theta; out[2] = ax[2] * theta;
    return;
  }
  float f = theta / (2.0f * st);
  out[0] = f * (m[7] - m[5]); out[1] = f * (m[2] - m[6]); out[2] = f * (m[3] - m[1]);
}

struct Pose { Mat3 R; float t[3]; };  // util::SE3 (float)

B2_HD Pose pose_from_T16(const float* T) {
  Pose p;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) p.R.m[i * 3 + j] = T[i * 4 + j]; p.t[i] = T[i * 4 + 3]; }
  return p;
}
B2_HD void pose_to_T16(const Pose& p, float* T) {
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = p.R.m[i * 3 + j]; T[i * 4 + 3] = p.t[i]; }
  T[12] = 0.0f; T[13] = 0.0f; T[14] = 0.0f; T[15] = 1.0f;
}
// SE3::operator* : rotation product re-projected, t = t_a + R_a t_b
B2_HD Pose pose_mul(const Pose& a, const Pose& b) {
  Pose r;
  float rt[3];
  mat3_vec(a.R.m, b.t, rt);
  r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
  r.R = so3_project(mat3_mul(a.R, b.R));
  return r;
}
B2_HD Pose pose_inv(const Pose& a) {
  Pose r;
  r.R = so3_project(mat3_t(a.R));
  float mt[3] = {-a.t[0], -a.t[1], -a.t[2]};
  mat3_vec(r.R.m, mt, r.t);
  return r;
}

// ---- pivoted LDL^T (lower) solve of a 6x6 f32 system: x = H.ldlt().solve(b) -------------------------
// Eigen's unblocked LDLT with largest-|diagonal| pivoting.  Every index below is a compile-time constant once the
// loops are unrolled (the dynamic pivot only selects between statically indexed swap blocks), so the matrix lives in
// registers on the device: this runs on ONE thread at the end of every Gauss-Newton iteration.
#if defined(__CUDA_ARCH__)
#define B2_UNROLL _Pragma("unroll")
#else
#define B2_UNROLL
#endif
#define B2_FSWAP(a, b) do { float _t = (a); (a) = (b); (b) = _t; } while (0)
B2_HD void ldlt6_solve(const float* Hin, const float* b, float* x) {
  float A[36];
  B2_UNROLL
  for (int i = 0; i < 36; ++i) A[i] = Hin[i];
  int perm[6] = {0, 1, 2, 3, 4, 5};
  float tmp[6];
  bool all_zero = false;
  B2_UNROLL
  for (int k = 0; k < 6; ++k) {
    if (all_zero) continue;
    int piv = k; float best = fabsf(A[k * 7]);
    B2_UNROLL
    for (int i = k + 1; i < 6; ++i) { float a = fabsf(A[i * 7]); if (a > best) { best = a; piv = i; } }
    perm[k] = piv;
    B2_UNROLL
    for (int p = k + 1; p < 6; ++p) {
      if (piv == p) {
        B2_UNROLL
        for (int c = 0; c < k; ++c) B2_FSWAP(A[k * 6 + c], A[p * 6 + c]);
        B2_UNROLL
        for (int r = p + 1; r < 6; ++r) B2_FSWAP(A[r * 6 + k], A[r * 6 + p]);
        B2_FSWAP(A[k * 7], A[p * 7]);
        B2_UNROLL
        for (int i = k + 1; i < p; ++i) B2_FSWAP(A[i * 6 + k], A[p * 6 + i]);
      }
    }
    if (k > 0) {
      B2_UNROLL
      for (int c = 0; c < k; ++c) tmp[c] = A[c * 7] * A[k * 6 + c];
      float acc = A[k * 6] * tmp[0];
      B2_UNROLL
      for (int c = 1; c < k; ++c) acc = acc + A[k * 6 + c] * tmp[c];
      A[k * 7] -= acc;
      B2_UNROLL
      for (int r = k + 1; r < 6; ++r) {
        float a2 = A[r * 6] * tmp[0];
        B2_UNROLL
        for (int c = 1; c < k; ++c) a2 = a2 + A[r * 6 + c] * tmp[c];
        A[r * 6 + k] -= a2;
      }
    }
    const float akk = A[k * 7];
    const bool valid = fabsf(akk) > 0.0f;
    if (k == 0 && !valid) { all_zero = true; B2_UNROLL for (int j = 0; j < 6; ++j) perm[j] = j; continue; }
    if (valid) {
      B2_UNROLL
      for (int r = k + 1; r < 6; ++r) A[r * 6 + k] /= akk;
    }
  }
  float y[6];
  B2_UNROLL
  for (int i = 0; i < 6; ++i) y[i] = b[i];
  B2_UNROLL
  for (int k = 0; k < 6; ++k) {
    B2_UNROLL
    for (int p = k + 1; p < 6; ++p) if (perm[k] == p) B2_FSWAP(y[k], y[p]);
  }
  B2_UNROLL
  for (int i = 0; i < 6; ++i) {
    float acc = y[i];
    B2_UNROLL
    for (int c = 0; c < i; ++c) acc -= A[i * 6 + c] * y[c];
    y[i] = acc;
  }
  B2_UNROLL
  for (int i = 0; i < 6; ++i) { if (fabsf(A[i * 7]) > FLT_MIN) y[i] /= A[i * 7]; else y[i] = 0.0f; }
  B2_UNROLL
  for (int i = 5; i >= 0; --i) {
    float acc = y[i];
    B2_UNROLL
    for (int c = i + 1; c < 6; ++c) acc -= A[c * 6 + i] * y[c];
    y[i] = acc;
  }
  B2_UNROLL
  for (int k = 5; k >= 0; --k) {
    B2_UNROLL
    for (int p = k + 1; p < 6; ++p) if (perm[k] == p) B2_FSWAP(y[k], y[p]);
  }
  B2_UNROLL
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// PCA plane of up to 27 centroids: mean, covariance / N, SVD -> normal = U[:,2], planarity = s2/(s0+1e-6)
B2_HD void fit_plane(const float* pts /*n x 3*/, int n, float* mu, float* normal, float* planarity) {
  float c[3] = {0.0f, 0.0f, 0.0f};
  for (int i = 0; i < n; ++i) { c[0] += pts[i * 3]; c[1] += pts[i * 3 + 1]; c[2] += pts[i * 3 + 2]; }
  float nf = (float)n;
  c[0] /= nf; c[1] /= nf; c[2] /= nf;
  Mat3 cov;
  for (int i = 0; i < 9; ++i) cov.m[i] = 0.0f;
  for (int i = 0; i < n; ++i) {
    float d[3] = {pts[i * 3] - c[0], pts[i * 3 + 1] - c[1], pts[i * 3 + 2] - c[2]};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov.m[a * 3 + b] += d[a] * d[b];
  }
  for (int i = 0; i < 9; ++i) cov.m[i] /= nf;
  Mat3 U, V; float S[3];
  svd3(cov, U, S, V);
  normal[0] = U.m[2]; normal[1] = U.m[5]; normal[2] = U.m[8];
  *planarity = S[2] / (S[0] + 1e-6f);
  mu[0] = c[0]; mu[1] = c[1]; mu[2] = c[2];
}

}  // namespace b2
