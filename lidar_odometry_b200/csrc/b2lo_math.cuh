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
struct Rot2 { float c, s; };
B2_HD void plane_rot(float* x, int sx, float* y, int sy, int n, Rot2 j) {
  if (j.c == 1.0f && j.s == 0.0f) return;
  for (int i = 0; i < n; ++i) {
    float xi = x[i * sx], yi = y[i * sy];
    x[i * sx] = j.c * xi + j.s * yi;
    y[i * sy] = j.c * yi - j.s * xi;
  }
}
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
B2_HD void svd2x2(const float* W, int p, int q, Rot2& jl, Rot2& jr) {
  float a = W[p * 3 + p], b = W[p * 3 + q], c = W[q * 3 + p], d = W[q * 3 + q];
  Rot2 r1;
  float t = a + d, dd = c - b;
  if (fabsf(dd) < FLT_MIN) { r1.s = 0.0f; r1.c = 1.0f; }
  else { float u = t / dd; float tmp = sqrtf(1.0f + u * u); r1.s = 1.0f / tmp; r1.c = u / tmp; }
  if (!(r1.c == 1.0f && r1.s == 0.0f)) {
    float na = r1.c * a + r1.s * c, nb = r1.c * b + r1.s * d;
    float nc = r1.c * c - r1.s * a, nd = r1.c * d - r1.s * b;
    a = na; b = nb; c = nc; d = nd;
  }
  (void)c;
  jr = sym_jacobi(a, b, d);
  Rot2 jrt{jr.c, -jr.s};
  jl.c = r1.c * jrt.c - r1.s * jrt.s;
  jl.s = r1.c * jrt.s + r1.s * jrt.c;
}
// A = U diag(S) V^T, S descending.
B2_HD void svd3(const Mat3& A, Mat3& U, float* S, Mat3& V) {
  const float precision = 2.0f * FLT_EPSILON;
  float scale = 0.0f;
  bool bad = false;
  for (int i = 0; i < 9; ++i) { float a = fabsf(A.m[i]); if (!(a == a) || a > FLT_MAX) bad = true; if (a > scale) scale = a; }
  U = mat3_identity(); V = mat3_identity();
  if (bad) { S[0] = S[1] = S[2] = 0.0f; return; }
  if (scale == 0.0f) scale = 1.0f;
  float W[9];
  for (int i = 0; i < 9; ++i) W[i] = A.m[i] / scale;
  float maxd = fmaxf(fabsf(W[0]), fmaxf(fabsf(W[4]), fabsf(W[8])));
  bool done = false;
  while (!done) {
    done = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        float thr = fmaxf(FLT_MIN, precision * maxd);
        if (fabsf(W[p * 3 + q]) > thr || fabsf(W[q * 3 + p]) > thr) {
          done = false;
          Rot2 jl, jr;
          svd2x2(W, p, q, jl, jr);
          plane_rot(&W[p * 3], 1, &W[q * 3], 1, 3, jl);
          plane_rot(&U.m[p], 3, &U.m[q], 3, 3, jl);
          Rot2 jrt{jr.c, -jr.s};
          plane_rot(&W[p], 3, &W[q], 3, 3, jrt);
          plane_rot(&V.m[p], 3, &V.m[q], 3, 3, jrt);
          maxd = fmaxf(maxd, fmaxf(fabsf(W[p * 3 + p]), fabsf(W[q * 3 + q])));
        }
      }
  }
  for (int i = 0; i < 3; ++i) {
    float a = W[i * 4];
    S[i] = fabsf(a);
    if (a < 0.0f) for (int r = 0; r < 3; ++r) U.m[r * 3 + i] = -U.m[r * 3 + i];
  }
  for (int i = 0; i < 3; ++i) S[i] *= scale;
  for (int i = 0; i < 3; ++i) {
    int pos = i; float mx = S[i];
    for (int k = i + 1; k < 3; ++k) if (S[k] > mx) { mx = S[k]; pos = k; }
    if (mx == 0.0f) break;
    if (pos != i) {
      float ts = S[i]; S[i] = S[pos]; S[pos] = ts;
      for (int r = 0; r < 3; ++r) {
        float tu = U.m[r * 3 + i]; U.m[r * 3 + i] = U.m[r * 3 + pos]; U.m[r * 3 + pos] = tu;
        float tv = V.m[r * 3 + i]; V.m[r * 3 + i] = V.m[r * 3 + pos]; V.m[r * 3 + pos] = tv;
      }
    }
  }
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
B2_HD Mat3 hat(const float* w) {
  Mat3 K;
  K.m[0] = 0.0f; K.m[1] = -w[2]; K.m[2] = w[1];
  K.m[3] = w[2]; K.m[4] = 0.0f; K.m[5] = -w[0];
  K.m[6] = -w[1]; K.m[7] = w[0]; K.m[8] = 0.0f;
  return K;
}
B2_HD float sin_f32(float x) {
#if defined(__CUDA_ARCH__)
  return (float)sin((double)x);  // correctly rounded in practice, as glibc's sinf is
#else
  return sinf(x);
#endif
}
B2_HD float cos_f32(float x) {
#if defined(__CUDA_ARCH__)
  return (float)cos((double)x);
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
B2_HD void ldlt6_solve(const float* Hin, const float* b, float* x) {
  float A[36];
  for (int i = 0; i < 36; ++i) A[i] = Hin[i];
  int perm[6];
  float tmp[6];
  for (int k = 0; k < 6; ++k) {
    int piv = k; float best = fabsf(A[k * 7]);
    for (int i = k + 1; i < 6; ++i) { float a = fabsf(A[i * 7]); if (a > best) { best = a; piv = i; } }
    perm[k] = piv;
    if (piv != k) {
      for (int c = 0; c < k; ++c) { float t = A[k * 6 + c]; A[k * 6 + c] = A[piv * 6 + c]; A[piv * 6 + c] = t; }
      for (int r = piv + 1; r < 6; ++r) { float t = A[r * 6 + k]; A[r * 6 + k] = A[r * 6 + piv]; A[r * 6 + piv] = t; }
      { float t = A[k * 7]; A[k * 7] = A[piv * 7]; A[piv * 7] = t; }
      for (int i = k + 1; i < piv; ++i) { float t = A[i * 6 + k]; A[i * 6 + k] = A[piv * 6 + i]; A[piv * 6 + i] = t; }
    }
    if (k > 0) {
      for (int c = 0; c < k; ++c) tmp[c] = A[c * 7] * A[k * 6 + c];
      float acc = A[k * 6] * tmp[0];
      for (int c = 1; c < k; ++c) acc = acc + A[k * 6 + c] * tmp[c];
      A[k * 7] -= acc;
      for (int r = k + 1; r < 6; ++r) {
        float a2 = A[r * 6] * tmp[0];
        for (int c = 1; c < k; ++c) a2 = a2 + A[r * 6 + c] * tmp[c];
        A[r * 6 + k] -= a2;
      }
    }
    float akk = A[k * 7];
    bool valid = fabsf(akk) > 0.0f;
    if (k == 0 && !valid) { for (int j = 0; j < 6; ++j) perm[j] = j; break; }
    if (valid) for (int r = k + 1; r < 6; ++r) A[r * 6 + k] /= akk;
  }
  float y[6];
  for (int i = 0; i < 6; ++i) y[i] = b[i];
  for (int k = 0; k < 6; ++k) if (perm[k] != k) { float t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  for (int i = 0; i < 6; ++i) { float acc = y[i]; for (int c = 0; c < i; ++c) acc -= A[i * 6 + c] * y[c]; y[i] = acc; }
  for (int i = 0; i < 6; ++i) { if (fabsf(A[i * 7]) > FLT_MIN) y[i] /= A[i * 7]; else y[i] = 0.0f; }
  for (int i = 5; i >= 0; --i) { float acc = y[i]; for (int c = i + 1; c < 6; ++c) acc -= A[c * 6 + i] * y[c]; y[i] = acc; }
  for (int k = 5; k >= 0; --k) if (perm[k] != k) { float t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
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
