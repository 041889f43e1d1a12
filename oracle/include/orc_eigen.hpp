// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the
// product path (lidar_odometry_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// orc_eigen.hpp — restatement of the Eigen3 fixed-size numerics the reference's hot path relies on.
//
// PARITY UNPINNED at this boundary: Eigen is an un-vendored, un-pinned dependency of the reference
// (/root/reference/CMakeLists.txt:25, build.sh:29) and is absent from this image, so the exact
// low-order bits of JacobiSVD / LDLT / small products cannot be checked against the real library.
// The algorithms below restate Eigen 3.4.0 (the libeigen3-dev of Ubuntu 22.04/24.04) from its
// published sources: JacobiSVD two-sided sweeps with real_2x2_jacobi_svd + makeJacobi, unblocked
// pivoted LDLT, and the summation order of small coefficient-based products/reductions
// (size-3 float reductions associate as x0 + (x1 + x2); 4x4*4 float products are packet products
// accumulating column by column).  They are THE contract for the CUDA path.
//
// Call sites restated (reference file:line):
//   JacobiSVD<Matrix3f>        src/database/VoxelMap.cpp:239,348 ; src/util/MathUtils.cpp:88
//   Matrix<float,6,6>::ldlt()  src/optimization/IterativeClosestPointOptimizer.cpp:418
//   Matrix4f * Vector4f        src/util/PointCloudUtils.cpp:120-121
//   Matrix3f * Vector3f, dot   src/optimization/IterativeClosestPointOptimizer.cpp:368-386
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <algorithm>

namespace orc {

// ---- small reductions (Eigen redux_novec_unroller for size 3: func(x0, func(x1,x2))) -------------
template <class T> static inline T sum3(T a, T b, T c) {
  if constexpr (sizeof(T) == 4) return a + (b + c);  // float: no packet fits, halves-split unroller
  else return (a + b) + c;                           // double: Packet2d predux, then the scalar tail
}
template <class T> static inline T dot3(const T* a, const T* b) { return sum3<T>(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
template <class T> static inline T sqnorm3(const T* a) { return sum3<T>(a[0] * a[0], a[1] * a[1], a[2] * a[2]); }
template <class T> static inline T norm3(const T* a) { return std::sqrt(sqnorm3<T>(a)); }

// 3x3 (row-major storage here; values identical to Eigen's col-major object) times 3-vector.
template <class T> static inline void mat3_mul_vec(const T* M, const T* v, T* out) {
  T r[3];
  for (int i = 0; i < 3; ++i) r[i] = sum3<T>(M[i * 3 + 0] * v[0], M[i * 3 + 1] * v[1], M[i * 3 + 2] * v[2]);
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
template <class T> static inline void mat3_mul_mat3(const T* A, const T* B, T* C) {
  T r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = sum3<T>(A[i * 3 + 0] * B[0 * 3 + j], A[i * 3 + 1] * B[1 * 3 + j], A[i * 3 + 2] * B[2 * 3 + j]);
  for (int i = 0; i < 9; ++i) C[i] = r[i];
}
template <class T> static inline void mat3_transpose(const T* A, T* At) {
  T r[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[j * 3 + i] = A[i * 3 + j];
  for (int i = 0; i < 9; ++i) At[i] = r[i];
}
template <class T> static inline T det3(const T* m) {
  // Eigen bruteforce_det3_helper
  auto h = [&](int a, int b, int c) { return m[0 * 3 + a] * (m[1 * 3 + b] * m[2 * 3 + c] - m[1 * 3 + c] * m[2 * 3 + b]); };
  return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}

// Homogeneous 4x4 * (x,y,z,1): packet product, ((c0*x + c1*y) + c2*z) + c3*1  (PointCloudUtils.cpp:120-121)
static inline void transform_point_4x4(const float* T /*row-major 4x4*/, float x, float y, float z, float* out) {
  for (int r = 0; r < 3; ++r) out[r] = ((T[r * 4 + 0] * x + T[r * 4 + 1] * y) + T[r * 4 + 2] * z) + T[r * 4 + 3] * 1.0f;
}

// ---- Jacobi rotation helpers (Eigen/src/Jacobi/Jacobi.h) ---------------------------------------
template <class T> struct JRot { T c, s; };

template <class T> static inline bool make_jacobi(T x, T y, T z, JRot<T>& j) {
  T deno = T(2) * std::abs(y);
  if (deno < std::numeric_limits<T>::min()) { j.c = T(1); j.s = T(0); return false; }
  T tau = (x - z) / deno;
  T w = std::sqrt(tau * tau + T(1));
  T t = (tau > T(0)) ? T(1) / (tau + w) : T(1) / (tau - w);
  T sign_t = t > T(0) ? T(1) : T(-1);
  T n = T(1) / std::sqrt(t * t + T(1));
  j.s = -sign_t * (y / std::abs(y)) * std::abs(t) * n;
  j.c = n;
  return true;
}
template <class T> static inline JRot<T> jrot_transpose(const JRot<T>& j) { return JRot<T>{j.c, -j.s}; }
template <class T> static inline JRot<T> jrot_mul(const JRot<T>& a, const JRot<T>& b) {
  return JRot<T>{a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c};
}
// apply_rotation_in_the_plane on two strided vectors of length n
template <class T> static inline void rot_plane(T* x, int sx, T* y, int sy, int n, const JRot<T>& j) {
  if (j.c == T(1) && j.s == T(0)) return;
  for (int i = 0; i < n; ++i) {
    T xi = x[i * sx], yi = y[i * sy];
    x[i * sx] = j.c * xi + j.s * yi;
    y[i * sy] = -j.s * xi + j.c * yi;
  }
}

// real_2x2_jacobi_svd (Eigen/src/misc/RealSvd2x2.h)
template <class T> static inline void real_2x2_jacobi_svd(const T* W /*3x3 row-major*/, int p, int q, JRot<T>& j_left, JRot<T>& j_right) {
  T m[4] = {W[p * 3 + p], W[p * 3 + q], W[q * 3 + p], W[q * 3 + q]};
  JRot<T> rot1;
  T t = m[0] + m[3];
  T d = m[2] - m[1];
  if (std::abs(d) < std::numeric_limits<T>::min()) { rot1.s = T(0); rot1.c = T(1); }
  else {
    T u = t / d;
    T tmp = std::sqrt(T(1) + u * u);
    rot1.s = T(1) / tmp;
    rot1.c = u / tmp;
  }
  rot_plane<T>(&m[0], 1, &m[2], 1, 2, rot1);  // m.applyOnTheLeft(0,1,rot1)
  make_jacobi<T>(m[0], m[1], m[3], j_right);
  j_left = jrot_mul(rot1, jrot_transpose(j_right));
}

// JacobiSVD of a square 3x3 matrix (Eigen/src/SVD/JacobiSVD.h, 3.4.0 compute()).
// A, U, V row-major; S descending.  U and V are always accumulated (ComputeFullU|ComputeFullV).
// jacobi_svd3_work: the sweeps + sign fix + scale + sort on an already prepared work matrix W (U, V hold the starting bases:
// identity for a square input, the preconditioner's factors for a tall one).
template <class T> static inline void jacobi_svd3_work(T* W, T* U, T* S, T* V, T scale) {
  const T precision = T(2) * std::numeric_limits<T>::epsilon();
  const T considerAsZero = std::numeric_limits<T>::min();
  T maxDiag = std::max(std::abs(W[0]), std::max(std::abs(W[4]), std::abs(W[8])));
  bool finished = false;
  while (!finished) {
    finished = true;
    for (int p = 1; p < 3; ++p) {
      for (int q = 0; q < p; ++q) {
        T threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::abs(W[p * 3 + q]) > threshold || std::abs(W[q * 3 + p]) > threshold) {
          finished = false;
          JRot<T> jl, jr;
          real_2x2_jacobi_svd<T>(W, p, q, jl, jr);
          rot_plane<T>(&W[p * 3], 1, &W[q * 3], 1, 3, jl);                 // W.applyOnTheLeft(p,q,jl)
          rot_plane<T>(&U[p], 3, &U[q], 3, 3, jl);                         // U.applyOnTheRight(p,q,jl^T)
          rot_plane<T>(&W[p], 3, &W[q], 3, 3, jrot_transpose(jr));         // W.applyOnTheRight(p,q,jr)
          rot_plane<T>(&V[p], 3, &V[q], 3, 3, jrot_transpose(jr));         // V.applyOnTheRight(p,q,jr)
          maxDiag = std::max(maxDiag, std::max(std::abs(W[p * 3 + p]), std::abs(W[q * 3 + q])));
        }
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    T a = W[i * 3 + i];
    S[i] = std::abs(a);
    if (a < T(0)) for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
  }
  for (int i = 0; i < 3; ++i) S[i] *= scale;
  for (int i = 0; i < 3; ++i) {
    int pos = 0; T mx = S[i];
    for (int k = 1; k < 3 - i; ++k) if (S[i + k] > mx) { mx = S[i + k]; pos = k; }
    if (mx == T(0)) break;
    if (pos) {
      pos += i;
      std::swap(S[i], S[pos]);
      for (int r = 0; r < 3; ++r) { std::swap(U[r * 3 + pos], U[r * 3 + i]); std::swap(V[r * 3 + pos], V[r * 3 + i]); }
    }
  }
}
template <class T> static inline void jacobi_svd3(const T* A, T* U, T* S, T* V) {
  T scale = T(0);
  for (int i = 0; i < 9; ++i) { T a = std::abs(A[i]); if (a > scale || a != a) scale = a; }
  if (!std::isfinite(scale)) {  // InvalidInput: Eigen leaves U,V,S unspecified; we return identity/zero
    for (int i = 0; i < 9; ++i) { U[i] = (i % 4 == 0) ? T(1) : T(0); V[i] = U[i]; }
    S[0] = S[1] = S[2] = T(0);
    return;
  }
  if (scale == T(0)) scale = T(1);
  T W[9];
  for (int i = 0; i < 9; ++i) { W[i] = A[i] / scale; U[i] = (i % 4 == 0) ? T(1) : T(0); V[i] = U[i]; }
  jacobi_svd3_work<T>(W, U, S, V, scale);
}

// Pivoted LDLT (lower) of a 6x6 float matrix and solve; Eigen/src/Cholesky/LDLT.h unblocked + _solve_impl.
// Only the lower triangle of H is read.  Inner products are accumulated sequentially.
static inline void ldlt6_solve(const float* Hin /*row-major 6x6*/, const float* b, float* x) {
  const int N = 6;
  float m[36];
  for (int i = 0; i < 36; ++i) m[i] = Hin[i];
  int tr[N];
  float temp[N];
  auto M = [&](int r, int c) -> float& { return m[r * N + c]; };
  bool all_zero_diag = false;
  for (int k = 0; k < N; ++k) {
    int big = k; float best = std::abs(M(k, k));
    for (int i = k + 1; i < N; ++i) { float a = std::abs(M(i, i)); if (a > best) { best = a; big = i; } }
    tr[k] = big;
    if (k != big) {
      int s = N - big - 1;
      for (int c = 0; c < k; ++c) std::swap(M(k, c), M(big, c));
      for (int r = 0; r < s; ++r) std::swap(M(big + 1 + r, k), M(big + 1 + r, big));
      std::swap(M(k, k), M(big, big));
      for (int i = k + 1; i < big; ++i) { float t = M(i, k); M(i, k) = M(big, i); M(big, i) = t; }
    }
    int rs = N - k - 1;
    if (k > 0) {
      for (int c = 0; c < k; ++c) temp[c] = M(c, c) * M(k, c);
      float acc = 0.0f;
      for (int c = 0; c < k; ++c) acc = (c == 0) ? M(k, 0) * temp[0] : acc + M(k, c) * temp[c];
      M(k, k) -= acc;
      for (int r = 0; r < rs; ++r) {
        float a2 = 0.0f;
        for (int c = 0; c < k; ++c) a2 = (c == 0) ? M(k + 1 + r, 0) * temp[0] : a2 + M(k + 1 + r, c) * temp[c];
        M(k + 1 + r, k) -= a2;
      }
    }
    float akk = M(k, k);
    bool pivot_valid = std::abs(akk) > 0.0f;
    if (k == 0 && !pivot_valid) { for (int j = 0; j < N; ++j) tr[j] = j; all_zero_diag = true; break; }
    if (rs > 0 && pivot_valid) for (int r = 0; r < rs; ++r) M(k + 1 + r, k) /= akk;
  }
  (void)all_zero_diag;
  // solve: dst = P b ; L^-1 ; D^-1 (pseudo-inverse) ; L^-T ; P^T
  float y[N];
  for (int i = 0; i < N; ++i) y[i] = b[i];
  for (int k = 0; k < N; ++k) if (tr[k] != k) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < N; ++i) {  // unit-lower forward substitution
    float acc = y[i];
    for (int c = 0; c < i; ++c) acc -= M(i, c) * y[c];
    y[i] = acc;
  }
  const float tol = std::numeric_limits<float>::min();
  for (int i = 0; i < N; ++i) { if (std::abs(M(i, i)) > tol) y[i] /= M(i, i); else y[i] = 0.0f; }
  for (int i = N - 1; i >= 0; --i) {  // unit-upper (L^T) back substitution
    float acc = y[i];
    for (int c = i + 1; c < N; ++c) acc -= M(c, i) * y[c];
    y[i] = acc;
  }
  for (int k = N - 1; k >= 0; --k) if (tr[k] != k) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < N; ++i) x[i] = y[i];
}

// Smallest right-singular vector of an n x 3 double matrix (n = 5 in the KDTree path,
// src/optimization/IterativeClosestPointOptimizer.cpp:739-746, JacobiSVD<MatrixXd>(A, ComputeFullV)).
// PARITY UNPINNED: Eigen pre-conditions non-square input with ColPivHouseholderQR; we use a one-sided
// (Hestenes) Jacobi on the columns, which has the same high relative accuracy.  The sign of the
// normal is irrelevant downstream (|r| gates, r*J and J^T J are sign-invariant).
static inline void smallest_right_singular_vec_nx3(const double* A /*n x 3 row-major*/, int n, double* normal) {
  double B[15 * 3];
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (n > 15) n = 15;
  for (int i = 0; i < n * 3; ++i) B[i] = A[i];
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < n; ++i) { alpha += B[i * 3 + p] * B[i * 3 + p]; beta += B[i * 3 + q] * B[i * 3 + q]; gamma += B[i * 3 + p] * B[i * 3 + q]; }
        if (gamma == 0.0 || std::abs(gamma) <= 1e-15 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < n; ++i) { double bp = B[i * 3 + p], bq = B[i * 3 + q]; B[i * 3 + p] = c * bp - s * bq; B[i * 3 + q] = s * bp + c * bq; }
        for (int i = 0; i < 3; ++i) { double vp = V[i * 3 + p], vq = V[i * 3 + q]; V[i * 3 + p] = c * vp - s * vq; V[i * 3 + q] = s * vp + c * vq; }
      }
    }
    if (!rotated) break;
  }
  double nrm[3] = {0, 0, 0};
  for (int j = 0; j < 3; ++j) for (int i = 0; i < n; ++i) nrm[j] += B[i * 3 + j] * B[i * 3 + j];
  int best = 0;
  for (int j = 1; j < 3; ++j) if (nrm[j] < nrm[best]) best = j;
  for (int i = 0; i < 3; ++i) normal[i] = V[i * 3 + best];
}

}  // namespace orc
