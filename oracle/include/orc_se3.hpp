// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_se3.hpp — CPU restatement of util::SO3 / util::SE3 (float)
//   /root/reference/src/util/MathUtils.h:57-168, MathUtils.cpp:23-181
// Every SO3 constructed from a matrix is re-projected with JacobiSVD (MathUtils.cpp:86-99).
#pragma once
#include <cmath>
#include "orc_eigen.hpp"

namespace orc {

struct SO3f {
  float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  SO3f() = default;
  static SO3f FromMatrix(const float* R) {  // SO3::SO3(const Matrix3f&), MathUtils.cpp:86-99
    SO3f r;
    float U[9], S[3], V[9], Vt[9];
    jacobi_svd3<float>(R, U, S, V);
    mat3_transpose<float>(V, Vt);
    mat3_mul_mat3<float>(U, Vt, r.m);
    if (det3<float>(r.m) < 0.0f) {
      for (int i = 0; i < 3; ++i) U[i * 3 + 2] *= -1.0f;
      mat3_mul_mat3<float>(U, Vt, r.m);
    }
    return r;
  }
  static SO3f Identity() { const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; return FromMatrix(I); }  // MathUtils.h:93-95
  static void Hat(const float* w, float* K) {
    K[0] = 0; K[1] = -w[2]; K[2] = w[1];
    K[3] = w[2]; K[4] = 0; K[5] = -w[0];
    K[6] = -w[1]; K[7] = w[0]; K[8] = 0;
  }
  static SO3f Exp(const float* omega) {  // MathUtils.cpp:23-39
    const float theta = norm3<float>(omega);
    const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float M[9];
    if (theta < 1e-6f) {
      float K[9]; Hat(omega, K);
      for (int i = 0; i < 9; ++i) M[i] = I[i] + K[i];
      return FromMatrix(M);
    }
    const float theta_inv = 1.0f / theta;
    const float k[3] = {omega[0] * theta_inv, omega[1] * theta_inv, omega[2] * theta_inv};
    float K[9]; Hat(k, K);
    const float s = std::sin(theta);
    const float omc = 1.0f - std::cos(theta);
    float omcK[9], KK[9];
    for (int i = 0; i < 9; ++i) omcK[i] = omc * K[i];
    mat3_mul_mat3<float>(omcK, K, KK);
    for (int i = 0; i < 9; ++i) M[i] = (I[i] + s * K[i]) + KK[i];
    return FromMatrix(M);
  }
  SO3f operator*(const SO3f& o) const { float P[9]; mat3_mul_mat3<float>(m, o.m, P); return FromMatrix(P); }  // MathUtils.h:78-80
  SO3f Inverse() const { float Tt[9]; mat3_transpose<float>(m, Tt); return FromMatrix(Tt); }                 // MathUtils.h:88-90
  // MathUtils.cpp:41-84
  void Log(float* out) const {
    const float trace = sum3<float>(m[0], m[4], m[8]);
    const float cos_theta = (trace - 1.0f) * 0.5f;
    const float cc = std::max(-1.0f, std::min(1.0f, cos_theta));
    const float theta = std::acos(cc);
    if (theta < 1e-6f) { out[0] = m[7]; out[1] = m[2]; out[2] = m[3]; return; }  // Vee(R - I)
    const float sin_theta = std::sin(theta);
    if (std::abs(sin_theta) < 1e-6f) {
      float axis[3] = {0, 0, 0};
      int mi = 0;
      if (m[4] > m[0]) mi = 1;
      if (m[8] > m[mi * 4]) mi = 2;
      axis[mi] = std::sqrt((m[mi * 4] + 1.0f) * 0.5f);
      for (int i = 0; i < 3; ++i) if (i != mi) axis[i] = m[mi * 3 + i] / (2.0f * axis[mi]);
      float sk[3] = {(m[7] - m[5]) * 0.5f, (m[2] - m[6]) * 0.5f, (m[3] - m[1]) * 0.5f};
      float d = dot3<float>(axis, sk);
      if (d < 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; }
      out[0] = axis[0] * theta; out[1] = axis[1] * theta; out[2] = axis[2] * theta;
      return;
    }
    const float factor = theta / (2.0f * sin_theta);
    out[0] = factor * (m[7] - m[5]); out[1] = factor * (m[2] - m[6]); out[2] = factor * (m[3] - m[1]);
  }
};

struct SE3f {
  SO3f R;
  float t[3] = {0, 0, 0};
  SE3f() = default;
  SE3f(const SO3f& r, const float* tt) : R(r) { t[0] = tt[0]; t[1] = tt[1]; t[2] = tt[2]; }
  // SE3(const Matrix3f& R, const Vector3f& t): m_rotation(R) -> SO3(Matrix3f) re-projection (MathUtils.h:116-117)
  static SE3f FromRt(const float* Rm, const float* tt) { return SE3f(SO3f::FromMatrix(Rm), tt); }
  static SE3f FromMatrix4(const float* T) {  // MathUtils.cpp:109-112
    float Rm[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    float tt[3] = {T[3], T[7], T[11]};
    return FromRt(Rm, tt);
  }
  void Matrix(float* T) const {  // MathUtils.cpp:176-181
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = R.m[i * 3 + j]; T[i * 4 + 3] = t[i]; }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
  }
  SE3f operator*(const SE3f& o) const {  // MathUtils.h:144-147
    float rt[3]; mat3_mul_vec<float>(R.m, o.t, rt);
    float nt[3] = {t[0] + rt[0], t[1] + rt[1], t[2] + rt[2]};
    return SE3f(R * o.R, nt);
  }
  SE3f Inverse() const {  // MathUtils.h:155-158
    SO3f Ri = R.Inverse();
    float mt[3] = {-t[0], -t[1], -t[2]}, it[3];
    mat3_mul_vec<float>(Ri.m, mt, it);
    return SE3f(Ri, it);
  }
};

}  // namespace orc
