// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the OKVIS hot-path math.
// Nothing in the product path (okvis_b200/, libokvis_b200.so) may include or link this.
// PARITY UNPINNED where noted: the reference (Ceres 1.9, Eigen, brisk 2.0.5) cannot be built in
// this environment and its tests hold no golden values for these functions; the restatement is
// validated relationally (numeric differentiation with the reference's own protocol).
//
// Small fixed-size linear algebra + the kinematics of
//   okvis_kinematics/include/okvis/kinematics/operators.hpp:61-112
//   okvis_kinematics/include/okvis/kinematics/implementation/Transformation.hpp:45-83, 246-299
// All matrices are row-major double arrays.  Quaternions are [x,y,z,w] (Eigen coeffs order).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace oko {

// ---- tiny dense helpers (row-major) ------------------------------------------------------
// C(m x n) = A(m x k) * B(k x n)
inline void matmul(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[i * k + p] * B[p * n + j];
      C[i * n + j] = s;
    }
}
// C(m x n) = A(m x k) * B^T, B is (n x k)
inline void matmul_nt(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[i * k + p] * B[j * k + p];
      C[i * n + j] = s;
    }
}
// C(k x n) = A^T * B, A is (m x k), B is (m x n)
inline void matmul_tn(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < m; ++p) s += A[p * k + i] * B[p * n + j];
      C[i * n + j] = s;
    }
}
inline void transpose(const double* A, double* At, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}
inline double norm3(const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// Eigen::LLT<>::matrixL() semantics for small fixed-size matrices: unblocked in-place lower
// Cholesky that stops at the first non-positive pivot and leaves the rest of the (lower) matrix
// untouched (Eigen/src/Cholesky/LLT.h, llt_inplace<Lower>::unblocked).  Returns -1 on success or
// the failing column.  L is the full n x n output with the strict upper part zeroed.
inline int llt_lower_eigen(const double* A, double* L, int n) {
  std::vector<double> M(A, A + n * n);
  int fail = -1;
  for (int k = 0; k < n; ++k) {
    double x = M[k * n + k];
    for (int p = 0; p < k; ++p) x -= M[k * n + p] * M[k * n + p];
    if (x <= 0.0) { fail = k; break; }
    x = std::sqrt(x);
    M[k * n + k] = x;
    for (int i = k + 1; i < n; ++i) {
      double s = M[i * n + k];
      for (int p = 0; p < k; ++p) s -= M[i * n + p] * M[k * n + p];
      M[i * n + k] = s / x;
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) L[i * n + j] = (j <= i) ? M[i * n + j] : 0.0;
  return fail;
}
// squareRootInformation_ = lltOfInformation.matrixL().transpose()  (e.g. PoseError.cpp:70-76)
inline int sqrt_information(const double* info, double* sqrtInfo, int n) {
  std::vector<double> L(n * n);
  int f = llt_lower_eigen(info, L.data(), n);
  transpose(L.data(), sqrtInfo, n, n);
  return f;
}

// General inverse by LU with partial pivoting (what Eigen's MatrixBase::inverse() does for n > 4).
inline bool inverse_lu(const double* A, double* Ainv, int n) {
  std::vector<double> M(n * 2 * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      M[i * 2 * n + j] = A[i * n + j];
      M[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(M[c * 2 * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(M[r * 2 * n + c]) > best) { best = std::fabs(M[r * 2 * n + c]); piv = r; }
    if (best == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(M[c * 2 * n + j], M[piv * 2 * n + j]);
    const double d = 1.0 / M[c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) M[c * 2 * n + j] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[r * 2 * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; ++j) M[r * 2 * n + j] -= f * M[c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Ainv[i * n + j] = M[i * 2 * n + n + j];
  return true;
}

// ---- kinematics --------------------------------------------------------------------------
// operators.hpp:61-76
inline void crossMx(const double* v, double* C) {
  C[0] = 0.0;   C[1] = -v[2]; C[2] = v[1];
  C[3] = v[2];  C[4] = 0.0;   C[5] = -v[0];
  C[6] = -v[1]; C[7] = v[0];  C[8] = 0.0;
}
// Eigen quaternion product a*b (Hamilton), coeffs [x,y,z,w]
inline void qmul(const double* a, const double* b, double* o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
// Eigen::Quaternion::inverse(): conjugate / squaredNorm
inline void qinv(const double* q, double* o) {
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2;
}
inline void qnormalize(double* q) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
// Eigen::Quaternion::toRotationMatrix (no normalisation)
inline void q2R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// operators.hpp:86-96: q_AB*q_BC = plus(q_AB)*q_BC.coeffs()
inline void qplusMat(const double* q, double* Q) {
  Q[0] = q[3];  Q[1] = -q[2]; Q[2] = q[1];  Q[3] = q[0];
  Q[4] = q[2];  Q[5] = q[3];  Q[6] = -q[0]; Q[7] = q[1];
  Q[8] = -q[1]; Q[9] = q[0];  Q[10] = q[3]; Q[11] = q[2];
  Q[12] = -q[0]; Q[13] = -q[1]; Q[14] = -q[2]; Q[15] = q[3];
}
// operators.hpp:100-110: q_AB*q_BC = oplus(q_BC)*q_AB.coeffs()
inline void qoplusMat(const double* q, double* Q) {
  Q[0] = q[3];  Q[1] = q[2];  Q[2] = -q[1]; Q[3] = q[0];
  Q[4] = -q[2]; Q[5] = q[3];  Q[6] = q[0];  Q[7] = q[1];
  Q[8] = q[1];  Q[9] = -q[0]; Q[10] = q[3]; Q[11] = q[2];
  Q[12] = -q[0]; Q[13] = -q[1]; Q[14] = -q[2]; Q[15] = q[3];
}
// Transformation.hpp(impl):45-58 and ode.hpp:58-71 (identical)
inline double sinc(double x) {
  if (std::fabs(x) > 1e-6) return std::sin(x) / x;
  const double c_2 = 1.0 / 6.0, c_4 = 1.0 / 120.0, c_6 = 1.0 / 5040.0;
  const double x_2 = x * x, x_4 = x_2 * x_2, x_6 = x_2 * x_2 * x_2;
  return 1.0 - c_2 * x_2 + c_4 * x_4 - c_6 * x_6;
}
// Transformation.hpp(impl):60-67
inline void deltaQ(const double* dAlpha, double* dq) {
  const double halfnorm = 0.5 * norm3(dAlpha);
  const double s = sinc(halfnorm) * 0.5;
  dq[0] = s * dAlpha[0]; dq[1] = s * dAlpha[1]; dq[2] = s * dAlpha[2];
  dq[3] = std::cos(halfnorm);
}
// Transformation.hpp(impl):69-83
inline void rightJacobian(const double* PhiVec, double* R) {
  const double Phi = norm3(PhiVec);
  double X[9], X2[9];
  crossMx(PhiVec, X);
  matmul(X, X, X2, 3, 3, 3);
  double a, b;
  if (Phi < 1.0e-4) {
    a = -0.5; b = 1.0 / 6.0;
  } else {
    const double Phi2 = Phi * Phi, Phi3 = Phi2 * Phi;
    a = -(1.0 - std::cos(Phi)) / Phi2;
    b = (Phi - std::sin(Phi)) / Phi3;
  }
  for (int i = 0; i < 9; ++i) R[i] = a * X[i] + b * X2[i];
  R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
}

// okvis::kinematics::Transformation: r + unit q, cached C.  Constructor normalises q
// (Transformation.hpp(impl):105-112).
struct Transformation {
  double r[3];
  double q[4];
  double C[9];
  Transformation() { r[0] = r[1] = r[2] = 0; q[0] = q[1] = q[2] = 0; q[3] = 1; q2R(q, C); }
  Transformation(const double* r_, const double* q_) { set(r_, q_); }
  explicit Transformation(const double* pose7) { set(pose7, pose7 + 3); }
  void set(const double* r_, const double* q_) {
    std::memcpy(r, r_, 24); std::memcpy(q, q_, 32);
    qnormalize(q); q2R(q, C);
  }
  // Transformation.hpp(impl):171-173
  Transformation inverse() const {
    double ri[3], qi[4];
    for (int i = 0; i < 3; ++i) ri[i] = -(C[0 * 3 + i] * r[0] + C[1 * 3 + i] * r[1] + C[2 * 3 + i] * r[2]);
    qinv(q, qi);
    return Transformation(ri, qi);
  }
  // Transformation.hpp(impl):215-218
  Transformation operator*(const Transformation& rhs) const {
    double rr[3], qq[4];
    for (int i = 0; i < 3; ++i)
      rr[i] = C[i * 3 + 0] * rhs.r[0] + C[i * 3 + 1] * rhs.r[1] + C[i * 3 + 2] * rhs.r[2] + r[i];
    qmul(q, rhs.q, qq);
    return Transformation(rr, qq);
  }
  // Transformation.hpp(impl):246-259
  void oplus(const double* delta6) {
    r[0] += delta6[0]; r[1] += delta6[1]; r[2] += delta6[2];
    double dq[4], qn[4];
    deltaQ(delta6 + 3, dq);
    qmul(dq, q, qn);
    std::memcpy(q, qn, 32);
    qnormalize(q);
    q2R(q, C);
  }
  void to7(double* p) const { std::memcpy(p, r, 24); std::memcpy(p + 3, q, 32); }
};

// PoseLocalParameterization::plus (okvis_ceres/src/PoseLocalParameterization.cpp:60-87)
inline void pose_plus(const double* x, const double* delta, double* xpd) {
  Transformation T(x);
  T.oplus(delta);
  T.to7(xpd);
}
// PoseLocalParameterization::minus (:103-116): delta = x_plus_delta [-] x
inline void pose_minus(const double* x, const double* xpd, double* delta) {
  delta[0] = xpd[0] - x[0]; delta[1] = xpd[1] - x[1]; delta[2] = xpd[2] - x[2];
  double qi[4], dq[4];
  qinv(x + 3, qi);
  qmul(xpd + 3, qi, dq);
  delta[3] = 2 * dq[0]; delta[4] = 2 * dq[1]; delta[5] = 2 * dq[2];
}
// PoseLocalParameterization::liftJacobian (:133-147), 6x7 row-major
inline void pose_lift_jacobian(const double* x, double* J) {
  std::memset(J, 0, sizeof(double) * 42);
  J[0 * 7 + 0] = J[1 * 7 + 1] = J[2 * 7 + 2] = 1.0;
  const double qinv_[4] = {-x[3], -x[4], -x[5], x[6]};
  double Q[16];
  qoplusMat(qinv_, Q);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) J[(3 + i) * 7 + 3 + j] = 2.0 * Q[i * 4 + j];
}
// Transformation::oplusJacobian (Transformation.hpp(impl):271-285), 7x6 row-major
inline void pose_plus_jacobian(const double* x, double* J) {
  std::memset(J, 0, sizeof(double) * 42);
  J[0 * 6 + 0] = J[1 * 6 + 1] = J[2 * 6 + 2] = 1.0;
  Transformation T(x);
  double Q[16];
  qoplusMat(T.q, Q);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) J[(3 + i) * 6 + 3 + j] = 0.5 * Q[i * 4 + j];
}

}  // namespace oko
