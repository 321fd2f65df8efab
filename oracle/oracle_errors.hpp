// TEST INFRASTRUCTURE ONLY -- CPU restatement of the OKVIS error functors and camera models.
// Each function cites the reference file:line it follows.  PARITY UNPINNED (see oracle_math.hpp):
// validated by numeric differentiation with the reference's own test protocol, not by golden values.
#pragma once
#include "../include/okvis_b200.h"
#include "oracle_math.hpp"

namespace oko {

// ---- distortion models -------------------------------------------------------------------
// Returns false where the reference's distort() returns false (only radtan8, rho > 9).
// J is the 2x2 Jacobian wrt the undistorted point (row-major), may be null.
inline bool distort(const okb_camera& cam, const double* u, double* d, double* J) {
  const double u0 = u[0], u1 = u[1];
  switch (cam.model) {
    case OKB_DIST_NONE:
      d[0] = u0; d[1] = u1;
      if (J) { J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; }
      return true;
    case OKB_DIST_RADTAN: {
      // okvis_cv/.../implementation/RadialTangentialDistortion.hpp:104-152
      const double k1 = cam.dist[0], k2 = cam.dist[1], p1 = cam.dist[2], p2 = cam.dist[3];
      const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
      const double rho_u = mx_u + my_u;
      const double rad_dist_u = k1 * rho_u + k2 * rho_u * rho_u;
      d[0] = u0 + u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
      d[1] = u1 + u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
      if (J) {
        J[0] = 1 + rad_dist_u + k1 * 2.0 * mx_u + k2 * rho_u * 4 * mx_u + 2.0 * p1 * u1 + 6 * p2 * u0;
        J[2] = k1 * 2.0 * u0 * u1 + k2 * 4 * rho_u * u0 * u1 + p1 * 2.0 * u0 + 2.0 * p2 * u1;
        J[1] = J[2];
        J[3] = 1 + rad_dist_u + k1 * 2.0 * my_u + k2 * rho_u * 4 * my_u + 6 * p1 * u1 + 2.0 * p2 * u0;
      }
      return true;
    }
    case OKB_DIST_EQUIDISTANT: {
      // okvis_cv/.../implementation/EquidistantDistortion.hpp:105-206
      const double k1 = cam.dist[0], k2 = cam.dist[1], k3 = cam.dist[2], k4 = cam.dist[3];
      const double r = std::sqrt(u0 * u0 + u1 * u1);
      const double theta = std::atan(r);
      const double theta2 = theta * theta, theta4 = theta2 * theta2;
      const double theta6 = theta4 * theta2, theta8 = theta4 * theta4;
      const double thetad = theta * (1 + k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8);
      const double scaling = (r > 1e-8) ? thetad / r : 1.0;
      d[0] = scaling * u0; d[1] = scaling * u1;
      if (J) {
        if (r > 1e-8) {
          double t2, t3, t4, t6, t7, t8, t9, t11, t17, t18, t19, t20, t25;
          t2 = u0 * u0; t3 = u1 * u1; t4 = t2 + t3;
          t6 = std::atan(std::sqrt(t4));
          t7 = t6 * t6;
          t8 = 1.0 / std::sqrt(t4);
          t9 = t7 * t7;
          t11 = 1.0 / ((t2 + t3) + 1.0);
          t17 = (((k1 * t7 + k2 * t9) + k3 * t7 * t9) + k4 * (t9 * t9)) + 1.0;
          t18 = 1.0 / t4;
          t19 = 1.0 / std::sqrt(t4 * t4 * t4);
          t20 = t6 * t8 * t17;
          t25 = ((k2 * t6 * t7 * t8 * t11 * u1 * 4.0 + k3 * t6 * t8 * t9 * t11 * u1 * 6.0) +
                 k4 * t6 * t7 * t8 * t9 * t11 * u1 * 8.0) + k1 * t6 * t8 * t11 * u1 * 2.0;
          t4 = ((k2 * t6 * t7 * t8 * t11 * u0 * 4.0 + k3 * t6 * t8 * t9 * t11 * u0 * 6.0) +
                k4 * t6 * t7 * t8 * t9 * t11 * u0 * 8.0) + k1 * t6 * t8 * t11 * u0 * 2.0;
          t7 = t11 * t17 * t18 * u0 * u1;
          J[1] = (t7 + t6 * t8 * t25 * u0) - t6 * t17 * t19 * u0 * u1;
          J[3] = ((t20 - t3 * t6 * t17 * t19) + t3 * t11 * t17 * t18) + t6 * t8 * t25 * u1;
          J[0] = ((t20 - t2 * t6 * t17 * t19) + t2 * t11 * t17 * t18) + t6 * t8 * t4 * u0;
          J[2] = (t7 + t6 * t8 * t4 * u1) - t6 * t17 * t19 * u0 * u1;
        } else {
          J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1;
        }
      }
      return true;
    }
    case OKB_DIST_RADTAN8: {
      // okvis_cv/.../implementation/RadialTangentialDistortion8.hpp:103-160
      const double k1 = cam.dist[0], k2 = cam.dist[1], p1 = cam.dist[2], p2 = cam.dist[3];
      const double k3 = cam.dist[4], k4 = cam.dist[5], k5 = cam.dist[6], k6 = cam.dist[7];
      const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
      const double rho_u = mx_u + my_u;
      if (rho_u > 9.0) return false;
      const double c = rho_u * (k4 + rho_u * (k5 + k6 * rho_u)) + 1.0;
      const double c2 = c * c;
      const double rad_dist_u = (1.0 + ((k3 * rho_u + k2) * rho_u + k1) * rho_u) /
                                (1.0 + ((k6 * rho_u + k5) * rho_u + k4) * rho_u);
      d[0] = u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
      d[1] = u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
      if (J) {
        const double num = rho_u * (k1 + rho_u * (k2 + k3 * rho_u)) + 1.0;
        const double den = rho_u * (k4 + rho_u * (k5 + k6 * rho_u)) + 1.0;
        auto dn = [&](double ua) {  // d(num)/d(ua)
          return rho_u * (ua * (k2 + k3 * rho_u) * 2.0 + k3 * ua * rho_u * 2.0) + ua * (k1 + rho_u * (k2 + k3 * rho_u)) * 2.0;
        };
        auto dd = [&](double ua) {  // d(den)/d(ua)
          return rho_u * (ua * (k5 + k6 * rho_u) * 2.0 + k6 * ua * rho_u * 2.0) + ua * (k4 + rho_u * (k5 + k6 * rho_u)) * 2.0;
        };
        J[0] = p1 * u1 * 2.0 + p2 * u0 * 6.0 + num / den + (u0 * dn(u0)) / den - u0 * dd(u0) * num * 1.0 / c2;
        J[1] = p1 * u0 * 2.0 + p2 * u1 * 2.0 + (u0 * dn(u1)) / den - u0 * dd(u1) * num * 1.0 / c2;
        J[2] = p1 * u0 * 2.0 + p2 * u1 * 2.0 + (u1 * dn(u0)) / den - u1 * dd(u0) * num * 1.0 / c2;
        J[3] = p1 * u1 * 6.0 + p2 * u0 * 2.0 + num / den + (u1 * dn(u1)) / den - u1 * dd(u1) * num * 1.0 / c2;
      }
      return true;
    }
  }
  return false;
}

// PinholeCamera::project (okvis_cv/.../implementation/PinholeCamera.hpp:147-226), without the
// image-bounds status.  Returns false when the reference returns Invalid: |z| < 1e-12 (outputs
// uninitialised in the reference -- we define them as zero) or distortion failure (radtan8: the
// reference leaves the distorted point uninitialised -- we define it as zero).
inline bool project(const okb_camera& cam, const double* pt, double* ip, double* J /*2x3 or null*/) {
  if (std::fabs(pt[2]) < 1.0e-12) {
    ip[0] = ip[1] = 0.0;
    if (J) std::memset(J, 0, 6 * sizeof(double));
    return false;
  }
  const double rz = 1.0 / pt[2];
  const double rz2 = rz * rz;
  const double u[2] = {pt[0] * rz, pt[1] * rz};
  double d[2] = {0, 0}, Jd[4] = {0, 0, 0, 0};
  const bool ok = distort(cam, u, d, J ? Jd : nullptr);
  if (J) {
    J[0] = cam.fu * Jd[0] * rz;
    J[1] = cam.fu * Jd[1] * rz;
    J[2] = -cam.fu * (pt[0] * Jd[0] + pt[1] * Jd[1]) * rz2;
    J[3] = cam.fv * Jd[2] * rz;
    J[4] = cam.fv * Jd[3] * rz;
    J[5] = -cam.fv * (pt[0] * Jd[2] + pt[1] * Jd[3]) * rz2;
  }
  ip[0] = cam.fu * d[0] + cam.cu;
  ip[1] = cam.fv * d[1] + cam.cv;
  return ok;
}

// PinholeCamera::projectHomogeneous (:345-378): w<0 projects -xyz, Jacobian NOT negated, 4th column 0.
inline bool project_homogeneous(const okb_camera& cam, const double* hp, double* ip, double* J /*2x4 or null*/) {
  double head[3] = {hp[0], hp[1], hp[2]};
  if (hp[3] < 0) { head[0] = -head[0]; head[1] = -head[1]; head[2] = -head[2]; }
  double J3[6];
  const bool ok = project(cam, head, ip, J ? J3 : nullptr);
  if (J) {
    for (int r = 0; r < 2; ++r) {
      J[r * 4 + 0] = J3[r * 3 + 0]; J[r * 4 + 1] = J3[r * 3 + 1]; J[r * 4 + 2] = J3[r * 3 + 2];
      J[r * 4 + 3] = 0.0;
    }
  }
  return ok;
}

// ---- ReprojectionError -------------------------------------------------------------------
// okvis_ceres/include/okvis/ceres/implementation/ReprojectionError.hpp:87-242, isotropic information
// (sqrt_info * I2, Estimator.hpp(impl):62-65).  Minimal Jacobians, row-major: J0 2x6, J1 2x3, J2 2x6.
// Any Jacobian pointer may be null.
inline void reprojection_error(const okb_camera& cam, const double* pose, const double* hp_W,
                               const double* ext, const double* z, double sqrt_info,
                               double* res, double* J0, double* J1, double* J2) {
  const double* t_WS_W = pose;
  const double* t_SC_S = ext;
  double C_SC[9], C_CS[9], C_WS[9], C_SW[9];
  q2R(ext + 3, C_SC); transpose(C_SC, C_CS, 3, 3);
  q2R(pose + 3, C_WS); transpose(C_WS, C_SW, 3, 3);
  double T_CS[16] = {0}, T_SW[16] = {0};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) { T_CS[i * 4 + j] = C_CS[i * 3 + j]; T_SW[i * 4 + j] = C_SW[i * 3 + j]; }
    T_CS[i * 4 + 3] = -(C_CS[i * 3] * t_SC_S[0] + C_CS[i * 3 + 1] * t_SC_S[1] + C_CS[i * 3 + 2] * t_SC_S[2]);
    T_SW[i * 4 + 3] = -(C_SW[i * 3] * t_WS_W[0] + C_SW[i * 3 + 1] * t_WS_W[1] + C_SW[i * 3 + 2] * t_WS_W[2]);
  }
  T_CS[15] = 1.0; T_SW[15] = 1.0;
  double hp_S[4], hp_C[4];
  matmul(T_SW, hp_W, hp_S, 4, 4, 1);
  matmul(T_CS, hp_S, hp_C, 4, 4, 1);

  const bool wantJ = (J0 || J1 || J2);
  double kp[2], Jh[8], Jh_weighted[8];
  project_homogeneous(cam, hp_C, kp, wantJ ? Jh : nullptr);
  if (wantJ) for (int i = 0; i < 8; ++i) Jh_weighted[i] = sqrt_info * Jh[i];
  res[0] = sqrt_info * (z[0] - kp[0]);
  res[1] = sqrt_info * (z[1] - kp[1]);

  bool valid = true;
  if (std::fabs(hp_C[3]) > 1.0e-8) {
    if (hp_C[2] / hp_C[3] < 0.2) valid = false;
  }
  if (J0) {
    const double p[3] = {hp_W[0] - t_WS_W[0] * hp_W[3], hp_W[1] - t_WS_W[1] * hp_W[3], hp_W[2] - t_WS_W[2] * hp_W[3]};
    double J[24] = {0}, px[9], Cpx[9];
    crossMx(p, px);
    matmul(C_SW, px, Cpx, 3, 3, 3);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { J[i * 6 + j] = C_SW[i * 3 + j] * hp_W[3]; J[i * 6 + 3 + j] = -Cpx[i * 3 + j]; }
    double JT[8];
    matmul(Jh_weighted, T_CS, JT, 2, 4, 4);
    matmul(JT, J, J0, 2, 4, 6);
    if (!valid) std::memset(J0, 0, 12 * sizeof(double));
  }
  if (J1) {
    double T_CW[16], Jf[8];
    matmul(T_CS, T_SW, T_CW, 4, 4, 4);
    matmul(Jh_weighted, T_CW, Jf, 2, 4, 4);
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 3; ++c) J1[r * 3 + c] = valid ? -Jf[r * 4 + c] : 0.0;
  }
  if (J2) {
    const double p[3] = {hp_S[0] - t_SC_S[0] * hp_S[3], hp_S[1] - t_SC_S[1] * hp_S[3], hp_S[2] - t_SC_S[2] * hp_S[3]};
    double J[24] = {0}, px[9], Cpx[9];
    crossMx(p, px);
    matmul(C_CS, px, Cpx, 3, 3, 3);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { J[i * 6 + j] = C_CS[i * 3 + j] * hp_S[3]; J[i * 6 + 3 + j] = -Cpx[i * 3 + j]; }
    matmul(Jh_weighted, J, J2, 2, 4, 6);
    if (!valid) std::memset(J2, 0, 12 * sizeof(double));
  }
}

// ---- ImuError ----------------------------------------------------------------------------
// The mutable preintegration cache of ImuError (okvis_ceres/include/okvis/ceres/ImuError.hpp:251-276).
struct ImuCache {
  double Delta_q[4];
  double C_integral[9], C_doubleintegral[9];
  double acc_integral[3], acc_doubleintegral[3];
  double cross[9];
  double dalpha_db_g[9], dv_db_g[9], dp_db_g[9];
  double P_delta[225];
  double information[225], squareRootInformation[225];
  double sb_ref[9];
  bool redo = true;
  int redoCounter = 0;
};

inline double ns_to_sec(int64_t ns) {
  // okvis::Duration::toSec(): sec + 1e-9*nsec with sec,nsec split (okvis_time/include/okvis/Duration.hpp)
  int64_t sec = ns / 1000000000LL, nsec = ns % 1000000000LL;
  if (nsec < 0) { nsec += 1000000000LL; sec -= 1; }
  return (double)sec + 1e-9 * (double)nsec;
}

// Shared body of ImuError::redoPreintegration (okvis_ceres/src/ImuError.cpp:76-284) and the static
// ImuError::propagation (:287-504).  `preint` selects the variant:
//   true : dalpha_db_g += C_1*rightJacobian(omega*dt)*dt (:200), sigma2_v = dt*sigma_a_c^2 (:231)
//   false: dalpha_db_g += dt*C_1 (:412),                 sigma2_v = dt*sigma_a_c*params.sigma_a_c (:438)
// Returns the number of integration steps, -1 if the samples do not reach t1.
inline int imu_integrate(const okb_imu_sample* s, int n, const okb_imu_params& prm, int64_t t0, int64_t t1,
                         const double* sb, bool preint, bool wantCov, ImuCache& c, double* Delta_t_out) {
  int64_t time = t0;
  const int64_t end = t1;
  if (!(s[n - 1].t_ns >= end)) return -1;
  c.Delta_q[0] = c.Delta_q[1] = c.Delta_q[2] = 0; c.Delta_q[3] = 1;
  std::memset(c.C_integral, 0, sizeof c.C_integral);
  std::memset(c.C_doubleintegral, 0, sizeof c.C_doubleintegral);
  std::memset(c.acc_integral, 0, sizeof c.acc_integral);
  std::memset(c.acc_doubleintegral, 0, sizeof c.acc_doubleintegral);
  std::memset(c.cross, 0, sizeof c.cross);
  std::memset(c.dalpha_db_g, 0, sizeof c.dalpha_db_g);
  std::memset(c.dv_db_g, 0, sizeof c.dv_db_g);
  std::memset(c.dp_db_g, 0, sizeof c.dp_db_g);
  std::memset(c.P_delta, 0, sizeof c.P_delta);
  double Delta_t = 0;
  bool hasStarted = false;
  int i = 0;
  for (int it = 0; it < n; ++it) {
    double omega_S_0[3], acc_S_0[3], omega_S_1[3], acc_S_1[3];
    const bool last = (it + 1 == n);
    for (int k = 0; k < 3; ++k) {
      omega_S_0[k] = s[it].gyro[k]; acc_S_0[k] = s[it].acc[k];
      // the reference dereferences (it+1) even at end(); the values are unused there because
      // nexttime == t1 terminates the loop.  We read the last sample instead.
      omega_S_1[k] = s[last ? it : it + 1].gyro[k]; acc_S_1[k] = s[last ? it : it + 1].acc[k];
    }
    int64_t nexttime = last ? t1 : s[it + 1].t_ns;
    double dt = ns_to_sec(nexttime - time);
    if (end < nexttime) {
      const double interval = ns_to_sec(nexttime - s[it].t_ns);
      nexttime = t1;
      dt = ns_to_sec(nexttime - time);
      const double r = dt / interval;
      for (int k = 0; k < 3; ++k) {
        omega_S_1[k] = (1.0 - r) * omega_S_0[k] + r * omega_S_1[k];
        acc_S_1[k] = (1.0 - r) * acc_S_0[k] + r * acc_S_1[k];
      }
    }
    if (dt <= 0.0) continue;
    Delta_t += dt;
    if (!hasStarted) {
      hasStarted = true;
      const double r = dt / ns_to_sec(nexttime - s[it].t_ns);
      for (int k = 0; k < 3; ++k) {
        omega_S_0[k] = r * omega_S_0[k] + (1.0 - r) * omega_S_1[k];
        acc_S_0[k] = r * acc_S_0[k] + (1.0 - r) * acc_S_1[k];
      }
    }
    double sigma_g_c = prm.sigma_g_c, sigma_a_c = prm.sigma_a_c;
    bool gsat = false, asat = false;
    for (int k = 0; k < 3; ++k) {
      if (std::fabs(omega_S_0[k]) > prm.g_max || std::fabs(omega_S_1[k]) > prm.g_max) gsat = true;
      if (std::fabs(acc_S_0[k]) > prm.a_max || std::fabs(acc_S_1[k]) > prm.a_max) asat = true;
    }
    if (gsat) sigma_g_c *= 100;
    if (asat) sigma_a_c *= 100;

    double omega_S_true[3], acc_S_true[3];
    for (int k = 0; k < 3; ++k) {
      omega_S_true[k] = 0.5 * (omega_S_0[k] + omega_S_1[k]) - sb[3 + k];
      acc_S_true[k] = 0.5 * (acc_S_0[k] + acc_S_1[k]) - sb[6 + k];
    }
    const double theta_half = norm3(omega_S_true) * 0.5 * dt;
    const double sinc_theta_half = sinc(theta_half);
    const double cos_theta_half = std::cos(theta_half);
    double dq[4] = {sinc_theta_half * omega_S_true[0] * 0.5 * dt, sinc_theta_half * omega_S_true[1] * 0.5 * dt,
                    sinc_theta_half * omega_S_true[2] * 0.5 * dt, cos_theta_half};
    double Delta_q_1[4];
    qmul(c.Delta_q, dq, Delta_q_1);
    double C[9], C_1[9], CC[9];
    q2R(c.Delta_q, C);
    q2R(Delta_q_1, C_1);
    for (int k = 0; k < 9; ++k) CC[k] = C[k] + C_1[k];
    double CCa[3];
    matmul(CC, acc_S_true, CCa, 3, 3, 1);
    double C_integral_1[9], acc_integral_1[3];
    for (int k = 0; k < 9; ++k) C_integral_1[k] = c.C_integral[k] + 0.5 * CC[k] * dt;
    for (int k = 0; k < 3; ++k) acc_integral_1[k] = c.acc_integral[k] + 0.5 * CCa[k] * dt;
    for (int k = 0; k < 9; ++k) c.C_doubleintegral[k] += c.C_integral[k] * dt + 0.25 * CC[k] * dt * dt;
    double acc_dd_inc[3];
    for (int k = 0; k < 3; ++k) {
      acc_dd_inc[k] = c.acc_integral[k] * dt + 0.25 * CCa[k] * dt * dt;
      c.acc_doubleintegral[k] += acc_dd_inc[k];
    }
    // Jacobian parts
    double wdt[3] = {omega_S_true[0] * dt, omega_S_true[1] * dt, omega_S_true[2] * dt};
    double Jr[9];
    rightJacobian(wdt, Jr);
    if (preint) {
      double CJ[9];
      matmul(C_1, Jr, CJ, 3, 3, 3);
      for (int k = 0; k < 9; ++k) c.dalpha_db_g[k] += CJ[k] * dt;
    } else {
      for (int k = 0; k < 9; ++k) c.dalpha_db_g[k] += dt * C_1[k];
    }
    double dqi[4], Rdqi[9], cross_1[9];
    qinv(dq, dqi);
    q2R(dqi, Rdqi);
    matmul(Rdqi, c.cross, cross_1, 3, 3, 3);
    for (int k = 0; k < 9; ++k) cross_1[k] += Jr[k] * dt;
    double acc_S_x[9], t1m[9], t2m[9], A1[9], A2[9], sumA[9];
    crossMx(acc_S_true, acc_S_x);
    matmul(C, acc_S_x, t1m, 3, 3, 3);
    matmul(t1m, c.cross, A1, 3, 3, 3);
    matmul(C_1, acc_S_x, t2m, 3, 3, 3);
    matmul(t2m, cross_1, A2, 3, 3, 3);
    for (int k = 0; k < 9; ++k) sumA[k] = A1[k] + A2[k];
    double dv_db_g_1[9];
    for (int k = 0; k < 9; ++k) dv_db_g_1[k] = c.dv_db_g[k] + 0.5 * dt * sumA[k];
    double F09[9];
    for (int k = 0; k < 9; ++k) {
      F09[k] = dt * c.dv_db_g[k] + 0.25 * dt * dt * sumA[k];
      c.dp_db_g[k] += F09[k];
    }
    if (wantCov) {
      double F[225];
      std::memset(F, 0, sizeof F);
      for (int k = 0; k < 15; ++k) F[k * 15 + k] = 1.0;
      auto setblk = [&](int r0, int c0, const double* B, double sgn) {
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) F[(r0 + a) * 15 + c0 + b] = sgn * B[a * 3 + b];
      };
      double X[9];
      crossMx(acc_dd_inc, X); setblk(0, 3, X, -1.0);
      double Idt[9] = {dt, 0, 0, 0, dt, 0, 0, 0, dt}; setblk(0, 6, Idt, 1.0);
      setblk(0, 9, F09, 1.0);
      double F012[9];
      for (int k = 0; k < 9; ++k) F012[k] = -c.C_integral[k] * dt + 0.25 * CC[k] * dt * dt;
      setblk(0, 12, F012, 1.0);
      double F39[9];
      for (int k = 0; k < 9; ++k) F39[k] = -dt * C_1[k];
      setblk(3, 9, F39, 1.0);
      double v63[3] = {0.5 * CCa[0] * dt, 0.5 * CCa[1] * dt, 0.5 * CCa[2] * dt};
      crossMx(v63, X); setblk(6, 3, X, -1.0);
      double F69[9];
      for (int k = 0; k < 9; ++k) F69[k] = 0.5 * dt * sumA[k];
      setblk(6, 9, F69, 1.0);
      double F612[9];
      for (int k = 0; k < 9; ++k) F612[k] = -0.5 * CC[k] * dt;
      setblk(6, 12, F612, 1.0);
      double FP[225], P2[225];
      matmul(F, c.P_delta, FP, 15, 15, 15);
      matmul_nt(FP, F, P2, 15, 15, 15);
      std::memcpy(c.P_delta, P2, sizeof P2);
      const double sigma2_dalpha = dt * sigma_g_c * sigma_g_c;
      const double sigma2_v = preint ? dt * sigma_a_c * sigma_a_c : dt * sigma_a_c * prm.sigma_a_c;
      const double sigma2_p = 0.5 * dt * dt * sigma2_v;
      const double sigma2_b_g = dt * prm.sigma_gw_c * prm.sigma_gw_c;
      const double sigma2_b_a = dt * prm.sigma_aw_c * prm.sigma_aw_c;
      for (int k = 0; k < 3; ++k) {
        c.P_delta[(3 + k) * 15 + 3 + k] += sigma2_dalpha;
        c.P_delta[(6 + k) * 15 + 6 + k] += sigma2_v;
        c.P_delta[(0 + k) * 15 + 0 + k] += sigma2_p;
        c.P_delta[(9 + k) * 15 + 9 + k] += sigma2_b_g;
        c.P_delta[(12 + k) * 15 + 12 + k] += sigma2_b_a;
      }
    }
    std::memcpy(c.Delta_q, Delta_q_1, sizeof Delta_q_1);
    std::memcpy(c.C_integral, C_integral_1, sizeof C_integral_1);
    std::memcpy(c.acc_integral, acc_integral_1, sizeof acc_integral_1);
    std::memcpy(c.cross, cross_1, sizeof cross_1);
    std::memcpy(c.dv_db_g, dv_db_g_1, sizeof dv_db_g_1);
    time = nexttime;
    ++i;
    if (nexttime == t1) break;
  }
  if (Delta_t_out) *Delta_t_out = Delta_t;
  return i;
}

// ImuError::redoPreintegration (ImuError.cpp:76-284)
inline int imu_redo_preintegration(const okb_imu_sample* s, int n, const okb_imu_params& prm, int64_t t0, int64_t t1,
                                   const double* sb, ImuCache& c) {
  const int i = imu_integrate(s, n, prm, t0, t1, sb, true, true, c, nullptr);
  if (i < 0) return i;
  std::memcpy(c.sb_ref, sb, 9 * sizeof(double));
  double Pt[225];
  transpose(c.P_delta, Pt, 15, 15);
  for (int k = 0; k < 225; ++k) c.P_delta[k] = 0.5 * c.P_delta[k] + 0.5 * Pt[k];
  inverse_lu(c.P_delta, c.information, 15);
  double It[225];
  transpose(c.information, It, 15, 15);
  for (int k = 0; k < 225; ++k) c.information[k] = 0.5 * c.information[k] + 0.5 * It[k];
  sqrt_information(c.information, c.squareRootInformation, 15);
  return i;
}

// ImuError::propagation (ImuError.cpp:287-504).  pose/sb in-out; covariance/jacobian 15x15 or null.
inline int imu_propagation(const okb_imu_sample* s, int n, const okb_imu_params& prm, double* pose, double* sb,
                           int64_t t0, int64_t t1, double* covariance, double* jacobian) {
  Transformation T_WS(pose);
  double r_0[3], q_WS_0[4], C_WS_0[9];
  std::memcpy(r_0, T_WS.r, 24); std::memcpy(q_WS_0, T_WS.q, 32); std::memcpy(C_WS_0, T_WS.C, 72);
  ImuCache c;
  double Delta_t = 0;
  const int i = imu_integrate(s, n, prm, t0, t1, sb, false, covariance != nullptr, c, &Delta_t);
  if (i < 0) return i;
  const double g_W[3] = {0, 0, prm.g};  // g * (0,0,6371009).normalized()
  double Cadd[3], rn[3], qn[4];
  matmul(C_WS_0, c.acc_doubleintegral, Cadd, 3, 3, 1);
  for (int k = 0; k < 3; ++k) rn[k] = r_0[k] + sb[k] * Delta_t + Cadd[k] - 0.5 * g_W[k] * Delta_t * Delta_t;
  qmul(q_WS_0, c.Delta_q, qn);
  T_WS.set(rn, qn);
  T_WS.to7(pose);
  double Cai[3];
  matmul(C_WS_0, c.acc_integral, Cai, 3, 3, 1);
  for (int k = 0; k < 3; ++k) sb[k] += Cai[k] - g_W[k] * Delta_t;
  if (jacobian) {
    double* F = jacobian;
    std::memset(F, 0, 225 * sizeof(double));
    for (int k = 0; k < 15; ++k) F[k * 15 + k] = 1.0;
    auto setblk = [&](int r0, int c0, const double* B, double sgn) {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) F[(r0 + a) * 15 + c0 + b] = sgn * B[a * 3 + b];
    };
    double X[9], M[9];
    crossMx(Cadd, X); setblk(0, 3, X, -1.0);
    double Idt[9] = {Delta_t, 0, 0, 0, Delta_t, 0, 0, 0, Delta_t}; setblk(0, 6, Idt, 1.0);
    matmul(C_WS_0, c.dp_db_g, M, 3, 3, 3); setblk(0, 9, M, 1.0);
    matmul(C_WS_0, c.C_doubleintegral, M, 3, 3, 3); setblk(0, 12, M, -1.0);
    matmul(C_WS_0, c.dalpha_db_g, M, 3, 3, 3); setblk(3, 9, M, -1.0);
    crossMx(Cai, X); setblk(6, 3, X, -1.0);
    matmul(C_WS_0, c.dv_db_g, M, 3, 3, 3); setblk(6, 9, M, 1.0);
    matmul(C_WS_0, c.C_integral, M, 3, 3, 3); setblk(6, 12, M, -1.0);
  }
  if (covariance) {
    double T[225], TP[225];
    std::memset(T, 0, sizeof T);
    for (int k = 0; k < 15; ++k) T[k * 15 + k] = 1.0;
    for (int b = 0; b < 3; ++b)
      for (int a = 0; a < 3; ++a)
        for (int d = 0; d < 3; ++d) T[(3 * b + a) * 15 + 3 * b + d] = C_WS_0[a * 3 + d];
    matmul(T, c.P_delta, TP, 15, 15, 15);
    matmul_nt(TP, T, covariance, 15, 15, 15);
  }
  return i;
}

// ImuError::EvaluateWithMinimalJacobians (ImuError.cpp:514-685).  Mutates the cache exactly like
// the reference (redo when redo_ || |Delta_b_g|*Delta_t > 1e-4, then Delta_b := 0).
// Minimal Jacobians row-major: J0 15x6, J1 15x9, J2 15x6, J3 15x9 (any may be null).
inline bool imu_error(const okb_imu_sample* s, int n, const okb_imu_params& prm, int64_t t0, int64_t t1,
                      const double* pose0, const double* sb0, const double* pose1, const double* sb1,
                      ImuCache& c, double* res, double* J0, double* J1, double* J2, double* J3) {
  const Transformation T_WS_0(pose0), T_WS_1(pose1);
  const double* C_WS_0 = T_WS_0.C;
  double C_S0_W[9];
  transpose(C_WS_0, C_S0_W, 3, 3);
  const double Delta_t = ns_to_sec(t1 - t0);
  double Delta_b[6];
  for (int k = 0; k < 6; ++k) Delta_b[k] = sb0[3 + k] - c.sb_ref[3 + k];
  c.redo = c.redo || (norm3(Delta_b) * Delta_t > 0.0001);
  if (c.redo) {
    imu_redo_preintegration(s, n, prm, t0, t1, sb0, c);
    c.redoCounter++;
    for (int k = 0; k < 6; ++k) Delta_b[k] = 0.0;
    c.redo = false;
  }
  const double g_W[3] = {0, 0, prm.g};
  double F0[225], F1[225];
  std::memset(F0, 0, sizeof F0); std::memset(F1, 0, sizeof F1);
  for (int k = 0; k < 15; ++k) { F0[k * 15 + k] = 1.0; F1[k * 15 + k] = -1.0; }
  double delta_p_est_W[3], delta_v_est_W[3];
  for (int k = 0; k < 3; ++k) {
    delta_p_est_W[k] = T_WS_0.r[k] - T_WS_1.r[k] + sb0[k] * Delta_t - 0.5 * g_W[k] * Delta_t * Delta_t;
    delta_v_est_W[k] = sb0[k] - sb1[k] - g_W[k] * Delta_t;
  }
  double mdb[3], dqb[4], Dq[4];
  {
    double t[3];
    matmul(c.dalpha_db_g, Delta_b, t, 3, 3, 1);
    mdb[0] = -t[0]; mdb[1] = -t[1]; mdb[2] = -t[2];
  }
  deltaQ(mdb, dqb);
  qmul(dqb, c.Delta_q, Dq);
  auto setblk = [&](double* F, int r0, int c0, const double* B, double sgn) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) F[(r0 + a) * 15 + c0 + b] = sgn * B[a * 3 + b];
  };
  double X[9], M[9];
  setblk(F0, 0, 0, C_S0_W, 1.0);
  crossMx(delta_p_est_W, X); matmul(C_S0_W, X, M, 3, 3, 3); setblk(F0, 0, 3, M, 1.0);
  for (int k = 0; k < 9; ++k) M[k] = C_S0_W[k] * Delta_t;
  setblk(F0, 0, 6, M, 1.0);
  setblk(F0, 0, 9, c.dp_db_g, 1.0);
  setblk(F0, 0, 12, c.C_doubleintegral, -1.0);
  double q1inv[4], Dq_q1inv[4], Qp[16], Qo[16], Q44[16];
  qinv(T_WS_1.q, q1inv);
  qmul(Dq, q1inv, Dq_q1inv);
  qplusMat(Dq_q1inv, Qp); qoplusMat(T_WS_0.q, Qo);
  matmul(Qp, Qo, Q44, 4, 4, 4);
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) F0[(3 + a) * 15 + 3 + b] = Q44[a * 4 + b];
  double q1inv_q0[4], Qa[16], Qb[16];
  qmul(q1inv, T_WS_0.q, q1inv_q0);
  qoplusMat(q1inv_q0, Qa); qoplusMat(Dq, Qb);
  matmul(Qa, Qb, Q44, 4, 4, 4);
  {
    double T33[9], R[9];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) T33[a * 3 + b] = Q44[a * 4 + b];
    matmul(T33, c.dalpha_db_g, R, 3, 3, 3);
    setblk(F0, 3, 9, R, -1.0);
  }
  crossMx(delta_v_est_W, X); matmul(C_S0_W, X, M, 3, 3, 3); setblk(F0, 6, 3, M, 1.0);
  setblk(F0, 6, 6, C_S0_W, 1.0);
  setblk(F0, 6, 9, c.dv_db_g, 1.0);
  setblk(F0, 6, 12, c.C_integral, -1.0);

  setblk(F1, 0, 0, C_S0_W, -1.0);
  {
    double QpDq[16], Qo0[16], Qp1[16], T1[16], T2[16];
    qplusMat(Dq, QpDq); qoplusMat(T_WS_0.q, Qo0); qplusMat(q1inv, Qp1);
    matmul(QpDq, Qo0, T1, 4, 4, 4);
    matmul(T1, Qp1, T2, 4, 4, 4);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) F1[(3 + a) * 15 + 3 + b] = -T2[a * 4 + b];
  }
  setblk(F1, 6, 6, C_S0_W, -1.0);

  double error[15];
  {
    double t[3], fb[3];
    matmul(C_S0_W, delta_p_est_W, t, 3, 3, 1);
    for (int a = 0; a < 3; ++a) {
      fb[a] = 0;
      for (int b = 0; b < 6; ++b) fb[a] += F0[(0 + a) * 15 + 9 + b] * Delta_b[b];
      error[a] = t[a] + c.acc_doubleintegral[a] + fb[a];
    }
    double qe[4];
    qmul(Dq, q1inv_q0, qe);
    error[3] = 2 * qe[0]; error[4] = 2 * qe[1]; error[5] = 2 * qe[2];
    matmul(C_S0_W, delta_v_est_W, t, 3, 3, 1);
    for (int a = 0; a < 3; ++a) {
      fb[a] = 0;
      for (int b = 0; b < 6; ++b) fb[a] += F0[(6 + a) * 15 + 9 + b] * Delta_b[b];
      error[6 + a] = t[a] + c.acc_integral[a] + fb[a];
    }
    for (int a = 0; a < 6; ++a) error[9 + a] = sb0[3 + a] - sb1[3 + a];
  }
  matmul(c.squareRootInformation, error, res, 15, 15, 1);
  auto wblock = [&](const double* F, int c0, int w, double* J) {
    if (!J) return;
    for (int r = 0; r < 15; ++r)
      for (int cc = 0; cc < w; ++cc) {
        double sum = 0;
        for (int k = 0; k < 15; ++k) sum += c.squareRootInformation[r * 15 + k] * F[k * 15 + c0 + cc];
        J[r * w + cc] = sum;
      }
  };
  wblock(F0, 0, 6, J0); wblock(F0, 6, 9, J1); wblock(F1, 0, 6, J2); wblock(F1, 6, 9, J3);
  return true;
}

// ---- priors ------------------------------------------------------------------------------
// SpeedAndBiasError (okvis_ceres/src/SpeedAndBiasError.cpp:89-116)
inline void speed_bias_error(const double* meas, const double* sqrtInfo, const double* sb, double* res, double* J) {
  double e[9];
  for (int k = 0; k < 9; ++k) e[k] = meas[k] - sb[k];
  matmul(sqrtInfo, e, res, 9, 9, 1);
  if (J) for (int k = 0; k < 81; ++k) J[k] = -sqrtInfo[k];
}
// PoseError (okvis_ceres/src/PoseError.cpp:86-136); J minimal 6x6
inline void pose_error(const double* meas7, const double* sqrtInfo, const double* pose, double* res, double* J) {
  const Transformation T_WS(pose), Tm(meas7);
  const Transformation dp = Tm * T_WS.inverse();
  double e[6];
  for (int k = 0; k < 3; ++k) { e[k] = Tm.r[k] - T_WS.r[k]; e[3 + k] = 2 * dp.q[k]; }
  matmul(sqrtInfo, e, res, 6, 6, 1);
  if (J) {
    double Jm[36], Qp[16];
    std::memset(Jm, 0, sizeof Jm);
    for (int k = 0; k < 3; ++k) Jm[k * 6 + k] = -1.0;
    qplusMat(dp.q, Qp);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Jm[(3 + a) * 6 + 3 + b] = -Qp[a * 4 + b];
    matmul(sqrtInfo, Jm, J, 6, 6, 6);
  }
}
// RelativePoseError (okvis_ceres/src/RelativePoseError.cpp:84-162); J0, J1 minimal 6x6
inline void relative_pose_error(const double* sqrtInfo, const double* pose0, const double* pose1, double* res,
                                double* J0, double* J1) {
  const Transformation T0(pose0), T1(pose1);
  const Transformation dp = T1 * T0.inverse();
  double e[6];
  for (int k = 0; k < 3; ++k) { e[k] = T1.r[k] - T0.r[k]; e[3 + k] = 2 * dp.q[k]; }
  matmul(sqrtInfo, e, res, 6, 6, 1);
  if (J0) {
    double Jm[36], Qp[16];
    std::memset(Jm, 0, sizeof Jm);
    for (int k = 0; k < 3; ++k) Jm[k * 6 + k] = -1.0;
    qplusMat(dp.q, Qp);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Jm[(3 + a) * 6 + 3 + b] = -Qp[a * 4 + b];
    matmul(sqrtInfo, Jm, J0, 6, 6, 6);
  }
  if (J1) {
    double Jm[36], Qo[16];
    std::memset(Jm, 0, sizeof Jm);
    for (int k = 0; k < 3; ++k) Jm[k * 6 + k] = 1.0;
    qoplusMat(dp.q, Qo);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Jm[(3 + a) * 6 + 3 + b] = Qo[a * 4 + b];
    matmul(sqrtInfo, Jm, J1, 6, 6, 6);
  }
}

// ---- MarginalizationError ------------------------------------------------------------------
inline int block_dim(int kind) { return kind == OKB_BLOCK_SPEED_BIAS ? 9 : 7; }
inline int block_min_dim(int kind) { return kind == OKB_BLOCK_SPEED_BIAS ? 9 : 6; }

// MarginalizationError::EvaluateWithMinimalJacobians (okvis_ceres/src/MarginalizationError.cpp:893-946)
// + computeDeltaChi (:867-882).  `x` = current values of the connected blocks, concatenated like x0.
// `fixed[i]` blocks contribute no columns.  Outputs: res[n] and J_eff (n x n_cols row-major), the
// minimal Jacobian Ceres ends up with: J[:,idx_i] * liftJacobian(x0_i) * plusJacobian(x_i).
inline void marginalization_error(const okb_marg_prior& m, const uint8_t* fixed, const double* x, double* res,
                                  double* J_eff) {
  const int n = m.n;
  std::vector<double> dchi(n, 0.0);
  int col = 0, off = 0;
  for (int i = 0; i < m.n_blocks; ++i) {
    const int kind = m.block_kind[i];
    const int dim = block_dim(kind), md = block_min_dim(kind);
    if (fixed && fixed[i]) { off += dim; continue; }
    if (kind == OKB_BLOCK_SPEED_BIAS) {
      for (int k = 0; k < 9; ++k) dchi[col + k] = x[off + k] - m.x0[off + k];
    } else {
      pose_minus(m.x0 + off, x + off, &dchi[col]);
    }
    col += md; off += dim;
  }
  for (int r = 0; r < n; ++r) {
    double s = m.e0[r];
    for (int c2 = 0; c2 < n; ++c2) s += m.J[r * n + c2] * dchi[c2];
    res[r] = s;
  }
  if (J_eff) {
    col = 0; off = 0;
    for (int i = 0; i < m.n_blocks; ++i) {
      const int kind = m.block_kind[i];
      const int dim = block_dim(kind), md = block_min_dim(kind);
      if (fixed && fixed[i]) { off += dim; continue; }
      if (kind == OKB_BLOCK_SPEED_BIAS) {
        for (int r = 0; r < n; ++r)
          for (int k = 0; k < 9; ++k) J_eff[r * n + col + k] = m.J[r * n + col + k];
      } else {
        double lift[42], plus[42], LP[36];
        pose_lift_jacobian(m.x0 + off, lift);
        pose_plus_jacobian(x + off, plus);
        matmul(lift, plus, LP, 6, 7, 6);
        for (int r = 0; r < n; ++r)
          for (int k = 0; k < 6; ++k) {
            double s = 0;
            for (int p = 0; p < 6; ++p) s += m.J[r * n + col + p] * LP[p * 6 + k];
            J_eff[r * n + col + k] = s;
          }
      }
      col += md; off += dim;
    }
  }
}

}  // namespace oko
