// TEST INFRASTRUCTURE ONLY -- CPU restatement of the geometric match gate of OKVIS' production matching algorithm:
//   VioKeyframeWindowMatchingAlgorithm::verifyMatch        okvis_frontend/src/VioKeyframeWindowMatchingAlgorithm.cpp:304-339
//   ProbabilisticStereoTriangulator::stereoTriangulate     okvis_frontend/src/ProbabilisticStereoTriangulator.cpp:168-227
//   ProbabilisticStereoTriangulator::computeReprojectionError4   :359-384
//   triangulation::triangulateFast                         okvis_frontend/src/stereo_triangulation.cpp:51-125
//   PinholeCamera::projectHomogeneous status               okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:147-226, 345-378,
//                                                          CameraBase::isInImage  implementation/CameraBase.hpp:95-104
// PARITY UNPINNED: the reference has no test with asserted values for these functions (TestFrontend does not exist);
// the restatement follows the source line by line.
#pragma once
#include <algorithm>
#include <cmath>

#include "../include/okvis_b200.h"
#include "oracle_errors.hpp"

namespace oko {

// returns the homogeneous point (unit 4-vector) and sets isValid / isParallel
inline void triangulate_fast(const double* p1, const double* e1, const double* p2, const double* e2, double sigma, double* hp, bool& isValid,
                             bool& isParallel) {
  isParallel = false;
  isValid = false;
  const double t12[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double b0 = dot(t12, e1), b1 = dot(t12, e2);
  double A00 = dot(e1, e1), A10 = dot(e1, e2), A01 = -A10, A11 = -dot(e2, e2);
  if (A10 < 0.0) { A10 = -A10; A01 = -A01; }
  const double det = A00 * A11 - A01 * A10;
  const bool invertible = std::fabs(det) > 1.0e-6;            // computeInverseWithCheck(..., 1.0e-6) of a fixed 2x2
  if (!invertible) {
    isParallel = true;
    const double c[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    if (std::sqrt(dot(c, c)) < 6 * sigma) isValid = true;
    double v[4] = {(e1[0] + e2[0]) / 2.0, (e1[1] + e2[1]) / 2.0, (e1[2] + e2[2]) / 2.0, 1e-3};
    const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    for (int k = 0; k < 4; ++k) hp[k] = v[k] / n;
    return;
  }
  const double i00 = A11 / det, i01 = -A01 / det, i10 = -A10 / det, i11 = A00 / det;
  const double l0 = i00 * b0 + i01 * b1, l1 = i10 * b0 + i11 * b1;
  double xm[3], xn[3], mid[3], err[3], diff[3];
  for (int k = 0; k < 3; ++k) { xm[k] = l0 * e1[k] + p1[k]; xn[k] = l1 * e2[k] + p2[k]; mid[k] = (xm[k] + xn[k]) / 2.0; }
  for (int k = 0; k < 3; ++k) { err[k] = mid[k] - xm[k]; diff[k] = mid[k] - (p1[k] + 0.5 * t12[k]); }
  const double diff_sq = dot(diff, diff);
  const double chi2 = dot(err, err) * (1.0 / (diff_sq * sigma * sigma));
  isValid = true;
  if (chi2 > 9) isValid = false;
  if (dot(diff, e1) < 0)
    for (int k = 0; k < 3; ++k) mid[k] = (p1[k] + 0.5 * t12[k]) - diff[k];
  const double n = std::sqrt(mid[0] * mid[0] + mid[1] * mid[1] + mid[2] * mid[2] + 1.0);
  hp[0] = mid[0] / n; hp[1] = mid[1] / n; hp[2] = mid[2] / n; hp[3] = 1.0 / n;
}

// projectHomogeneous(...) == Successful ?  (image point in y)
inline bool project_successful(const okb_camera& cam, const double* hp, double* y) {
  double head[3] = {hp[0], hp[1], hp[2]};
  if (hp[3] < 0) { head[0] = -head[0]; head[1] = -head[1]; head[2] = -head[2]; }
  if (!project(cam, head, y, nullptr)) return false;                                   // Invalid
  if (y[0] < 0.0 || y[1] < 0.0 || y[0] >= cam.width || y[1] >= cam.height) return false;   // OutsideImage
  return head[2] > 0.0;                                                               // else Behind
}

inline bool reprojection_error4(const okb_camera& cam, const double* kp, double size, const double* hp, double& err) {
  double y[2];
  if (!project_successful(cam, hp, y)) return false;
  const double sd = 0.8 * size / 12.0;
  const double w = 1.0 / (sd * sd);
  const double d0 = y[0] - kp[0], d1 = y[1] - kp[1];
  err = d0 * (w * d0) + d1 * (w * d1);
  return true;
}

inline bool verify_match(const okb_match_gate& g, int a, int b) {
  if (g.mode == OKB_GATE_2D2D) {
    Transformation T_AB(g.T_AB);
    auto unit = [](const double* v, double* o) { const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); o[0] = v[0] / n; o[1] = v[1] / n; o[2] = v[2] / n; };
    double e1[3], eb[3], e2[3];
    unit(g.bearing_a + 3 * a, e1);
    matmul(T_AB.C, g.bearing_b + 3 * b, eb, 3, 3, 1);
    unit(eb, e2);
    const double p1[3] = {0, 0, 0};
    double hpA[4];
    bool isValid, isParallel;
    triangulate_fast(p1, e1, T_AB.r, e2, std::max(g.ray_sigma_a[a], g.ray_sigma_b[b]), hpA, isValid, isParallel);
    if (!isValid) return false;
    double errA, errB;
    if (!reprojection_error4(g.cam_a, g.kp_a + 2 * a, g.kp_size_a[a], hpA, errA)) return false;
    // T_BA * hpA
    double Ct[9];
    transpose(T_AB.C, Ct, 3, 3);
    double hpB[4];
    for (int k = 0; k < 3; ++k) {
      const double tk = -(Ct[k * 3] * T_AB.r[0] + Ct[k * 3 + 1] * T_AB.r[1] + Ct[k * 3 + 2] * T_AB.r[2]);
      hpB[k] = Ct[k * 3] * hpA[0] + Ct[k * 3 + 1] * hpA[1] + Ct[k * 3 + 2] * hpA[2] + tk * hpA[3];
    }
    hpB[3] = hpA[3];
    if (!reprojection_error4(g.cam_b, g.kp_b + 2 * b, g.kp_size_b[b], hpB, errB)) return false;
    return !(errA > 4.0 || errB > 4.0);
  }
  // 3D-2D: chi2 of the projection of landmark a into B against keypoint b
  double sd = 0.8 * g.kp_size_b[b] / 12.0;
  const double* P = g.proj_uncertainty + 4 * a;
  const double U00 = sd * sd + P[0], U01 = P[1], U10 = P[2], U11 = sd * sd + P[3];
  const double det = U00 * U11 - U01 * U10;
  const double e0 = g.proj_into_b[2 * a] - g.kp_b[2 * b], e1 = g.proj_into_b[2 * a + 1] - g.kp_b[2 * b + 1];
  const double v0 = (U11 * e0 - U01 * e1) / det, v1 = (-U10 * e0 + U00 * e1) / det;    // U^-1 err
  const int chi2 = (int)(e0 * v0 + e1 * v1);        // `const int chi2 = ...` in the reference
  return chi2 < 4.0;
}

}  // namespace oko
