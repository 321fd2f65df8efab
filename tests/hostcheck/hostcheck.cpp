// TEST INFRASTRUCTURE ONLY -- compiles the product's host/device math headers (okb_math.cuh,
// okb_imu.cuh) for the HOST so that tests can compare them with the oracle without a GPU.
// This library is never loaded by the product; the product path is CUDA only.
#include <cstring>
#include "../../okvis_b200/csrc/okb_imu.cuh"
using namespace okb;
extern "C" {
void hc_reproj_full(const okb_camera* cam, const double* pose, const double* X, const double* ext, const double* z,
                    double sqrt_info, double* r, double* J0, double* J1, double* J2) {
  reproj_full(*cam, pose, X, ext, z, sqrt_info, r, J0, J1, J2);
}
void hc_pose_error(const double* meas, const double* S, const double* pose, double* r, double* J) { pose_error(meas, S, pose, r, J); }
void hc_relative_pose(const double* S, const double* p0, const double* p1, double* r, double* J0, double* J1) {
  relative_pose_error(S, p0, p1, r, J0, J1);
}
void hc_pose_plus(const double* x, const double* d, double* o) { pose_plus(x, d, o); }
void hc_pose_minus(const double* x, const double* xpd, double* d) { pose_minus(x, xpd, d); }
void hc_marg_rot_block(const double* x0, const double* x, double* B) { marg_pose_rot_block(x0, x, B); }
void hc_eig3(const double* S6, double* ev) { eig3sym(S6, ev); }
int hc_imu_eval(const okb_imu_params* prm, const okb_imu_sample* s, int n, int64_t t0, int64_t t1, const double* pose0,
                const double* sb0, const double* pose1, const double* sb1, const double* sb_ref, double* r, double* J0,
                double* J1, double* J2, double* J3, double* sqrt_info) {
  SeqCtx cx;
  static ImuCache cache;
  std::memset(&cache, 0, sizeof cache);
  double P[225], F[225], T[225], P2[225], Sb[32 * kImuPre], F01[450], SF[450], e[15];
  ImuWork wk{P, F, T, P2, Sb};
  if (sb_ref) imu_preintegrate(cx, s, n, *prm, t0, t1, sb_ref, &cache, wk);
  const int before = cache.redo_count;
  imu_evaluate(cx, s, n, *prm, t0, t1, pose0, sb0, pose1, sb1, &cache, wk, F01, e, r, SF);
  for (int rr = 0; rr < 15; ++rr) {
    for (int c = 0; c < 6; ++c) { J0[rr * 6 + c] = SF[rr * 30 + c]; J2[rr * 6 + c] = SF[rr * 30 + 15 + c]; }
    for (int c = 0; c < 9; ++c) { J1[rr * 9 + c] = SF[rr * 30 + 6 + c]; J3[rr * 9 + c] = SF[rr * 30 + 21 + c]; }
  }
  if (sqrt_info) std::memcpy(sqrt_info, cache.sqrt_info, sizeof cache.sqrt_info);
  return cache.redo_count - before;
}
int hc_imu_propagate(const okb_imu_params* prm, const okb_imu_sample* s, int n, int64_t t0, int64_t t1, double* pose,
                     double* sb, double* cov, double* jac) {
  SeqCtx cx;
  double P[225], F[225], T[225], P2[225], Sb[32 * kImuPre];
  ImuWork wk{P, F, T, P2, Sb};
  return imu_propagate(cx, s, n, *prm, t0, t1, pose, sb, cov, jac, wk);
}
}


// ---- host packing logic (okb_hostpack.hpp)
#include "../../okvis_b200/csrc/okb_hostpack.hpp"
extern "C" void hc_sort_landmarks(const uint32_t* vis, int L, uint32_t* perm, uint32_t* inv, uint32_t* tile_range) {
  okb::sort_landmarks_by_frame_range(vis, L, perm, inv, tile_range);
}
