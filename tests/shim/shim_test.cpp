// Exercises the C++ host shim (include/okvis_b200_estimator.hpp) the way ThreadedKFVio drives
// okvis::Estimator: addCamera/addImu, then per frame addStates -> addLandmark/addObservation -> optimize.
// Prints one JSON line; tests/test_gpu_shim.py checks it.  Usage: shim_test.bin [compile-only smoke: --no-gpu]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "okvis_b200_estimator.hpp"

using namespace okvis_b200;

int main(int argc, char** argv) {
  if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
    try { Estimator e(0); std::printf("{\"created\": true}\n"); }
    catch (const std::exception& ex) { std::printf("{\"created\": false, \"error\": \"%s\"}\n", ex.what()); }
    return 0;
  }
  const double g = 9.81007;
  Estimator est(0);
  okb_camera cam{};
  cam.model = OKB_DIST_NONE; cam.width = 752; cam.height = 480; cam.fu = 450; cam.fv = 450; cam.cu = 376; cam.cv = 240;
  // camera z = S x (forward), camera x = -S y, camera y = -S z :  C_SC = [[0,0,1],[-1,0,0],[0,-1,0]]
  // as a quaternion (x,y,z,w) = (-0.5, 0.5, -0.5, 0.5)
  const Pose7 T_SC0{{0.0, 0.055, 0.0, -0.5, 0.5, -0.5, 0.5}}, T_SC1{{0.0, -0.055, 0.0, -0.5, 0.5, -0.5, 0.5}};
  est.addCamera(ExtrinsicsEstimationParameters(), cam, T_SC0);
  est.addCamera(ExtrinsicsEstimationParameters(), cam, T_SC1);
  okb_imu_params imu{};
  imu.a_max = 176; imu.g_max = 7.8; imu.sigma_g_c = 12e-4; imu.sigma_a_c = 8e-3; imu.sigma_bg = 0.03; imu.sigma_ba = 0.1;
  imu.sigma_gw_c = 4e-6; imu.sigma_aw_c = 4e-5; imu.tau = 3600; imu.g = g; imu.rate = 200;
  est.addImu(imu);

  std::mt19937 rng(7);
  std::uniform_real_distribution<double> ux(3.0, 9.0), uy(-2.5, 3.5), uz(-1.5, 1.5);
  std::normal_distribution<double> noise(0.0, 0.05);
  const int L = 120, K = 5;
  double lmk[L][3];
  for (int l = 0; l < L; ++l) { lmk[l][0] = ux(rng); lmk[l][1] = uy(rng); lmk[l][2] = uz(rng); }
  const double vy = 0.5, dtf = 0.2;
  auto project = [&](const double* p_W, double y_cam, double ybase, double* px) {
    // S = W translated by (0, y_cam, 0); camera centre at S + (0, ybase, 0)
    const double xs = p_W[0], ys = p_W[1] - y_cam - ybase, zs = p_W[2];
    const double xc = -ys, yc = -zs, zc = xs;
    px[0] = cam.fu * xc / zc + cam.cu; px[1] = cam.fv * yc / zc + cam.cv;
    return zc > 0.5 && px[0] >= 0 && px[0] < 752 && px[1] >= 0 && px[1] < 480;
  };
  double final_err = 0;
  for (int k = 0; k < K; ++k) {
    const int64_t t_k = 1000000000LL + (int64_t)(k * dtf * 1e9);
    std::vector<ImuMeasurement> meas;
    const int64_t t_prev = 1000000000LL + (int64_t)((k - 1) * dtf * 1e9);
    for (int64_t t = (k == 0 ? t_k - 50000000LL : t_prev - 5000000LL); t <= t_k + 5000000LL; t += 5000000LL) {
      ImuMeasurement m{}; m.t_ns = t; m.acc[2] = g; meas.push_back(m);
    }
    if (!est.addStates(100 + k, t_k, meas, true)) { std::printf("{\"error\": \"addStates failed at %d\"}\n", k); return 1; }
    for (int l = 0; l < L; ++l) {
      for (int c = 0; c < 2; ++c) {
        double px[2];
        if (!project(lmk[l], vy * k * dtf, c == 0 ? 0.055 : -0.055, px)) continue;
        if (!est.isLandmarkAdded(1000 + l)) {
          Vec4 hp{{lmk[l][0] + noise(rng), lmk[l][1] + noise(rng), lmk[l][2] + noise(rng), 1.0}};
          est.addLandmark(1000 + l, hp);
        }
        est.addObservation(1000 + l, 100 + k, c, l, px, 8.0);
      }
    }
    est.optimize(10, 2, false);
    Pose7 T;
    est.get_T_WS(100 + k, T);
    final_err = std::sqrt(T[0] * T[0] + (T[1] - vy * k * dtf) * (T[1] - vy * k * dtf) + T[2] * T[2]);
  }
  SpeedAndBias sb;
  est.getSpeedAndBias(100 + K - 1, 0, sb);
  MapPoint mp;
  est.getLandmark(1000, mp);
  std::printf("{\"frames\": %zu, \"landmarks\": %zu, \"final_cost\": %.6e, \"initial_cost\": %.6e, \"iterations\": %d, \"pos_err\": %.6e, "
              "\"vy\": %.6f, \"lm_quality\": %.4f, \"dup_obs\": %llu}\n",
              est.numFrames(), est.numLandmarks(), est.summary().final_cost, est.summary().initial_cost, est.summary().iterations, final_err, sb[1],
              mp.quality, (unsigned long long)est.addObservation(1000, 100, 0, 0, mp.point.data(), 8.0));
  return 0;
}
