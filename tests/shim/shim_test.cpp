// Exercises the C++ host shim (include/okvis_b200_estimator.hpp) the way ThreadedKFVio drives
// okvis::Estimator: addCamera/addImu, then per frame addStates -> addLandmark/addObservation -> optimize.
// Prints one JSON line; tests/test_gpu_shim.py checks it.  Usage: shim_test.bin [compile-only smoke: --no-gpu]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>
#include <set>

#include "okvis_b200_estimator.hpp"
#include "okvis_b200_frontend.hpp"

using namespace okvis_b200;

static std::set<uint64_t> g_removed;
static bool l_was_removed(uint64_t id) { return g_removed.count(id) != 0; }
static void mark_removed(uint64_t id) { g_removed.insert(id); }

// The ThreadedKFVio cycle (okvis_multisensor_processing/src/ThreadedKFVio.cpp:733-770): per frame addStates ->
// addLandmark / addObservation -> optimize -> applyMarginalizationStrategy(numKeyframes = 5, numImuFrames = 3).
// Constant-velocity motion along +y past a field of landmarks, exact measurements; every third frame is a keyframe.
static int sliding(int n_frames) {
  const double g = 9.81007;
  Estimator est(0);
  okb_camera cam{};
  cam.model = OKB_DIST_NONE; cam.width = 752; cam.height = 480; cam.fu = 450; cam.fv = 450; cam.cu = 376; cam.cv = 240;
  const Pose7 T_SC0{{0.0, 0.055, 0.0, -0.5, 0.5, -0.5, 0.5}}, T_SC1{{0.0, -0.055, 0.0, -0.5, 0.5, -0.5, 0.5}};
  est.addCamera(ExtrinsicsEstimationParameters(), cam, T_SC0);
  est.addCamera(ExtrinsicsEstimationParameters(), cam, T_SC1);
  okb_imu_params imu{};
  imu.a_max = 176; imu.g_max = 7.8; imu.sigma_g_c = 12e-4; imu.sigma_a_c = 8e-3; imu.sigma_bg = 0.03; imu.sigma_ba = 0.1;
  imu.sigma_gw_c = 4e-6; imu.sigma_aw_c = 4e-5; imu.tau = 3600; imu.g = g; imu.rate = 200;
  est.addImu(imu);
  std::mt19937 rng(11);
  const int L = 600;
  const double vy = 1.5, dtf = 0.2;
  std::uniform_real_distribution<double> ux(3.0, 9.0), uy(-4.0, 4.0 + vy * dtf * n_frames), uz(-1.5, 1.5);
  std::normal_distribution<double> noise(0.0, 0.05);
  std::vector<std::array<double, 3>> lmk(L);
  for (auto& p : lmk) p = {ux(rng), uy(rng), uz(rng)};
  auto project = [&](const double* p_W, double y_cam, double ybase, double* px) {
    const double xs = p_W[0], ys = p_W[1] - y_cam - ybase, zs = p_W[2];
    const double xc = -ys, yc = -zs, zc = xs;
    px[0] = cam.fu * xc / zc + cam.cu; px[1] = cam.fv * yc / zc + cam.cv;
    return zc > 0.5 && px[0] >= 0 && px[0] < 752 && px[1] >= 0 && px[1] < 480;
  };
  size_t max_frames = 0, removed_total = 0, max_landmarks = 0;
  double max_err = 0, last_err = 0, max_cost = 0;
  long long first_bytes = 0, steady_bytes = 0;
  int marg_calls = 0, imu_window_ok = 1;
  for (int k = 0; k < n_frames; ++k) {
    const int64_t t_k = 1000000000LL + (int64_t)(k * dtf * 1e9);
    const int64_t t_prev = 1000000000LL + (int64_t)((k - 1) * dtf * 1e9);
    std::vector<ImuMeasurement> meas;
    for (int64_t t = (k == 0 ? t_k - 50000000LL : t_prev - 5000000LL); t <= t_k + 5000000LL; t += 5000000LL) {
      ImuMeasurement m{}; m.t_ns = t; m.acc[2] = g; meas.push_back(m);
    }
    const uint64_t fid = 100 + k;
    if (!est.addStates(fid, t_k, meas, k % 3 == 0)) { std::printf("{\"error\": \"addStates failed at %d\"}\n", k); return 1; }
    for (int l = 0; l < L; ++l)
      for (int c = 0; c < 2; ++c) {
        double px[2];
        if (!project(lmk[l].data(), vy * k * dtf, c == 0 ? 0.055 : -0.055, px)) continue;
        const uint64_t lid = 100000 + l;
        if (!est.isLandmarkAdded(lid)) {
          MapPoint gone;
          if (l_was_removed(lid)) continue;       // a marginalised landmark does not come back (the frontend would create a new id)
          Vec4 hp{{lmk[l][0] + noise(rng), lmk[l][1] + noise(rng), lmk[l][2] + noise(rng), 1.0}};
          est.addLandmark(lid, hp);
        }
        est.addObservation(lid, fid, c, l, px, 8.0);
      }
    est.optimize(10, 2, false);
    if (k == 0) first_bytes = est.lastUploadBytes(); else steady_bytes = est.lastUploadBytes();
    max_cost = std::max(max_cost, est.summary().final_cost);
    MapPointVector removed;
    if (!est.applyMarginalizationStrategy(5, 3, removed)) { std::printf("{\"error\": \"marginalization failed at %d\"}\n", k); return 1; }
    for (auto& mp : removed) mark_removed(mp.id);
    removed_total += removed.size();
    ++marg_calls;
    max_frames = std::max(max_frames, est.numFrames());
    max_landmarks = std::max(max_landmarks, est.numLandmarks());
    Pose7 T;
    est.get_T_WS(fid, T);
    last_err = std::sqrt(T[0] * T[0] + (T[1] - vy * k * dtf) * (T[1] - vy * k * dtf) + T[2] * T[2]);
    max_err = std::max(max_err, last_err);
    // the newest three frames keep their speed/bias blocks, older ones do not (Estimator.cpp:483-554)
    for (size_t age = 0; age < est.numFrames(); ++age) {
      const bool in = est.isInImuWindow(est.frameIdByAge(age));
      if ((age < 3) != in && est.numFrames() > 3) imu_window_ok = 0;
    }
  }
  SpeedAndBias sb;
  est.getSpeedAndBias(est.currentFrameId(), 0, sb);
  std::printf("{\"frames\": %zu, \"max_frames\": %zu, \"landmarks\": %zu, \"max_landmarks\": %zu, \"removed_landmarks\": %zu, \"marg_calls\": %d, "
              "\"pos_err\": %.6e, \"max_pos_err\": %.6e, \"vy\": %.6f, \"final_cost\": %.6e, \"max_cost\": %.6e, \"first_upload_bytes\": %lld, "
              "\"steady_upload_bytes\": %lld, \"imu_window_ok\": %d, \"iterations\": %d}\n",
              est.numFrames(), max_frames, est.numLandmarks(), max_landmarks, removed_total, marg_calls, last_err, max_err, sb[1], est.summary().final_cost,
              max_cost, first_bytes, steady_bytes, imu_window_ok, est.summary().iterations);
  return 0;
}

// okvis_b200::Frontend + okvis_b200::DenseMatcher the way Frontend::detectAndDescribe / matchStereo use their
// reference counterparts: detect + describe two views of the same texture (12 px disparity), match with the plain
// Hamming MatchingAlgorithm (okvis_matcher/include/okvis/MatchingAlgorithm.hpp interface).
struct HammingAlgorithm {
  const std::vector<uint8_t>*dA, *dB;
  int bytes;
  std::vector<std::array<int, 2>> matches;
  std::vector<float> distances;
  void doSetup() { matches.clear(); distances.clear(); }
  size_t sizeA() const { return dA->size() / bytes; }
  size_t sizeB() const { return dB->size() / bytes; }
  bool skipA(size_t) const { return false; }
  bool skipB(size_t) const { return false; }
  float distanceThreshold() const { return 60.0f; }
  float distanceRatioThreshold() const { return 3.0f; }
  void reserveMatches(size_t n) { matches.reserve(n); }
  void setBestMatch(size_t a, size_t b, double d) { matches.push_back({(int)a, (int)b}); distances.push_back((float)d); }
  const uint8_t* descriptorsA() const { return dA->data(); }
  const uint8_t* descriptorsB() const { return dB->data(); }
  int descriptorBytes() const { return bytes; }
};

// The geometry-gated variant (what VioKeyframeWindowMatchingAlgorithm<...> is for matchStereo): the 2D-2D gate of a
// rectified pair.  doSetup() of the reference back-projects the keypoints and derives the ray sigmas
// (VioKeyframeWindowMatchingAlgorithm.cpp:236-276); here the same quantities for an undistorted pinhole.
struct GatedStereoAlgorithm : HammingAlgorithm {
  const std::vector<okb_keypoint>*kA, *kB;
  okb_camera cam;
  double baseline_x;
  std::vector<double> xyA, xyB, szA, szB, rayA, rayB, sgA, sgB;
  okb_match_gate gate;
  void doSetup() {
    HammingAlgorithm::doSetup();
    auto fill = [&](const std::vector<okb_keypoint>& k, std::vector<double>& xy, std::vector<double>& sz, std::vector<double>& ray, std::vector<double>& sg) {
      xy.clear(); sz.clear(); ray.clear(); sg.clear();
      for (const auto& p : k) {
        xy.push_back(p.x); xy.push_back(p.y); sz.push_back(p.size);
        ray.push_back((p.x - cam.cu) / cam.fu); ray.push_back((p.y - cam.cv) / cam.fv); ray.push_back(1.0);
        sg.push_back(std::sqrt(std::sqrt(2.0)) * (0.8 * p.size / 12.0) / cam.fu);
      }
    };
    fill(*kA, xyA, szA, rayA, sgA);
    fill(*kB, xyB, szB, rayB, sgB);
    gate = okb_match_gate{};
    gate.mode = OKB_GATE_2D2D;
    gate.kp_a = xyA.data(); gate.kp_size_a = szA.data(); gate.bearing_a = rayA.data(); gate.ray_sigma_a = sgA.data();
    gate.kp_b = xyB.data(); gate.kp_size_b = szB.data(); gate.bearing_b = rayB.data(); gate.ray_sigma_b = sgB.data();
    gate.cam_a = cam; gate.cam_b = cam;
    gate.T_AB[0] = baseline_x; gate.T_AB[6] = 1.0;
  }
  const okb_match_gate& matchGate() const { return gate; }
};

static int frontend_demo() {
  const int W = 752, H = 480, shift = 12;
  std::vector<uint8_t> base((size_t)(W + shift) * H), left((size_t)W * H), right((size_t)W * H);
  std::mt19937 rng(3);
  // blocky random texture with corners: 16 px cells of random brightness, smoothed by a 2 px ramp
  std::uniform_int_distribution<int> u(20, 235);
  std::vector<int> cell((size_t)((W + shift) / 16 + 2) * (H / 16 + 2));
  for (auto& c : cell) c = u(rng);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W + shift; ++x) base[(size_t)y * (W + shift) + x] = (uint8_t)cell[(size_t)(y / 16) * ((W + shift) / 16 + 2) + x / 16];
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) { left[(size_t)y * W + x] = base[(size_t)y * (W + shift) + x + shift]; right[(size_t)y * W + x] = base[(size_t)y * (W + shift) + x]; }
  okb_camera cam{};
  cam.model = OKB_DIST_NONE; cam.width = W; cam.height = H; cam.fu = 450; cam.fv = 450; cam.cu = 376; cam.cv = 240;
  Frontend fe(2);
  fe.setBriskDetectionThreshold(30.0);
  fe.setBriskDetectionMaximumKeypoints(400);
  const double T_WC[7] = {0, 0, 0, -0.5, 0.5, -0.5, 0.5};
  std::vector<okb_keypoint> kl, kr;
  std::vector<uint8_t> dl, dr;
  const int nl = fe.detectAndDescribe(0, left.data(), W, H, W, cam, T_WC, kl, dl);
  const int nr = fe.detectAndDescribe(1, right.data(), W, H, W, cam, T_WC, kr, dr);
  DenseMatcher matcher(fe.context(), 4, 4, false);
  HammingAlgorithm algo{&dl, &dr, fe.descriptorBytes(), {}, {}};
  matcher.match(algo);
  int consistent = 0;
  for (auto& m : algo.matches) {
    const float dx = kr[m[1]].x - kl[m[0]].x, dy = kr[m[1]].y - kl[m[0]].y;
    if (std::fabs(dx - shift) < 1.5f && std::fabs(dy) < 1.5f) ++consistent;
  }
  // gated: B sits 12 px * Z / f to the left of A for a fronto-parallel plane at Z = 5 m
  GatedStereoAlgorithm galgo;
  galgo.dA = &dl; galgo.dB = &dr; galgo.bytes = fe.descriptorBytes(); galgo.kA = &kl; galgo.kB = &kr; galgo.cam = cam;
  galgo.baseline_x = -(double)shift * 5.0 / cam.fu;
  matcher.matchGated(galgo);
  int gated_consistent = 0;
  for (auto& m : galgo.matches) {
    const float dx = kr[m[1]].x - kl[m[0]].x, dy = kr[m[1]].y - kl[m[0]].y;
    if (std::fabs(dx - shift) < 1.5f && std::fabs(dy) < 1.5f) ++gated_consistent;
  }
  std::vector<uint32_t> rp, ci; std::vector<uint16_t> cd;
  matcher.candidates(dl.data(), nl, dr.data(), nr, fe.descriptorBytes(), 60.f, rp, ci, cd);
  double pose[7] = {0, 0, 0, 0, 0, 0, 1}, sb[9] = {0.2, 0, 0, 0, 0, 0, 0, 0, 0};
  okb_imu_params imu{};
  imu.a_max = 176; imu.g_max = 7.8; imu.sigma_g_c = 12e-4; imu.sigma_a_c = 8e-3; imu.sigma_bg = 0.03; imu.sigma_ba = 0.1;
  imu.sigma_gw_c = 4e-6; imu.sigma_aw_c = 4e-5; imu.tau = 3600; imu.g = 9.81007; imu.rate = 200;
  std::vector<okb_imu_sample> smp;
  for (int64_t t = 0; t <= 210000000LL; t += 5000000LL) { okb_imu_sample m{}; m.t_ns = t; m.acc[2] = 9.81007; smp.push_back(m); }
  const bool prop = fe.propagation(smp.data(), (int)smp.size(), imu, pose, sb, 5000000LL, 205000000LL, nullptr, nullptr);
  std::printf("{\"keypoints\": [%d, %d], \"matches\": %zu, \"consistent\": %d, \"gated_matches\": %zu, \"gated_consistent\": %d, \"candidates\": %zu, "
              "\"propagated_x\": %.6f, \"propagation_ok\": %d, \"initialized\": %d}\n", nl, nr, algo.matches.size(), consistent, galgo.matches.size(),
              gated_consistent, ci.size(), pose[0], prop ? 1 : 0, fe.isInitialized() ? 1 : 0);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
    try { Estimator e(0); std::printf("{\"created\": true}\n"); }
    catch (const std::exception& ex) { std::printf("{\"created\": false, \"error\": \"%s\"}\n", ex.what()); }
    return 0;
  }
  if (argc > 1 && !std::strcmp(argv[1], "--frontend")) return frontend_demo();
  if (argc > 1 && !std::strcmp(argv[1], "--sliding")) return sliding(argc > 2 ? std::atoi(argv[2]) : 30);
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
