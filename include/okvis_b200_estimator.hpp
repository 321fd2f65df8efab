// okvis_b200_estimator.hpp -- C++ host shim above the C-ABI (include/okvis_b200.h), mirroring the public
// surface of okvis::Estimator (okvis_ceres/include/okvis/Estimator.hpp:77-412) for the methods the hot
// path needs.  It owns the host book-keeping the reference keeps in statesMap_ / landmarksMap_ /
// okvis::ceres::Map (id <-> block maps, observation list) and forwards all numeric work to the device:
// optimize() = okb_window_upload + okb_optimize + okb_window_download, state prediction in addStates() =
// okb_imu_propagate.  Eigen-free: poses are [t(3), q_xyzw(4)] arrays exactly as PoseParameterBlock
// stores them (okvis_ceres/src/PoseParameterBlock.cpp:68-79), so an OKVIS build can wrap these calls
// one-to-one (see INTEGRATION.md for the adaptor a maintainer would add).
//
// Not provided in round 1 (SURVEY.md 8f "next"): applyMarginalizationStrategy (device-side
// marginalisation), estimated extrinsics.  Both fail loudly.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "okvis_b200.h"

namespace okvis_b200 {

using Pose7 = std::array<double, 7>;        // [t, q_xyzw]
using SpeedAndBias = std::array<double, 9>; // [v_W, b_g, b_a]
using Vec4 = std::array<double, 4>;

struct ImuMeasurement { int64_t t_ns; double gyro[3]; double acc[3]; };

// okvis::ExtrinsicsEstimationParameters (okvis_common/include/okvis/Parameters.hpp:60-96)
struct ExtrinsicsEstimationParameters {
  double sigma_absolute_translation = 0, sigma_absolute_orientation = 0;
  double sigma_c_relative_translation = 0, sigma_c_relative_orientation = 0;
};

// okvis::MapPoint (okvis_common/include/okvis/FrameTypedefs.hpp)
struct MapPoint {
  uint64_t id = 0;
  Vec4 point{{0, 0, 0, 1}};
  double quality = 0, distance = 0;
  std::map<std::array<uint64_t, 3>, uint64_t> observations;   // (poseId, camIdx, keypointIdx) -> residual id
};

class Estimator {
 public:
  explicit Estimator(int device = 0) {
    if (okb_ctx_create(device, 1, &ctx_) != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(nullptr));
  }
  ~Estimator() { okb_ctx_destroy(ctx_); }
  Estimator(const Estimator&) = delete;
  Estimator& operator=(const Estimator&) = delete;

  // ---- sensor configuration (Estimator.cpp:82-110).  The camera geometry and T_SC travel with the
  // call because the reference reads them from the MultiFrame in addStates/addObservation.
  int addCamera(const ExtrinsicsEstimationParameters& p, const okb_camera& geometry, const Pose7& T_SC) {
    if (p.sigma_absolute_translation > 1e-8 || p.sigma_absolute_orientation > 1e-8 || p.sigma_c_relative_translation > 1e-12 ||
        p.sigma_c_relative_orientation > 1e-12)
      throw std::runtime_error("okvis_b200: online extrinsics estimation is not supported by the device solver (round 1)");
    extrinsicsParams_.push_back(p); cameras_.push_back(geometry); T_SC_.push_back(T_SC);
    return (int)cameras_.size() - 1;
  }
  int addImu(const okb_imu_params& p) { if (hasImu_) return -1; imu_ = p; hasImu_ = true; return 0; }
  void clearCameras() { extrinsicsParams_.clear(); cameras_.clear(); T_SC_.clear(); }
  void clearImus() { hasImu_ = false; }

  // ---- Estimator::addStates (Estimator.cpp:110-343): first frame -> gravity-aligned pose + priors;
  // later frames -> propagate the last state through the IMU samples, add an ImuError term.
  bool addStates(uint64_t frameId, int64_t timestamp_ns, const std::vector<ImuMeasurement>& imu, bool asKeyframe) {
    if (!hasImu_ || states_.count(frameId)) return false;
    State st;
    st.id = frameId; st.t_ns = timestamp_ns; st.isKeyframe = asKeyframe;
    if (states_.empty()) {
      if (!initPoseFromImu(imu, st.T_WS)) return false;
      st.sb.fill(0.0);
      for (int k = 0; k < 3; ++k) st.sb[6 + k] = imu_.a0[k];
      // pose prior, information diag(1e8,1e8,1e8,0,0,1e8) with the reference's LLT result
      okb_pose_prior pp{};
      for (int k = 0; k < 7; ++k) pp.meas[k] = st.T_WS[k];
      const double d[6] = {1e4, 1e4, 1e4, 0, 0, 1e8};
      for (int k = 0; k < 6; ++k) pp.sqrt_info[k * 6 + k] = d[k];
      posePrior_ = pp; posePriorFrame_ = frameId; hasPosePrior_ = true;
      okb_sb_prior sp{};
      for (int k = 0; k < 9; ++k) sp.meas[k] = st.sb[k];
      const double s[9] = {1, 1, 1, 1 / imu_.sigma_bg, 1 / imu_.sigma_bg, 1 / imu_.sigma_bg, 1 / imu_.sigma_ba, 1 / imu_.sigma_ba, 1 / imu_.sigma_ba};
      for (int k = 0; k < 9; ++k) sp.sqrt_info[k * 9 + k] = s[k];
      sbPrior_ = sp; sbPriorFrame_ = frameId; hasSbPrior_ = true;
    } else {
      const State& prev = states_.rbegin()->second;
      st.T_WS = prev.T_WS; st.sb = prev.sb;
      std::vector<okb_imu_sample> s = toSamples(imu);
      int used = 0;
      if (okb_imu_propagate(ctx_, &imu_, s.data(), (int)s.size(), prev.t_ns, timestamp_ns, st.T_WS.data(), st.sb.data(), nullptr, nullptr, &used) != OKB_OK || used < 1)
        return false;
      st.imuFromPrev = s;
      st.hasImuTerm = true;
    }
    states_[frameId] = st;
    return true;
  }

  // ---- landmarks / observations (Estimator.cpp:345-413, implementation/Estimator.hpp:43-90)
  bool addLandmark(uint64_t landmarkId, const Vec4& hp) {
    if (landmarks_.count(landmarkId)) return false;
    MapPoint mp; mp.id = landmarkId; mp.point = hp;
    mp.distance = std::fabs(hp[3]) > 1e-8 ? std::sqrt(hp[0] * hp[0] + hp[1] * hp[1] + hp[2] * hp[2]) / std::fabs(hp[3]) : 1e300;
    landmarks_[landmarkId] = mp;
    return true;
  }
  // returns the residual block id (0 = duplicate observation, like the reference's NULL)
  uint64_t addObservation(uint64_t landmarkId, uint64_t poseId, size_t camIdx, size_t keypointIdx, const double kp[2], double keypointSize) {
    auto it = landmarks_.find(landmarkId);
    if (it == landmarks_.end() || !states_.count(poseId) || camIdx >= cameras_.size()) throw std::runtime_error("okvis_b200: addObservation on unknown ids");
    const std::array<uint64_t, 3> kid{{poseId, (uint64_t)camIdx, (uint64_t)keypointIdx}};
    if (it->second.observations.count(kid)) return 0;
    Obs o; o.lm = landmarkId; o.pose = poseId; o.cam = (uint32_t)camIdx; o.z[0] = kp[0]; o.z[1] = kp[1];
    o.sqrt_info = 8.0 / keypointSize;       // information = 64/size^2 * I2
    const uint64_t id = ++nextResidualId_;
    obs_[id] = o;
    it->second.observations[kid] = id;
    return id;
  }
  bool removeObservation(uint64_t residualBlockId) {
    auto it = obs_.find(residualBlockId);
    if (it == obs_.end()) return false;
    auto& m = landmarks_[it->second.lm].observations;
    for (auto o = m.begin(); o != m.end();) o = (o->second == residualBlockId) ? m.erase(o) : std::next(o);
    obs_.erase(it);
    return true;
  }
  bool applyMarginalizationStrategy(size_t, size_t) {
    throw std::runtime_error("okvis_b200: applyMarginalizationStrategy is a next-tier row (SURVEY 8f-1), not built in round 1");
  }

  // ---- Estimator::optimize (Estimator.cpp:843-906) + setOptimizationTimeLimit (:909-929)
  bool setOptimizationTimeLimit(double timeLimit, int minIterations) { timeLimit_ = timeLimit; minIterations_ = minIterations; return true; }
  void optimize(size_t numIter, size_t /*numThreads*/, bool /*verbose*/) {
    // pack: dense indices in insertion (id) order, as Map does with its hash maps
    std::vector<double> poses, sbs, ext, lms;
    std::map<uint64_t, uint32_t> poseIdx, lmIdx;
    for (auto& kv : states_) { poseIdx[kv.first] = (uint32_t)poseIdx.size(); poses.insert(poses.end(), kv.second.T_WS.begin(), kv.second.T_WS.end()); sbs.insert(sbs.end(), kv.second.sb.begin(), kv.second.sb.end()); }
    for (auto& t : T_SC_) ext.insert(ext.end(), t.begin(), t.end());
    std::vector<uint8_t> extFixed(T_SC_.size(), 1);
    std::vector<uint64_t> lmIds;
    for (auto& kv : landmarks_) if (!kv.second.observations.empty()) { lmIdx[kv.first] = (uint32_t)lmIds.size(); lmIds.push_back(kv.first); lms.insert(lms.end(), kv.second.point.begin(), kv.second.point.end()); }
    std::vector<okb_observation> obs;
    for (auto& kv : obs_) {
      okb_observation o{};
      o.pose_idx = poseIdx.at(kv.second.pose); o.lm_idx = lmIdx.at(kv.second.lm); o.ext_idx = kv.second.cam; o.cam_idx = kv.second.cam;
      o.z[0] = kv.second.z[0]; o.z[1] = kv.second.z[1]; o.sqrt_info = kv.second.sqrt_info;
      obs.push_back(o);
    }
    std::vector<okb_imu_term> terms; std::vector<okb_imu_sample> samples;
    uint64_t prevId = 0; bool havePrev = false;
    for (auto& kv : states_) {
      if (havePrev && kv.second.hasImuTerm) {
        okb_imu_term t{};
        t.pose0 = t.sb0 = poseIdx.at(prevId); t.pose1 = t.sb1 = poseIdx.at(kv.first);
        t.t0_ns = states_.at(prevId).t_ns; t.t1_ns = kv.second.t_ns;
        t.sample_offset = (uint32_t)samples.size(); t.sample_count = (uint32_t)kv.second.imuFromPrev.size();
        samples.insert(samples.end(), kv.second.imuFromPrev.begin(), kv.second.imuFromPrev.end());
        terms.push_back(t);
      }
      prevId = kv.first; havePrev = true;
    }
    std::vector<okb_pose_prior> pps; std::vector<okb_sb_prior> sps;
    if (hasPosePrior_ && poseIdx.count(posePriorFrame_)) { okb_pose_prior p = posePrior_; p.pose_idx = poseIdx.at(posePriorFrame_); pps.push_back(p); }
    if (hasSbPrior_ && poseIdx.count(sbPriorFrame_)) { okb_sb_prior p = sbPrior_; p.sb_idx = poseIdx.at(sbPriorFrame_); sps.push_back(p); }
    if (lmIds.empty() || obs.empty()) return;   // nothing the device path would change that the reference would not
    okb_window_desc d{};
    d.n_poses = d.n_speed_bias = (int)states_.size(); d.n_extrinsics = (int)T_SC_.size(); d.n_landmarks = (int)lmIds.size();
    d.n_cameras = (int)cameras_.size(); d.n_obs = (int)obs.size(); d.n_imu_terms = (int)terms.size(); d.n_imu_samples = (int)samples.size();
    d.n_pose_priors = (int)pps.size(); d.n_sb_priors = (int)sps.size();
    d.poses = poses.data(); d.speed_bias = sbs.data(); d.extrinsics = ext.data(); d.extrinsics_fixed = extFixed.data(); d.landmarks = lms.data();
    d.cameras = cameras_.data(); d.obs = obs.data(); d.imu_terms = terms.data(); d.imu_samples = samples.data(); d.imu_params = imu_;
    d.pose_priors = pps.data(); d.sb_priors = sps.data();
    check(okb_window_upload(ctx_, 0, &d));
    okb_solve_options opt{};
    opt.max_iterations = (int)numIter; opt.min_iterations = minIterations_; opt.time_limit_s = timeLimit_; opt.use_cauchy_loss = 1;
    check(okb_optimize(ctx_, 0, 1, &opt, &summary_));
    std::vector<double> q(lmIds.size());
    check(okb_window_download(ctx_, 0, poses.data(), sbs.data(), lms.data(), q.data()));
    size_t i = 0;
    for (auto& kv : states_) { std::copy(poses.begin() + 7 * i, poses.begin() + 7 * i + 7, kv.second.T_WS.begin()); std::copy(sbs.begin() + 9 * i, sbs.begin() + 9 * i + 9, kv.second.sb.begin()); ++i; }
    for (size_t l = 0; l < lmIds.size(); ++l) { MapPoint& mp = landmarks_[lmIds[l]]; std::copy(lms.begin() + 4 * l, lms.begin() + 4 * l + 4, mp.point.begin()); mp.quality = q[l]; }
  }
  const okb_summary& summary() const { return summary_; }

  // ---- getters / setters (Estimator.cpp:931-1230)
  bool get_T_WS(uint64_t poseId, Pose7& T) const { auto it = states_.find(poseId); if (it == states_.end()) return false; T = it->second.T_WS; return true; }
  bool getSpeedAndBias(uint64_t poseId, uint64_t /*imuIdx*/, SpeedAndBias& sb) const { auto it = states_.find(poseId); if (it == states_.end()) return false; sb = it->second.sb; return true; }
  bool set_T_WS(uint64_t poseId, const Pose7& T) { auto it = states_.find(poseId); if (it == states_.end()) return false; it->second.T_WS = T; return true; }
  bool setSpeedAndBias(uint64_t poseId, size_t, const SpeedAndBias& sb) { auto it = states_.find(poseId); if (it == states_.end()) return false; it->second.sb = sb; return true; }
  bool getLandmark(uint64_t id, MapPoint& mp) const { auto it = landmarks_.find(id); if (it == landmarks_.end()) return false; mp = it->second; return true; }
  bool setLandmark(uint64_t id, const Vec4& hp) { auto it = landmarks_.find(id); if (it == landmarks_.end()) return false; it->second.point = hp; return true; }
  bool isLandmarkAdded(uint64_t id) const { return landmarks_.count(id) != 0; }
  size_t numFrames() const { return states_.size(); }
  size_t numLandmarks() const { return landmarks_.size(); }
  uint64_t currentFrameId() const { return states_.empty() ? 0 : states_.rbegin()->first; }
  uint64_t frameIdByAge(size_t age) const { auto it = states_.rbegin(); for (size_t k = 0; k < age && it != states_.rend(); ++k) ++it; return it == states_.rend() ? 0 : it->first; }
  bool isKeyframe(uint64_t id) const { auto it = states_.find(id); return it != states_.end() && it->second.isKeyframe; }
  void setKeyframe(uint64_t id, bool kf) { auto it = states_.find(id); if (it != states_.end()) it->second.isKeyframe = kf; }
  int64_t timestamp(uint64_t id) const { return states_.at(id).t_ns; }

  // Estimator::initPoseFromImu (Estimator.cpp:811-840): align z_W with the mean specific force.
  static bool initPoseFromImu(const std::vector<ImuMeasurement>& imu, Pose7& T_WS) {
    T_WS = Pose7{{0, 0, 0, 0, 0, 0, 1}};
    if (imu.empty()) return false;
    double a[3] = {0, 0, 0};
    for (auto& m : imu) for (int k = 0; k < 3; ++k) a[k] += m.acc[k];
    const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (n == 0) return false;
    const double e[3] = {a[0] / n, a[1] / n, a[2] / n};
    // poseIncrement = -(ez x e_acc).normalized() * acos(ez . e_acc), applied with Transformation::oplus
    double ax[3] = {-e[1], e[0], 0.0};
    const double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1]);
    const double angle = std::acos(e[2]);
    if (an < 1e-15) return true;
    const double s = -angle / an;
    const double d[3] = {ax[0] * s, ax[1] * s, 0.0};
    const double half = 0.5 * std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double sinc = half > 1e-6 ? std::sin(half) / half : 1.0 - half * half / 6.0;
    T_WS[3] = sinc * 0.5 * d[0]; T_WS[4] = sinc * 0.5 * d[1]; T_WS[5] = sinc * 0.5 * d[2]; T_WS[6] = std::cos(half);
    return true;
  }

 private:
  struct State { uint64_t id = 0; int64_t t_ns = 0; bool isKeyframe = false; Pose7 T_WS{}; SpeedAndBias sb{}; bool hasImuTerm = false; std::vector<okb_imu_sample> imuFromPrev; };
  struct Obs { uint64_t lm, pose; uint32_t cam; double z[2]; double sqrt_info; };
  static std::vector<okb_imu_sample> toSamples(const std::vector<ImuMeasurement>& imu) {
    std::vector<okb_imu_sample> s(imu.size());
    for (size_t i = 0; i < imu.size(); ++i) { s[i].t_ns = imu[i].t_ns; for (int k = 0; k < 3; ++k) { s[i].gyro[k] = imu[i].gyro[k]; s[i].acc[k] = imu[i].acc[k]; } }
    return s;
  }
  void check(int rc) { if (rc != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(ctx_)); }

  okb_ctx* ctx_ = nullptr;
  std::vector<ExtrinsicsEstimationParameters> extrinsicsParams_;
  std::vector<okb_camera> cameras_;
  std::vector<Pose7> T_SC_;
  okb_imu_params imu_{};
  bool hasImu_ = false;
  std::map<uint64_t, State> states_;
  std::map<uint64_t, MapPoint> landmarks_;
  std::map<uint64_t, Obs> obs_;
  uint64_t nextResidualId_ = 0;
  okb_pose_prior posePrior_{}; uint64_t posePriorFrame_ = 0; bool hasPosePrior_ = false;
  okb_sb_prior sbPrior_{}; uint64_t sbPriorFrame_ = 0; bool hasSbPrior_ = false;
  double timeLimit_ = -1.0; int minIterations_ = 0;
  okb_summary summary_{};
};

}  // namespace okvis_b200
