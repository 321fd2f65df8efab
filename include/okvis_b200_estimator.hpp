// okvis_b200_estimator.hpp -- C++ host shim above the C-ABI (include/okvis_b200.h), mirroring the public surface of
// okvis::Estimator (okvis_ceres/include/okvis/Estimator.hpp:77-412).  It owns the host book-keeping the reference
// keeps in statesMap_ / landmarksMap_ / okvis::ceres::Map (id <-> block maps, which residual touches which block) and
// forwards all numeric work to the device:
//   * the keyframe window lives on the device (one slot of the context); addStates / addLandmark / addObservation /
//     removeObservation / set_* forward ONE command each (okb_window_add_frame, ...), so a frame's worth of data
//     crosses PCIe per optimize() -- never the window;
//   * optimize()                      = okb_optimize + okb_window_download           (Estimator.cpp:843-906)
//   * applyMarginalizationStrategy()  = the reference's bookkeeping (which blocks / residuals, Estimator.cpp:434-773)
//                                       on the host + okb_window_marginalize for MarginalizationError's numerics
//                                       (MarginalizationError.cpp:127-435, 507-846) + okb_window_remove_*;
//   * state prediction in addStates() = okb_imu_propagate                            (ImuError.cpp:287-504).
// Eigen-free: poses are [t(3), q_xyzw(4)] arrays exactly as PoseParameterBlock stores them
// (okvis_ceres/src/PoseParameterBlock.cpp:68-79), so an OKVIS build wraps these calls one-to-one (INTEGRATION.md
// shows the adaptor a maintainer would add; the MultiFrame / cv types stay on the OKVIS side of it).
//
// Not supported (fails loudly): online extrinsics estimation (sigma_absolute_* > 0 / sigma_c_relative_* > 0).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <iterator>
#include <map>
#include <ostream>
#include <set>
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

// okvis::KeypointIdentifier / okvis::MapPoint (okvis_common/include/okvis/FrameTypedefs.hpp)
struct KeypointIdentifier {
  uint64_t frameId = 0; size_t cameraIndex = 0, keypointIndex = 0;
  bool operator<(const KeypointIdentifier& o) const {
    if (frameId != o.frameId) return frameId < o.frameId;
    if (cameraIndex != o.cameraIndex) return cameraIndex < o.cameraIndex;
    return keypointIndex < o.keypointIndex;
  }
};
struct MapPoint {
  uint64_t id = 0;
  Vec4 point{{0, 0, 0, 1}};
  double quality = 0, distance = 0;
  std::map<KeypointIdentifier, uint64_t> observations;   // -> residual block id
};
using PointMap = std::map<uint64_t, MapPoint>;
using MapPointVector = std::vector<MapPoint>;

// Device-side capacities of an Estimator's resident window (okb_window_reserve).
struct EstimatorCapacity { int frames = 16, landmarks = 4096, observations = 65536, imu_samples = 8192, marg_dim = 160; };

class Estimator {
 public:
  using Capacity = EstimatorCapacity;

  explicit Estimator(int device = 0, Capacity cap = Capacity()) : cap_(cap) {
    if (okb_ctx_create(device, 1, &ctx_) != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(nullptr));
  }
  ~Estimator() { okb_ctx_destroy(ctx_); }
  Estimator(const Estimator&) = delete;
  Estimator& operator=(const Estimator&) = delete;

  // ---- sensor configuration (Estimator.cpp:82-110).  The camera geometry and T_SC travel with the call because the
  // reference reads them from the MultiFrame in addStates / addObservation.
  int addCamera(const ExtrinsicsEstimationParameters& p, const okb_camera& geometry, const Pose7& T_SC) {
    if (p.sigma_absolute_translation > 1e-8 || p.sigma_absolute_orientation > 1e-8 || p.sigma_c_relative_translation > 1e-12 ||
        p.sigma_c_relative_orientation > 1e-12)
      throw std::runtime_error("okvis_b200: online extrinsics estimation is not supported by the device solver");
    if (resident_) throw std::runtime_error("okvis_b200: cameras must be added before the first optimize()");
    extrinsicsParams_.push_back(p); cameras_.push_back(geometry); T_SC_.push_back(T_SC);
    return (int)cameras_.size() - 1;
  }
  int addImu(const okb_imu_params& p) { if (hasImu_) return -1; imu_ = p; hasImu_ = true; return 0; }
  void clearCameras() { extrinsicsParams_.clear(); cameras_.clear(); T_SC_.clear(); }
  void clearImus() { hasImu_ = false; }

  // ---- Estimator::addStates (Estimator.cpp:110-343): first frame -> gravity-aligned pose + priors; later frames ->
  // propagate the last state through the IMU samples, add an ImuError term.
  bool addStates(uint64_t frameId, int64_t timestamp_ns, const std::vector<ImuMeasurement>& imu, bool asKeyframe) {
    if (!hasImu_ || states_.count(frameId)) return false;
    if (!states_.empty() && frameId < states_.rbegin()->first) return false;      // ids grow with time, as IdProvider's do
    State st;
    st.id = frameId; st.t_ns = timestamp_ns; st.isKeyframe = asKeyframe; st.hasSb = true;
    std::vector<okb_imu_sample> s = toSamples(imu);
    if (states_.empty()) {
      if (!initPoseFromImu(imu, st.T_WS)) return false;
      st.sb.fill(0.0);
      for (int k = 0; k < 3; ++k) st.sb[6 + k] = imu_.a0[k];
      // PoseError, information diag(1e8,1e8,1e8,0,0,1e8) with the reference's LLT result
      okb_pose_prior pp{};
      for (int k = 0; k < 7; ++k) pp.meas[k] = st.T_WS[k];
      const double d[6] = {1e4, 1e4, 1e4, 0, 0, 1e8};
      for (int k = 0; k < 6; ++k) pp.sqrt_info[k * 6 + k] = d[k];
      posePriors_.push_back({frameId, pp});
      okb_sb_prior sp{};                      // SpeedAndBiasError(speedAndBias, 1.0, sigma_bg^2, sigma_ba^2)
      for (int k = 0; k < 9; ++k) sp.meas[k] = st.sb[k];
      const double si[9] = {1, 1, 1, 1 / imu_.sigma_bg, 1 / imu_.sigma_bg, 1 / imu_.sigma_bg, 1 / imu_.sigma_ba, 1 / imu_.sigma_ba, 1 / imu_.sigma_ba};
      for (int k = 0; k < 9; ++k) sp.sqrt_info[k * 9 + k] = si[k];
      sbPriors_.push_back({frameId, sp});
    } else {
      const State& prev = states_.rbegin()->second;
      if (!prev.hasSb) return false;
      st.T_WS = prev.T_WS; st.sb = prev.sb;
      int used = 0;
      if (okb_imu_propagate(ctx_, &imu_, s.data(), (int)s.size(), prev.t_ns, timestamp_ns, st.T_WS.data(), st.sb.data(), nullptr, nullptr, &used) != OKB_OK || used < 1)
        return false;
      st.imuFromPrev = s;
      st.hasImuTerm = true;
    }
    const bool first = states_.empty();
    const uint64_t prevId = first ? 0 : states_.rbegin()->first;
    states_[frameId] = st;
    if (!first) terms_.push_back({prevId, frameId});
    if (resident_) {
      okb_imu_term t{};
      const State& prev = states_.at(prevId);
      t.pose0 = posePos(prevId); t.sb0 = sbPos(prevId); t.pose1 = posePos(frameId); t.sb1 = sbPos(frameId);
      t.t0_ns = prev.t_ns; t.t1_ns = timestamp_ns; t.sample_offset = 0; t.sample_count = (uint32_t)s.size();
      check(okb_window_add_frame(ctx_, 0, st.T_WS.data(), st.sb.data(), &t, s.data(), (int)s.size()));
    }
    return true;
  }

  // ---- landmarks / observations (Estimator.cpp:345-413, implementation/Estimator.hpp:43-90)
  bool addLandmark(uint64_t landmarkId, const Vec4& hp) {
    if (landmarks_.count(landmarkId)) return false;
    MapPoint mp; mp.id = landmarkId; mp.point = hp;
    mp.distance = std::fabs(hp[3]) > 1e-8 ? std::sqrt(hp[0] * hp[0] + hp[1] * hp[1] + hp[2] * hp[2]) / std::fabs(hp[3]) : 1e300;
    landmarks_[landmarkId] = mp;
    uint32_t slot;
    if (!freeSlots_.empty()) { slot = *freeSlots_.begin(); freeSlots_.erase(freeSlots_.begin()); }
    else slot = slotCount_++;
    if ((int)slot >= cap_.landmarks) throw std::runtime_error("okvis_b200: landmark capacity of the resident window exceeded");
    lmSlot_[landmarkId] = slot; slotOwner_[slot] = landmarkId;
    if (resident_) check(okb_window_set_landmarks(ctx_, 0, 1, &slot, hp.data()));
    return true;
  }
  // returns the residual block id (0 = duplicate observation, like the reference's NULL)
  uint64_t addObservation(uint64_t landmarkId, uint64_t poseId, size_t camIdx, size_t keypointIdx, const double kp[2], double keypointSize) {
    auto it = landmarks_.find(landmarkId);
    if (it == landmarks_.end() || !states_.count(poseId) || camIdx >= cameras_.size()) throw std::runtime_error("okvis_b200: addObservation on unknown ids");
    KeypointIdentifier kid; kid.frameId = poseId; kid.cameraIndex = camIdx; kid.keypointIndex = keypointIdx;
    if (it->second.observations.count(kid)) return 0;
    // one reprojection error per (landmark, frame, camera): the device grid holds one cell per such triple
    for (auto& o : it->second.observations)
      if (o.first.frameId == poseId && o.first.cameraIndex == camIdx) return 0;
    Obs o; o.lm = landmarkId; o.pose = poseId; o.cam = (uint32_t)camIdx; o.kp = keypointIdx; o.z[0] = kp[0]; o.z[1] = kp[1];
    o.sqrt_info = 8.0 / keypointSize;       // information = 64/size^2 * I2
    const uint64_t id = ++nextResidualId_;
    obs_[id] = o;
    it->second.observations[kid] = id;
    if (resident_) { okb_observation d = deviceObs(o); check(okb_window_add_observations(ctx_, 0, 1, &d)); }
    return id;
  }
  bool removeObservation(uint64_t residualBlockId) {
    auto it = obs_.find(residualBlockId);
    if (it == obs_.end()) return false;
    if (resident_) {
      okb_obs_key k{posePos(it->second.pose), lmSlot_.at(it->second.lm), it->second.cam, 0};
      check(okb_window_remove_observations(ctx_, 0, 1, &k));
    }
    eraseObsHost(it);
    return true;
  }
  bool removeObservation(uint64_t landmarkId, uint64_t poseId, size_t camIdx, size_t keypointIdx) {
    auto it = landmarks_.find(landmarkId);
    if (it == landmarks_.end()) return false;
    KeypointIdentifier kid; kid.frameId = poseId; kid.cameraIndex = camIdx; kid.keypointIndex = keypointIdx;
    auto o = it->second.observations.find(kid);
    if (o == it->second.observations.end()) return false;
    return removeObservation(o->second);
  }

  // ---- Estimator::applyMarginalizationStrategy (Estimator.cpp:434-773).  Bookkeeping on the host, numerics
  // (linearisation at the first-estimate points, Schur complements, eigen-factorisation) on the device.
  bool applyMarginalizationStrategy(size_t numKeyframes, size_t numImuFrames, MapPointVector& removedLandmarks) {
    // keep the newest numImuFrames
    auto rit = states_.rbegin();
    for (size_t k = 0; k < numImuFrames; ++k) { ++rit; if (rit == states_.rend()) return true; }
    ensureResident();
    std::vector<uint64_t> removeFrames, removeAllButPose, allLinearizedFrames;
    size_t countedKeyframes = 0;
    for (; rit != states_.rend(); ++rit) {
      if (!rit->second.isKeyframe || countedKeyframes >= numKeyframes) removeFrames.push_back(rit->first);
      else countedKeyframes++;
      removeAllButPose.push_back(rit->first);
      allLinearizedFrames.push_back(rit->first);
    }
    auto contains = [](const std::vector<uint64_t>& v, uint64_t q) { return std::find(v.begin(), v.end(), q) != v.end(); };
    // the linear system: blocks of the current prior first (their order), newly connected blocks appended in the
    // order MarginalizationError::addResidualBlock meets them
    std::vector<BlockRef> blocks = margBlocks_;
    std::vector<int32_t> prev(blocks.size());
    for (size_t i = 0; i < blocks.size(); ++i) prev[i] = (int32_t)i;
    std::vector<uint8_t> margFlag(blocks.size(), 0);
    auto connect = [&](int kind, uint64_t id) -> size_t {
      for (size_t i = 0; i < blocks.size(); ++i) if (blocks[i].kind == kind && blocks[i].id == id) return i;
      blocks.push_back({kind, id}); prev.push_back(-1); margFlag.push_back(0);
      return blocks.size() - 1;
    };
    std::vector<uint32_t> jobTerms, jobSbPriors, jobLandmarks;
    std::set<size_t> termDone, sbPriorDone;
    auto linearizeTerm = [&](size_t ti) {
      if (!termDone.insert(ti).second) return;
      connect(OKB_BLOCK_POSE, terms_[ti].first); connect(OKB_BLOCK_SPEED_BIAS, terms_[ti].first);
      connect(OKB_BLOCK_POSE, terms_[ti].second); connect(OKB_BLOCK_SPEED_BIAS, terms_[ti].second);
      jobTerms.push_back((uint32_t)ti);
    };
    bool anything = false;
    // marginalize everything but pose (Estimator.cpp:483-554): the speed/bias block of every frame that left the IMU window
    for (uint64_t fid : removeAllButPose) {
      State& st = states_.at(fid);
      if (!st.hasSb) continue;
      margFlag[connect(OKB_BLOCK_SPEED_BIAS, fid)] = 1;
      anything = true;
      for (size_t ti = 0; ti < terms_.size(); ++ti)
        if (terms_[ti].first == fid || terms_[ti].second == fid) linearizeTerm(ti);
      for (size_t pi = 0; pi < sbPriors_.size(); ++pi)
        if (sbPriors_[pi].frame == fid && sbPriorDone.insert(pi).second) jobSbPriors.push_back((uint32_t)pi);
    }
    // marginalize ONLY pose now (:556-733)
    bool reDoFixation = false;
    std::vector<uint64_t> landmarksGone;             // marginalised or deleted (device slots to clear)
    if (!removeFrames.empty()) {
      const uint64_t currentKfId = allLinearizedFrames.at(0);
      for (uint64_t fid : removeFrames) {
        margFlag[connect(OKB_BLOCK_POSE, fid)] = 1;
        anything = true;
        for (auto& pp : posePriors_) if (pp.frame == fid) reDoFixation = true;       // PoseError is dropped, not linearised
        for (size_t ti = 0; ti < terms_.size(); ++ti)
          if (terms_[ti].first == fid || terms_[ti].second == fid) linearizeTerm(ti);
      }
      for (auto pit = landmarks_.begin(); pit != landmarks_.end();) {
        MapPoint& mp = pit->second;
        bool skipLandmark = true, hasNewObservations = false, justDelete = false, marginalize = true, errorTermAdded = false;
        size_t obsCount = 0;
        for (auto& o : mp.observations) {
          const uint64_t poseId = o.first.frameId;
          if (contains(removeFrames, poseId)) skipLandmark = false;
          if (poseId >= currentKfId) { marginalize = false; hasNewObservations = true; }
          if (contains(allLinearizedFrames, poseId)) obsCount++;
        }
        if (mp.observations.empty()) {
          removedLandmarks.push_back(mp); landmarksGone.push_back(pit->first); pit = landmarks_.erase(pit);
          continue;
        }
        if (skipLandmark) { ++pit; continue; }
        std::vector<uint64_t> toRemove;
        size_t remaining = mp.observations.size();
        for (auto& o : mp.observations) {
          const uint64_t poseId = o.first.frameId;
          if ((contains(removeFrames, poseId) && hasNewObservations) || (!contains(allLinearizedFrames, poseId) && marginalize)) {
            toRemove.push_back(o.second); --remaining;
          } else if (marginalize && contains(allLinearizedFrames, poseId)) {
            if (obsCount < 2) { toRemove.push_back(o.second); --remaining; }
            else { errorTermAdded = true; connect(OKB_BLOCK_POSE, poseId); }
          }
          if (remaining == 0) { justDelete = true; marginalize = false; }
        }
        for (uint64_t rid : toRemove) removeObservation(rid);
        if (justDelete) {
          removedLandmarks.push_back(mp); landmarksGone.push_back(pit->first); pit = landmarks_.erase(pit);
          continue;
        }
        if (marginalize && errorTermAdded) {
          jobLandmarks.push_back(lmSlot_.at(pit->first));
          removedLandmarks.push_back(mp); landmarksGone.push_back(pit->first); pit = landmarks_.erase(pit);
          continue;
        }
        ++pit;
      }
    }
    // ---- numerics on the device (positions as the window is NOW)
    if (anything) {
      std::vector<int32_t> kinds(blocks.size());
      std::vector<uint32_t> idx(blocks.size());
      for (size_t i = 0; i < blocks.size(); ++i) { kinds[i] = blocks[i].kind; idx[i] = blocks[i].kind == OKB_BLOCK_POSE ? posePos(blocks[i].id) : sbPos(blocks[i].id); }
      okb_marg_job job{};
      job.n_blocks = (int32_t)blocks.size(); job.n_imu_terms = (int32_t)jobTerms.size(); job.n_sb_priors = (int32_t)jobSbPriors.size();
      job.n_landmarks = (int32_t)jobLandmarks.size();
      job.block_kind = kinds.data(); job.block_idx = idx.data(); job.block_prev = prev.data(); job.block_marginalize = margFlag.data();
      job.imu_terms = jobTerms.data(); job.sb_priors = jobSbPriors.data(); job.landmarks = jobLandmarks.data();
      check(okb_window_marginalize(ctx_, 0, &job));
      std::vector<BlockRef> kept;
      for (size_t i = 0; i < blocks.size(); ++i) if (!margFlag[i]) kept.push_back(blocks[i]);
      margBlocks_.swap(kept);
    }
    // ---- the same structural edits on the device and in the host maps
    for (uint64_t lid : landmarksGone) {
      // observations still attached to a deleted landmark die with it on the device (okb_window_remove_landmarks)
      for (auto o = obs_.begin(); o != obs_.end();) o = (o->second.lm == lid) ? obs_.erase(o) : std::next(o);
      const uint32_t slot = lmSlot_.at(lid);
      check(okb_window_remove_landmarks(ctx_, 0, 1, &slot));
      lmSlot_.erase(lid); slotOwner_.erase(slot); freeSlots_.insert(slot);
    }
    for (uint64_t fid : removeAllButPose) {          // newest first: positions of older blocks do not move
      State& st = states_.at(fid);
      const bool whole = contains(removeFrames, fid);
      if (whole) {
        check(okb_window_remove_frame(ctx_, 0, posePos(fid), st.hasSb ? sbPos(fid) : 0xffffffffu));
        for (auto o = obs_.begin(); o != obs_.end();) {
          if (o->second.pose == fid) {
            auto lm = landmarks_.find(o->second.lm);
            if (lm != landmarks_.end()) eraseKey(lm->second, o->first);
            o = obs_.erase(o);
          } else ++o;
        }
      } else if (st.hasSb) {
        check(okb_window_remove_speed_bias(ctx_, 0, sbPos(fid)));
      }
      // host mirrors of the device lists (stable compaction, exactly what the interpreter does)
      eraseIf(terms_, [&](const TermRef& t) { return (t.first == fid || t.second == fid); });
      eraseIf(sbPriors_, [&](const SbPrior& p) { return p.frame == fid; });
      if (whole) { eraseIf(posePriors_, [&](const PosePrior& p) { return p.frame == fid; }); states_.erase(fid); }
      else st.hasSb = false;
    }
    if (reDoFixation && !states_.empty()) {
      // finally fix the first pose properly (Estimator.cpp:761-770): information diag(1e14 x3, 0, 0, 1e14); the
      // reference's LLT leaves diag(1e7,1e7,1e7,0,0,1e14) as the square root
      const State& first = states_.begin()->second;
      okb_pose_prior pp{};
      for (int k = 0; k < 7; ++k) pp.meas[k] = first.T_WS[k];
      const double d[6] = {1e7, 1e7, 1e7, 0, 0, 1e14};
      for (int k = 0; k < 6; ++k) pp.sqrt_info[k * 6 + k] = d[k];
      posePriors_.push_back({first.id, pp});
    }
    pushPriors();
    return true;
  }
  bool applyMarginalizationStrategy(size_t numKeyframes, size_t numImuFrames) { MapPointVector r; return applyMarginalizationStrategy(numKeyframes, numImuFrames, r); }

  // ---- Estimator::optimize (Estimator.cpp:843-906) + setOptimizationTimeLimit (:909-929)
  bool setOptimizationTimeLimit(double timeLimit, int minIterations) { timeLimit_ = timeLimit; minIterations_ = minIterations; return true; }
  void optimize(size_t numIter, size_t /*numThreads*/ = 1, bool /*verbose*/ = false) {
    if (states_.empty() || cameras_.empty()) return;
    ensureResident();
    okb_solve_options opt{};
    opt.max_iterations = (int)numIter; opt.min_iterations = minIterations_; opt.time_limit_s = timeLimit_; opt.use_cauchy_loss = 1;
    check(okb_optimize(ctx_, 0, 1, &opt, &summary_));
    const size_t K = states_.size(), NSB = numSb(), L = slotCount_ ? slotCount_ : 1;
    std::vector<double> poses(7 * K), sbs(9 * std::max<size_t>(NSB, 1)), lms(4 * L), q(L);
    check(okb_window_download(ctx_, 0, poses.data(), sbs.data(), lms.data(), q.data()));
    size_t i = 0, j = 0;
    for (auto& kv : states_) {
      std::copy(poses.begin() + 7 * i, poses.begin() + 7 * i + 7, kv.second.T_WS.begin()); ++i;
      if (kv.second.hasSb) { std::copy(sbs.begin() + 9 * j, sbs.begin() + 9 * j + 9, kv.second.sb.begin()); ++j; }
    }
    for (auto& kv : landmarks_) {            // Estimator.cpp:880-900: point + quality written back for every landmark
      const uint32_t s = lmSlot_.at(kv.first);
      std::copy(lms.begin() + 4 * s, lms.begin() + 4 * s + 4, kv.second.point.begin());
      kv.second.quality = q[s];
      const Vec4& hp = kv.second.point;
      kv.second.distance = std::fabs(hp[3]) > 1e-8 ? std::sqrt(hp[0] * hp[0] + hp[1] * hp[1] + hp[2] * hp[2]) / std::fabs(hp[3]) : 1e300;
    }
  }
  const okb_summary& summary() const { return summary_; }
  // bytes the last commit of the window moved host -> device (a frame's worth once the window is resident)
  int64_t lastUploadBytes() const { return okb_window_h2d_bytes(ctx_, 0); }

  // ---- getters / setters (Estimator.cpp:931-1230)
  bool get_T_WS(uint64_t poseId, Pose7& T) const { auto it = states_.find(poseId); if (it == states_.end()) return false; T = it->second.T_WS; return true; }
  bool getSpeedAndBias(uint64_t poseId, uint64_t /*imuIdx*/, SpeedAndBias& sb) const {
    auto it = states_.find(poseId);
    if (it == states_.end() || !it->second.hasSb) return false;
    sb = it->second.sb; return true;
  }
  bool getCameraSensorStates(uint64_t poseId, size_t cameraIdx, Pose7& T_SCi) const {
    if (!states_.count(poseId) || cameraIdx >= T_SC_.size()) return false;
    T_SCi = T_SC_[cameraIdx]; return true;
  }
  bool set_T_WS(uint64_t poseId, const Pose7& T) {
    auto it = states_.find(poseId); if (it == states_.end()) return false;
    it->second.T_WS = T;
    if (resident_) { const uint32_t p = posePos(poseId); check(okb_window_set_states(ctx_, 0, 1, &p, T.data(), 0, nullptr, nullptr)); }
    return true;
  }
  bool setSpeedAndBias(uint64_t poseId, size_t /*imuIdx*/, const SpeedAndBias& sb) {
    auto it = states_.find(poseId); if (it == states_.end() || !it->second.hasSb) return false;
    it->second.sb = sb;
    if (resident_) { const uint32_t p = sbPos(poseId); check(okb_window_set_states(ctx_, 0, 0, nullptr, nullptr, 1, &p, sb.data())); }
    return true;
  }
  bool setCameraSensorStates(uint64_t poseId, size_t cameraIdx, const Pose7& T_SCi) {
    if (!states_.count(poseId) || cameraIdx >= T_SC_.size()) return false;
    if (resident_) throw std::runtime_error("okvis_b200: extrinsics are fixed once the window is resident");
    T_SC_[cameraIdx] = T_SCi; return true;
  }
  bool getLandmark(uint64_t id, MapPoint& mp) const { auto it = landmarks_.find(id); if (it == landmarks_.end()) return false; mp = it->second; return true; }
  size_t getLandmarks(PointMap& landmarks) const { landmarks = landmarks_; return landmarks_.size(); }
  size_t getLandmarks(MapPointVector& landmarks) const {
    landmarks.clear(); landmarks.reserve(landmarks_.size());
    for (auto& kv : landmarks_) landmarks.push_back(kv.second);
    return landmarks_.size();
  }
  bool setLandmark(uint64_t id, const Vec4& hp) {
    auto it = landmarks_.find(id); if (it == landmarks_.end()) return false;
    it->second.point = hp;
    if (resident_) { const uint32_t s = lmSlot_.at(id); check(okb_window_set_landmarks(ctx_, 0, 1, &s, hp.data())); }
    return true;
  }
  bool isLandmarkAdded(uint64_t id) const { return landmarks_.count(id) != 0; }
  bool isLandmarkInitialized(uint64_t id) const { auto it = lmInitialized_.find(id); return it != lmInitialized_.end() && it->second; }
  void setLandmarkInitialized(uint64_t id, bool initialized) { if (landmarks_.count(id)) lmInitialized_[id] = initialized; }
  size_t numFrames() const { return states_.size(); }
  size_t numLandmarks() const { return landmarks_.size(); }
  uint64_t currentFrameId() const { return states_.empty() ? 0 : states_.rbegin()->first; }
  uint64_t currentKeyframeId() const {
    for (auto rit = states_.rbegin(); rit != states_.rend(); ++rit) if (rit->second.isKeyframe) return rit->first;
    return 0;
  }
  uint64_t frameIdByAge(size_t age) const { auto it = states_.rbegin(); for (size_t k = 0; k < age && it != states_.rend(); ++k) ++it; return it == states_.rend() ? 0 : it->first; }
  bool isKeyframe(uint64_t id) const { auto it = states_.find(id); return it != states_.end() && it->second.isKeyframe; }
  bool isInImuWindow(uint64_t id) const { auto it = states_.find(id); return it != states_.end() && it->second.hasSb; }
  void setKeyframe(uint64_t id, bool kf) { auto it = states_.find(id); if (it != states_.end()) it->second.isKeyframe = kf; }
  int64_t timestamp(uint64_t id) const { return states_.at(id).t_ns; }
  void printStates(uint64_t poseId, std::ostream& buffer) const {
    auto it = states_.find(poseId);
    if (it == states_.end()) return;
    buffer << "GLOBAL: id=" << poseId << ":pose " << (it->second.isKeyframe ? "(keyframe) " : "") << "SENSOR: ";
    if (it->second.hasSb) buffer << "speedAndBias ";
    for (size_t c = 0; c < T_SC_.size(); ++c) buffer << "(extrinsics " << c << ") ";
    buffer << "\n";
  }

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
  struct State { uint64_t id = 0; int64_t t_ns = 0; bool isKeyframe = false, hasSb = false; Pose7 T_WS{}; SpeedAndBias sb{}; bool hasImuTerm = false; std::vector<okb_imu_sample> imuFromPrev; };
  struct Obs { uint64_t lm, pose; uint32_t cam; size_t kp; double z[2]; double sqrt_info; };
  struct BlockRef { int kind; uint64_t id; };
  using TermRef = std::pair<uint64_t, uint64_t>;            // ImuError between two frame ids, in the device's list order
  struct PosePrior { uint64_t frame; okb_pose_prior p; };
  struct SbPrior { uint64_t frame; okb_sb_prior p; };

  template <class V, class F> static void eraseIf(V& v, F f) { v.erase(std::remove_if(v.begin(), v.end(), f), v.end()); }
  static std::vector<okb_imu_sample> toSamples(const std::vector<ImuMeasurement>& imu) {
    std::vector<okb_imu_sample> s(imu.size());
    for (size_t i = 0; i < imu.size(); ++i) { s[i].t_ns = imu[i].t_ns; for (int k = 0; k < 3; ++k) { s[i].gyro[k] = imu[i].gyro[k]; s[i].acc[k] = imu[i].acc[k]; } }
    return s;
  }
  void check(int rc) const { if (rc != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(ctx_)); }
  uint32_t posePos(uint64_t id) const { return (uint32_t)std::distance(states_.begin(), states_.find(id)); }
  uint32_t sbPos(uint64_t id) const {
    uint32_t p = 0;
    for (auto& kv : states_) { if (kv.first == id) return p; if (kv.second.hasSb) ++p; }
    return p;
  }
  size_t numSb() const { size_t n = 0; for (auto& kv : states_) n += kv.second.hasSb; return n; }
  okb_observation deviceObs(const Obs& o) const {
    okb_observation d{};
    d.pose_idx = posePos(o.pose); d.lm_idx = lmSlot_.at(o.lm); d.ext_idx = o.cam; d.cam_idx = o.cam;
    d.z[0] = o.z[0]; d.z[1] = o.z[1]; d.sqrt_info = o.sqrt_info;
    return d;
  }
  static void eraseKey(MapPoint& mp, uint64_t residualId) {
    for (auto o = mp.observations.begin(); o != mp.observations.end();) o = (o->second == residualId) ? mp.observations.erase(o) : std::next(o);
  }
  void eraseObsHost(std::map<uint64_t, Obs>::iterator it) {
    auto lm = landmarks_.find(it->second.lm);
    if (lm != landmarks_.end()) eraseKey(lm->second, it->first);
    obs_.erase(it);
  }
  void pushPriors() {
    if (!resident_) return;
    std::vector<okb_pose_prior> pps; std::vector<okb_sb_prior> sps;
    for (auto& p : posePriors_) { okb_pose_prior q = p.p; q.pose_idx = posePos(p.frame); pps.push_back(q); }
    for (auto& p : sbPriors_) { okb_sb_prior q = p.p; q.sb_idx = sbPos(p.frame); sps.push_back(q); }
    check(okb_window_set_priors(ctx_, 0, (int)pps.size(), pps.data(), (int)sps.size(), sps.data(), nullptr));
  }
  // First contact with the device: the whole graph once (okb_window_upload); everything after that is incremental.
  void ensureResident() {
    if (resident_) return;
    std::vector<double> poses, sbs, ext;
    for (auto& kv : states_) { poses.insert(poses.end(), kv.second.T_WS.begin(), kv.second.T_WS.end()); if (kv.second.hasSb) sbs.insert(sbs.end(), kv.second.sb.begin(), kv.second.sb.end()); }
    for (auto& t : T_SC_) ext.insert(ext.end(), t.begin(), t.end());
    std::vector<uint8_t> extFixed(T_SC_.size(), 1);
    const size_t L = slotCount_ ? slotCount_ : 1;
    std::vector<double> lms(4 * L, 0.0);
    for (auto& kv : landmarks_) std::copy(kv.second.point.begin(), kv.second.point.end(), lms.begin() + 4 * lmSlot_.at(kv.first));
    std::vector<okb_observation> obs;
    for (auto& kv : obs_) obs.push_back(deviceObs(kv.second));
    std::vector<okb_imu_term> terms; std::vector<okb_imu_sample> samples;
    for (auto& tr : terms_) {
      const State& s0 = states_.at(tr.first); const State& s1 = states_.at(tr.second);
      okb_imu_term t{};
      t.pose0 = posePos(tr.first); t.sb0 = sbPos(tr.first); t.pose1 = posePos(tr.second); t.sb1 = sbPos(tr.second);
      t.t0_ns = s0.t_ns; t.t1_ns = s1.t_ns;
      t.sample_offset = (uint32_t)samples.size(); t.sample_count = (uint32_t)s1.imuFromPrev.size();
      samples.insert(samples.end(), s1.imuFromPrev.begin(), s1.imuFromPrev.end());
      terms.push_back(t);
    }
    std::vector<okb_pose_prior> pps; std::vector<okb_sb_prior> sps;
    for (auto& p : posePriors_) { okb_pose_prior q = p.p; q.pose_idx = posePos(p.frame); pps.push_back(q); }
    for (auto& p : sbPriors_) { okb_sb_prior q = p.p; q.sb_idx = sbPos(p.frame); sps.push_back(q); }
    okb_window_desc d{};
    d.n_poses = (int)states_.size(); d.n_speed_bias = (int)numSb(); d.n_extrinsics = (int)T_SC_.size(); d.n_landmarks = (int)L;
    d.n_cameras = (int)cameras_.size(); d.n_obs = (int)obs.size(); d.n_imu_terms = (int)terms.size(); d.n_imu_samples = (int)samples.size();
    d.n_pose_priors = (int)pps.size(); d.n_sb_priors = (int)sps.size();
    d.poses = poses.data(); d.speed_bias = sbs.data(); d.extrinsics = ext.data(); d.extrinsics_fixed = extFixed.data(); d.landmarks = lms.data();
    d.cameras = cameras_.data(); d.obs = obs.data(); d.imu_terms = terms.data(); d.imu_samples = samples.data(); d.imu_params = imu_;
    d.pose_priors = pps.data(); d.sb_priors = sps.data();
    check(okb_window_reserve(ctx_, 0, cap_.frames, cap_.landmarks, cap_.observations, cap_.imu_samples, cap_.marg_dim));
    check(okb_window_upload(ctx_, 0, &d));
    resident_ = true;
  }

  okb_ctx* ctx_ = nullptr;
  Capacity cap_;
  bool resident_ = false;
  std::vector<ExtrinsicsEstimationParameters> extrinsicsParams_;
  std::vector<okb_camera> cameras_;
  std::vector<Pose7> T_SC_;
  okb_imu_params imu_{};
  bool hasImu_ = false;
  std::map<uint64_t, State> states_;
  PointMap landmarks_;
  std::map<uint64_t, bool> lmInitialized_;
  std::map<uint64_t, Obs> obs_;
  uint64_t nextResidualId_ = 0;
  // device mirrors: landmark slots, IMU term / prior lists in the device's order, blocks of the marginalisation prior
  std::map<uint64_t, uint32_t> lmSlot_;
  std::map<uint32_t, uint64_t> slotOwner_;
  std::set<uint32_t> freeSlots_;
  uint32_t slotCount_ = 0;
  std::vector<TermRef> terms_;
  std::vector<PosePrior> posePriors_;
  std::vector<SbPrior> sbPriors_;
  std::vector<BlockRef> margBlocks_;
  double timeLimit_ = -1.0; int minIterations_ = 0;
  okb_summary summary_{};
};

}  // namespace okvis_b200
