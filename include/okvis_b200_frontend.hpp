// okvis_b200_frontend.hpp -- C++ host shims above the C-ABI (include/okvis_b200.h) for the two frontend classes the
// hot path replaces:
//   okvis_b200::Frontend      mirrors okvis::Frontend      (okvis_frontend/include/okvis/Frontend.hpp:71-255):
//       detectAndDescribe (Frontend.cpp:92-114 -> Frame::detect / Frame::describe), propagation (Frontend.cpp:274-292
//       -> ImuError::propagation) and the brisk / keyframe parameter accessors (:137-255);
//   okvis_b200::DenseMatcher  mirrors okvis::DenseMatcher  (okvis_matcher/include/okvis/DenseMatcher.hpp:59-212):
//       match<MATCHING_ALGORITHM>() with the epilogue of matchBody (implementation/DenseMatcher.hpp:48-125) on the
//       host and listBIteration + assignbest (DenseMatcher.cpp:69-110) on the device.
// Eigen- and OpenCV-free: images are raw u8 rows, keypoints are okb_keypoint (the cv::KeyPoint fields the reference
// uses), poses are [t(3), q_xyzw(4)].  INTEGRATION.md shows the adaptor an OKVIS build adds around these.
//
// Frontend::dataAssociationAndInitialization (Frontend.cpp:116-272) is NOT provided: it is the caller of the matcher
// and owns MultiFrame / RANSAC / triangulation glue that stays on the OKVIS side.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "okvis_b200.h"

namespace okvis_b200 {

class Frontend {
 public:
  // own == nullptr: the Frontend creates its own context on `device`; otherwise it borrows the caller's.
  explicit Frontend(size_t numCameras, int device = 0, okb_ctx* shared = nullptr) : numCameras_(numCameras), ctx_(shared), owns_(shared == nullptr) {
    if (owns_ && okb_ctx_create(device, 1, &ctx_) != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(nullptr));
  }
  virtual ~Frontend() { if (owns_) okb_ctx_destroy(ctx_); }
  Frontend(const Frontend&) = delete;
  Frontend& operator=(const Frontend&) = delete;

  // Frontend::detectAndDescribe(cameraIndex, frameOut, T_WC, keypoints): detects on the image of camera `cameraIndex`
  // and describes with the keypoint orientation taken from the gravity direction in that camera (Frame.hpp(impl):128-156).
  // May be called concurrently for different camera indices (one device stream per camera slot), like the reference
  // (ThreadedKFVio.cpp:131, 425).  Returns the number of keypoints; descriptors come back row-contiguous [n][bytes].
  int detectAndDescribe(size_t cameraIndex, const uint8_t* image, int width, int height, int stride, const okb_camera& geometry,
                        const double T_WC[7], std::vector<okb_keypoint>& keypoints, std::vector<uint8_t>& descriptors) {
    if (cameraIndex >= numCameras_) throw std::runtime_error("okvis_b200: camera index out of range");
    // R_CW = C_WC^T from the quaternion of T_WC
    const double x = T_WC[3], y = T_WC[4], z = T_WC[5], w = T_WC[6];
    const double C[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    const double R_CW[9] = {C[0], C[3], C[6], C[1], C[4], C[7], C[2], C[5], C[8]};
    okb_detect_params p{};
    p.uniformity_radius = briskDetectionThreshold_;          // the "detection threshold" of the yaml is brisk's uniformity radius (Frontend.cpp:828-831)
    p.absolute_threshold = briskDetectionAbsoluteThreshold_;
    p.max_keypoints = (int32_t)briskDetectionMaximumKeypoints_;
    p.desc_bytes = descriptorBytes_;
    p.rotation_invariance = briskDescriptionRotationInvariance_ ? 1 : 0;
    keypoints.resize(briskDetectionMaximumKeypoints_);
    descriptors.resize(briskDetectionMaximumKeypoints_ * (size_t)descriptorBytes_);
    int n = 0;
    const int rc = okb_detect_describe(ctx_, (int)cameraIndex, image, width, height, stride, &geometry, R_CW, &p, keypoints.data(), descriptors.data(),
                                       (int)briskDetectionMaximumKeypoints_, &n);
    if (rc != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(ctx_));
    keypoints.resize(n);
    descriptors.resize((size_t)n * descriptorBytes_);
    isInitialized_ = true;
    return n;
  }

  // Frontend::propagation (Frontend.cpp:274-292): ImuError::propagation on the device; covariance / jacobian 15x15 or null.
  bool propagation(const okb_imu_sample* imuMeasurements, int n, const okb_imu_params& imuParams, double T_WS_propagated[7], double speedAndBiases[9],
                   int64_t t_start_ns, int64_t t_end_ns, double* covariance, double* jacobian) const {
    if (n < 2) return false;
    int used = 0;
    const int rc = okb_imu_propagate(ctx_, &imuParams, imuMeasurements, n, t_start_ns, t_end_ns, T_WS_propagated, speedAndBiases, covariance, jacobian, &used);
    return rc == OKB_OK && used > 0;
  }

  // ---- accessors (Frontend.hpp:137-255)
  size_t getBriskDetectionOctaves() const { return briskDetectionOctaves_; }
  double getBriskDetectionThreshold() const { return briskDetectionThreshold_; }
  double getBriskDetectionAbsoluteThreshold() const { return briskDetectionAbsoluteThreshold_; }
  size_t getBriskDetectionMaximumKeypoints() const { return briskDetectionMaximumKeypoints_; }
  bool getBriskDescriptionRotationInvariance() const { return briskDescriptionRotationInvariance_; }
  bool getBriskDescriptionScaleInvariance() const { return briskDescriptionScaleInvariance_; }
  double getBriskMatchingThreshold() const { return briskMatchingThreshold_; }
  float getKeyframeInsertionOverlapThershold() const { return keyframeInsertionOverlapThreshold_; }
  float getKeyframeInsertionMatchingRatioThreshold() const { return keyframeInsertionMatchingRatioThreshold_; }
  bool isInitialized() const { return isInitialized_; }
  void setBriskDetectionOctaves(size_t octaves) {
    if (octaves != 0) throw std::runtime_error("okvis_b200: the device detector is single scale (octaves = 0, the value OKVIS ships)");
    briskDetectionOctaves_ = octaves;
  }
  void setBriskDetectionThreshold(double threshold) { briskDetectionThreshold_ = threshold; }
  void setBriskDetectionAbsoluteThreshold(double threshold) { briskDetectionAbsoluteThreshold_ = threshold; }
  void setBriskDetectionMaximumKeypoints(size_t maxKeypoints) { briskDetectionMaximumKeypoints_ = maxKeypoints; }
  void setBriskDescriptionRotationInvariance(bool invariance) { briskDescriptionRotationInvariance_ = invariance; }
  void setBriskDescriptionScaleInvariance(bool invariance) {
    if (invariance) throw std::runtime_error("okvis_b200: scale-invariant description is not built (OKVIS constructs the extractor with scaleInvariant = false)");
    briskDescriptionScaleInvariance_ = invariance;
  }
  void setBriskMatchingThreshold(double threshold) { briskMatchingThreshold_ = threshold; }
  void setKeyframeInsertionOverlapThreshold(float threshold) { keyframeInsertionOverlapThreshold_ = threshold; }
  void setKeyframeInsertionMatchingRatioThreshold(float threshold) { keyframeInsertionMatchingRatioThreshold_ = threshold; }
  void setDescriptorBytes(int bytes) { if (bytes != 48 && bytes != 64) throw std::runtime_error("okvis_b200: 48 or 64 byte descriptors"); descriptorBytes_ = bytes; }
  int descriptorBytes() const { return descriptorBytes_; }
  okb_ctx* context() const { return ctx_; }

 private:
  size_t numCameras_;
  okb_ctx* ctx_;
  bool owns_;
  bool isInitialized_ = false;
  // defaults of Frontend::Frontend (Frontend.cpp:59-81)
  size_t briskDetectionOctaves_ = 0;
  double briskDetectionThreshold_ = 50.0;
  double briskDetectionAbsoluteThreshold_ = 800.0;
  size_t briskDetectionMaximumKeypoints_ = 450;
  bool briskDescriptionRotationInvariance_ = true;
  bool briskDescriptionScaleInvariance_ = false;
  double briskMatchingThreshold_ = 60.0;
  float keyframeInsertionOverlapThreshold_ = 0.6f;
  float keyframeInsertionMatchingRatioThreshold_ = 0.2f;
  int descriptorBytes_ = 48;
};

// okvis::DenseMatcher.  MATCHING_ALGORITHM_T provides what okvis::MatchingAlgorithm provides (doSetup, sizeA, sizeB,
// skipA, skipB, distanceThreshold, distanceRatioThreshold, reserveMatches, setBestMatch) plus direct access to the
// binary descriptors instead of the per-pair distance() callback:
//   const uint8_t* descriptorsA() / descriptorsB()   row-contiguous [size][descriptorBytes()]
// (with the pure Hamming MatchingAlgorithm of testMatcher.cpp / cfg-3 the two are the same thing; the production
// VioKeyframeWindowMatchingAlgorithm gates distance() geometrically -- use candidates() below for that one).
class DenseMatcher {
 public:
  typedef float distance_t;
  struct Pairing {
    Pairing() : indexA(-1), distance(std::numeric_limits<float>::max()) {}
    Pairing(int ia, distance_t d) : indexA(ia), distance(d) {}
    bool operator<(const Pairing& rhs) const { return distance < rhs.distance; }
    int indexA;
    distance_t distance;
  };
  DenseMatcher(okb_ctx* ctx, unsigned char /*numMatcherThreads*/ = 8, unsigned char numBest = 4, bool useDistanceRatioThreshold = false)
      : ctx_(ctx), numBest_(numBest), useDistanceRatioThreshold_(useDistanceRatioThreshold) {}

  // DenseMatcher::match (implementation/DenseMatcher.hpp:48-122) for a matching algorithm whose distance() is the plain
  // descriptor distance.
  template <typename MATCHING_ALGORITHM_T>
  void match(MATCHING_ALGORITHM_T& matchingAlgorithm) {
    matchingAlgorithm.doSetup();
    matchImpl(matchingAlgorithm, nullptr);
  }

  // The same for a geometry-gated algorithm (VioKeyframeWindowMatchingAlgorithm: distance() = Hamming AND verifyMatch,
  // src/VioKeyframeWindowMatchingAlgorithm.cpp:304-339): the algorithm hands over what its doSetup() prepared --
  // projections + uncertainties (3D-2D) or bearings, ray sigmas and T_AB (2D-2D) -- as an okb_match_gate through
  // matchGate(); the gate runs on the device inside the top-k kernel (okb_hamming_match_gated).
  template <typename MATCHING_ALGORITHM_T>
  void matchGated(MATCHING_ALGORITHM_T& matchingAlgorithm) {
    matchingAlgorithm.doSetup();
    const okb_match_gate& gate = matchingAlgorithm.matchGate();
    matchImpl(matchingAlgorithm, &gate);
  }

  // Candidate lists for a geometry-gated matching algorithm: every B with Hamming distance < threshold per A, ascending
  // B (CSR).  The caller applies verifyMatch (VioKeyframeWindowMatchingAlgorithm.cpp:304-339) on these pairs only.
  void candidates(const uint8_t* A, int nA, const uint8_t* B, int nB, int descriptorBytes, float threshold, std::vector<uint32_t>& rowPtr,
                  std::vector<uint32_t>& colIdx, std::vector<uint16_t>& dist) {
    rowPtr.assign((size_t)nA + 1, 0);
    size_t cap = (size_t)std::max(1024, 8 * (nA + nB));
    for (int attempt = 0; attempt < 2; ++attempt) {
      colIdx.resize(cap); dist.resize(cap);
      const int rc = okb_hamming_candidates(ctx_, A, nA, B, nB, descriptorBytes, threshold, rowPtr.data(), colIdx.data(), dist.data(), (int)cap);
      if (rc == OKB_OK) { colIdx.resize(rowPtr[nA]); dist.resize(rowPtr[nA]); return; }
      if (rc != OKB_ERR_CAPACITY) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(ctx_));
      cap = rowPtr[nA];
    }
    throw std::runtime_error("okvis_b200: candidate capacity");
  }

 private:
  template <typename MATCHING_ALGORITHM_T>
  void matchImpl(MATCHING_ALGORITHM_T& matchingAlgorithm, const okb_match_gate* gate) {
    const int nA = (int)matchingAlgorithm.sizeA(), nB = (int)matchingAlgorithm.sizeB();
    if (nA == 0 || nB == 0) return;
    std::vector<uint8_t> skipA(nA), skipB(nB);
    for (int i = 0; i < nA; ++i) skipA[i] = matchingAlgorithm.skipA(i) ? 1 : 0;
    for (int i = 0; i < nB; ++i) skipB[i] = matchingAlgorithm.skipB(i) ? 1 : 0;
    std::vector<okb_pair> topk((size_t)nA * numBest_), pairs(nB);
    const distance_t thr = matchingAlgorithm.distanceThreshold();
    const int rc = gate ? okb_hamming_match_gated(ctx_, matchingAlgorithm.descriptorsA(), nA, matchingAlgorithm.descriptorsB(), nB,
                                                  matchingAlgorithm.descriptorBytes(), skipA.data(), skipB.data(), thr, numBest_, 0, 0.f, gate,
                                                  topk.data(), pairs.data())
                        : okb_hamming_match(ctx_, matchingAlgorithm.descriptorsA(), nA, matchingAlgorithm.descriptorsB(), nB,
                                            matchingAlgorithm.descriptorBytes(), skipA.data(), skipB.data(), thr, numBest_, 0, 0.f, topk.data(),
                                            pairs.data());
    if (rc != OKB_OK) throw std::runtime_error(std::string("okvis_b200: ") + okb_last_error(ctx_));
    // epilogue of matchBody (implementation/DenseMatcher.hpp:92-122), ascending B
    matchingAlgorithm.reserveMatches((size_t)nB);
    const distance_t ratio = matchingAlgorithm.distanceRatioThreshold();
    for (int b = 0; b < nB; ++b) {
      if (!(pairs[b].distance < thr)) continue;
      if (useDistanceRatioThreshold_) {
        const okb_pair* best = &topk[(size_t)pairs[b].index_a * numBest_];
        if (numBest_ > 1 && best[1].index_a != -1) {
          if (!(best[0].distance == 0 || best[1].distance / best[0].distance > ratio)) continue;
        }
      }
      matchingAlgorithm.setBestMatch((size_t)pairs[b].index_a, (size_t)b, pairs[b].distance);
    }
  }

  okb_ctx* ctx_;
  int numBest_;
  bool useDistanceRatioThreshold_;
};

}  // namespace okvis_b200
