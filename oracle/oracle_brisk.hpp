// TEST INFRASTRUCTURE ONLY -- CPU restatement of the frontend hot path:
//   Frontend::detectAndDescribe  okvis_frontend/src/Frontend.cpp:92-114
//   Frame::detect / describe     okvis_cv/include/okvis/implementation/Frame.hpp:109-156
// PARITY UNPINNED: the detector and descriptor live in brisk 2.0.5 (CMakeLists.txt:113-121), which is
// not in /root/reference and whose tests assert nothing (TestFrame.cpp:47-84).  What is restated
// here is this project's own, fully specified BRISK-style pipeline (DESIGN.md "Frontend spec"):
//   * Harris score on the full-resolution image (octaves = 0): 3x3 Scharr gradients, 5x5 binomial
//     window, integer arithmetic, score = (ab - c^2 - (a+b)^2/16) >> 12 clamped to int32;
//   * 3x3 strict non-maximum suppression, absolute threshold, 16 px border;
//   * uniformity: greedy acceptance in descending score order (ties: raster order) of keypoints
//     farther than `uniformity_radius` from every accepted one, at most max_keypoints;
//   * keypoint angle from the gravity direction exactly as Frame::describe (in-tree, restated
//     line by line): backProject (Gauss-Newton undistort, 5 iterations), project with Jacobian,
//     angle = atan2((J g)[1], (J g)[0]) in degrees;
//   * descriptor: 60-point BRISK ring pattern (radii {0,2.9,4.9,7.4,10.8}*0.85, {1,10,14,15,20}
//     points), box-smoothed samples from an integral image, the 8*desc_bytes shortest point pairs,
//     bit = mean_i > mean_j, rotation quantised to 1024 steps.
// The GPU kernels implement the same spec; GPU == this oracle bit for bit is what tests assert.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/okvis_b200.h"
#include "oracle_errors.hpp"

namespace oko {

constexpr int kBorder = 16;
constexpr int kRotBins = 1024;
constexpr int kPatternPoints = 60;

struct BriskPattern {
  double px[kPatternPoints], py[kPatternPoints];
  int half[kPatternPoints];
  std::vector<uint8_t> pair_i, pair_j;
  std::vector<int16_t> lut;   // [kRotBins][60][2]
};

inline BriskPattern make_pattern(int desc_bytes) {
  BriskPattern P;
  const double rList[5] = {0.0, 2.9 * 0.85, 4.9 * 0.85, 7.4 * 0.85, 10.8 * 0.85};
  const int nList[5] = {1, 10, 14, 15, 20};
  int idx = 0;
  for (int ring = 0; ring < 5; ++ring) {
    const double sigma = 1.3 * (ring == 0 ? rList[1] * std::sin(M_PI / nList[1]) : rList[ring] * std::sin(M_PI / nList[ring]));
    for (int a = 0; a < nList[ring]; ++a) {
      const double alpha = 2.0 * M_PI * a / nList[ring];
      P.px[idx] = rList[ring] * std::cos(alpha);
      P.py[idx] = rList[ring] * std::sin(alpha);
      P.half[idx] = std::max(1, (int)std::lround(sigma));
      ++idx;
    }
  }
  // squared distances quantised to 1e-6 (ordering of equal distances must not depend on FP contraction)
  struct Pr { long long d; int i, j; };
  std::vector<Pr> all;
  for (int i = 0; i < kPatternPoints; ++i)
    for (int j = i + 1; j < kPatternPoints; ++j) {
      const double dx = P.px[i] - P.px[j], dy = P.py[i] - P.py[j];
      all.push_back({std::llround((dx * dx + dy * dy) * 1e6), i, j});
    }
  std::stable_sort(all.begin(), all.end(), [](const Pr& a, const Pr& b) {
    if (a.d != b.d) return a.d < b.d;
    if (a.i != b.i) return a.i < b.i;
    return a.j < b.j;
  });
  const int nbits = 8 * desc_bytes;
  for (int k = 0; k < nbits; ++k) { P.pair_i.push_back((uint8_t)all[k].i); P.pair_j.push_back((uint8_t)all[k].j); }
  P.lut.resize((size_t)kRotBins * kPatternPoints * 2);
  for (int r = 0; r < kRotBins; ++r) {
    const double ang = 2.0 * M_PI * r / kRotBins, c = std::cos(ang), s = std::sin(ang);
    for (int p = 0; p < kPatternPoints; ++p) {
      P.lut[((size_t)r * kPatternPoints + p) * 2 + 0] = (int16_t)std::lround(P.px[p] * c - P.py[p] * s);
      P.lut[((size_t)r * kPatternPoints + p) * 2 + 1] = (int16_t)std::lround(P.px[p] * s + P.py[p] * c);
    }
  }
  return P;
}

// integer Harris score image (int32), zero outside the valid interior (3 px)
inline void harris_score(const uint8_t* img, int W, int H, int stride, std::vector<int32_t>& score) {
  std::vector<int32_t> xx((size_t)W * H, 0), yy((size_t)W * H, 0), xy((size_t)W * H, 0);
  auto I = [&](int y, int x) { return (int)img[(size_t)y * stride + x]; };
  for (int y = 1; y < H - 1; ++y)
    for (int x = 1; x < W - 1; ++x) {
      const int gx = 3 * (I(y - 1, x + 1) - I(y - 1, x - 1)) + 10 * (I(y, x + 1) - I(y, x - 1)) + 3 * (I(y + 1, x + 1) - I(y + 1, x - 1));
      const int gy = 3 * (I(y + 1, x - 1) - I(y - 1, x - 1)) + 10 * (I(y + 1, x) - I(y - 1, x)) + 3 * (I(y + 1, x + 1) - I(y - 1, x + 1));
      xx[(size_t)y * W + x] = gx * gx; yy[(size_t)y * W + x] = gy * gy; xy[(size_t)y * W + x] = gx * gy;
    }
  score.assign((size_t)W * H, 0);
  static const int w5[5] = {1, 4, 6, 4, 1};
  for (int y = 3; y < H - 3; ++y)
    for (int x = 3; x < W - 3; ++x) {
      int64_t a = 0, b = 0, c = 0;
      for (int dy = -2; dy <= 2; ++dy)
        for (int dx = -2; dx <= 2; ++dx) {
          const int64_t wgt = w5[dy + 2] * w5[dx + 2];
          const size_t o = (size_t)(y + dy) * W + x + dx;
          a += wgt * xx[o]; b += wgt * yy[o]; c += wgt * xy[o];
        }
      a >>= 8; b >>= 8; c >>= 8;   // arithmetic shift (c may be negative): floor division by 256
      int64_t s = (a * b - c * c) - (((a + b) * (a + b)) >> 4);
      s >>= 12;
      if (s > 2147483647LL) s = 2147483647LL;
      if (s < -2147483647LL) s = -2147483647LL;
      score[(size_t)y * W + x] = (int32_t)s;
    }
}

// RadialTangentialDistortion::undistort (RadialTangentialDistortion.hpp(impl):210-249), used for all
// models in this spec; returns the undistorted normalised point.
inline void undistort_gn(const okb_camera& cam, const double* pd, double* pu) {
  double x[2] = {pd[0], pd[1]};
  for (int i = 0; i < 5; ++i) {
    double xt[2], E[4];
    distort(cam, x, xt, E);
    const double e[2] = {pd[0] - xt[0], pd[1] - xt[1]};
    // du = (E^T E)^-1 E^T e
    const double a = E[0] * E[0] + E[2] * E[2], b = E[0] * E[1] + E[2] * E[3], d = E[1] * E[1] + E[3] * E[3];
    const double r0 = E[0] * e[0] + E[2] * e[1], r1 = E[1] * e[0] + E[3] * e[1];
    const double det = a * d - b * b;
    x[0] += (d * r0 - b * r1) / det;
    x[1] += (-b * r0 + a * r1) / det;
    const double chi2 = e[0] * e[0] + e[1] * e[1];
    if (chi2 < 1e-15) break;
  }
  pu[0] = x[0]; pu[1] = x[1];
}

// Frame::describe orientation (Frame.hpp(impl):128-151)
inline float gravity_angle_deg(const okb_camera& cam, double kx, double ky, const double* g_C) {
  const double pd[2] = {(kx - cam.cu) / cam.fu, (ky - cam.cv) / cam.fv};
  double pu[2];
  undistort_gn(cam, pd, pu);
  const double ep[3] = {pu[0], pu[1], 1.0};
  double ip[2], J[6];
  project(cam, ep, ip, J);
  const double e0 = J[0] * g_C[0] + J[1] * g_C[1] + J[2] * g_C[2];
  const double e1 = J[3] * g_C[0] + J[4] * g_C[1] + J[5] * g_C[2];
  return (float)(std::atan2(e1, e0) / M_PI * 180.0);
}

inline int detect_describe(const uint8_t* img, int W, int H, int stride, const okb_camera& cam, const double* R_CW,
                           const okb_detect_params& prm, okb_keypoint* kps, uint8_t* desc, int max_out) {
  std::vector<int32_t> score;
  harris_score(img, W, H, stride, score);
  struct Cand { int32_t s; int idx; };
  std::vector<Cand> cands;
  const int32_t thr = (int32_t)std::ceil(prm.absolute_threshold);
  for (int y = kBorder; y < H - kBorder; ++y)
    for (int x = kBorder; x < W - kBorder; ++x) {
      const int32_t s = score[(size_t)y * W + x];
      if (s < thr) continue;
      bool mx = true;
      for (int dy = -1; dy <= 1 && mx; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if (!dy && !dx) continue;
          if (!(s > score[(size_t)(y + dy) * W + x + dx])) { mx = false; break; }
        }
      if (mx) cands.push_back({s, y * W + x});
    }
  std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.s != b.s ? a.s > b.s : a.idx < b.idx; });
  const double r2 = prm.uniformity_radius * prm.uniformity_radius;
  const int maxk = std::min(prm.max_keypoints, max_out);
  std::vector<int> acc;
  for (const Cand& c : cands) {
    if ((int)acc.size() >= maxk) break;
    const int x = c.idx % W, y = c.idx / W;
    bool ok = true;
    for (int j : acc) {
      const double dx = x - j % W, dy = y - j / W;
      if (dx * dx + dy * dy < r2) { ok = false; break; }
    }
    if (ok) {
      kps[acc.size()] = okb_keypoint{(float)x, (float)y, 12.0f, 0.0f, (float)c.s, 0};
      acc.push_back(c.idx);
    }
  }
  const int n = (int)acc.size();
  // describe
  const BriskPattern P = make_pattern(prm.desc_bytes);
  std::vector<uint32_t> II((size_t)(W + 1) * (H + 1), 0);
  for (int y = 0; y < H; ++y) {
    uint32_t row = 0;
    for (int x = 0; x < W; ++x) {
      row += img[(size_t)y * stride + x];
      II[(size_t)(y + 1) * (W + 1) + x + 1] = II[(size_t)y * (W + 1) + x + 1] + row;
    }
  }
  const double g_W[3] = {0, 0, -1};
  double g_C[3];
  matmul(R_CW, g_W, g_C, 3, 3, 1);   // extractionDirection = R_CW * (0,0,-1)  (Frontend.cpp:108-109)
  for (int k = 0; k < n; ++k) {
    const int x = acc[k] % W, y = acc[k] / W;
    float ang = 0.0f;
    if (prm.rotation_invariance) ang = gravity_angle_deg(cam, x, y, g_C);
    kps[k].angle = ang;
    const int bin = ((int)std::lround((double)ang / 360.0 * kRotBins)) & (kRotBins - 1);
    uint32_t S[kPatternPoints], area[kPatternPoints];
    for (int p = 0; p < kPatternPoints; ++p) {
      const int cx = x + P.lut[((size_t)bin * kPatternPoints + p) * 2], cy = y + P.lut[((size_t)bin * kPatternPoints + p) * 2 + 1];
      const int h = P.half[p];
      const int x0 = std::max(cx - h, 0), x1 = std::min(cx + h, W - 1), y0 = std::max(cy - h, 0), y1 = std::min(cy + h, H - 1);
      S[p] = II[(size_t)(y1 + 1) * (W + 1) + x1 + 1] - II[(size_t)y0 * (W + 1) + x1 + 1] - II[(size_t)(y1 + 1) * (W + 1) + x0] + II[(size_t)y0 * (W + 1) + x0];
      area[p] = (uint32_t)((x1 - x0 + 1) * (y1 - y0 + 1));
    }
    uint8_t* d = desc + (size_t)k * prm.desc_bytes;
    for (int b = 0; b < prm.desc_bytes; ++b) d[b] = 0;
    for (int b = 0; b < 8 * prm.desc_bytes; ++b) {
      const int i = P.pair_i[b], j = P.pair_j[b];
      if ((uint64_t)S[i] * area[j] > (uint64_t)S[j] * area[i]) d[b >> 3] |= (uint8_t)(1u << (b & 7));
    }
  }
  return n;
}

}  // namespace oko
