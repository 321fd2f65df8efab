// TEST INFRASTRUCTURE ONLY -- CPU restatement of what Estimator::optimize() does through Ceres 1.9:
//   okvis_ceres/src/Estimator.cpp:843-906 (options, Map::solve, per-landmark quality)
//   Ceres 1.9.0 (tag 7c57de50, NOT in /root/reference): trust_region_minimizer.cc, dogleg_strategy.cc
//   (TRADITIONAL_DOGLEG), schur_eliminator_impl.h, corrector.cc, residual_block.cc -- restated from
//   the published algorithm; the in-tree copy of the corrector is MarginalizationError.cpp:325-365.
// PARITY UNPINNED: no reference test asserts a cost value, iteration count or step vector
// (SURVEY.md 8c); convergence-level behaviour is what the reference tests check and what
// tests/ check against this file.
#pragma once
#include <cstdint>
#include <vector>

#include "../include/okvis_b200.h"
#include "oracle_errors.hpp"

namespace oko {

struct IterationRecord {
  double cost;            // cost after this iteration (accepted) or unchanged (rejected)
  double cost_change;
  double radius;
  double step_norm;
  double relative_decrease;
  int accepted;           // 1 accepted, 0 rejected, -1 invalid
};

struct PhaseTimes {  // seconds, accumulated over one solve()
  double evaluate_jac = 0, schur = 0, reduced_solve = 0, backsub = 0, evaluate_cost = 0, quality = 0, other = 0;
};

class Problem {
 public:
  explicit Problem(const okb_window_desc& d);
  // Ceres-1.9-like solve.  Returns the summary; `trace` receives one record per iteration.
  okb_summary solve(const okb_solve_options& opt, std::vector<IterationRecord>* trace, int num_threads);
  // Estimator.cpp:880-900 / Map.cpp:101-156
  void landmark_quality(std::vector<double>& quality) const;
  // cost at the current state (0.5 * sum rho(s)), no side effects on the IMU caches if `peek`.
  double cost_only();

  int K, NSB, NE, L, d, n_tan;
  std::vector<double> poses, sb, ext, lms;
  std::vector<uint8_t> ext_fixed;
  PhaseTimes times;
  int num_threads = 1;

  // ---- graph (deep copies)
  std::vector<okb_camera> cams;
  std::vector<okb_observation> obs;
  std::vector<okb_imu_term> imu_terms;
  std::vector<okb_imu_sample> samples;
  okb_imu_params imu_params;
  std::vector<okb_pose_prior> pose_priors;
  std::vector<okb_sb_prior> sb_priors;
  std::vector<okb_relpose_term> relpose;
  bool has_marg = false;
  okb_marg_prior marg;
  std::vector<int32_t> marg_kind;
  std::vector<uint32_t> marg_idx;
  std::vector<double> marg_x0, marg_J, marg_e0;
  std::vector<uint8_t> marg_fixed;
  bool use_cauchy = true;
  std::vector<ImuCache> imu_cache;

  // tangent-space layout: [poses 6 | free extrinsics 6 | speed/bias 9 | landmarks 3]
  std::vector<int> pose_off, ext_off, sb_off;
  int lm_off(int l) const { return d + 3 * l; }

  // ---- block-sparse Jacobian
  struct JBlock { int col, w, data; };
  struct RBlock { int row, m, b0, nb; };
  std::vector<RBlock> rblocks;   // [obs..., imu..., pose priors..., sb priors..., relpose..., marg]
  std::vector<JBlock> jblocks;
  int n_rows = 0;
  int n_jvals = 0;
  std::vector<int> lm_ptr, lm_obs;  // landmark -> observation CSR

  void build_structure();
  // Evaluates all residual blocks at state (P,S,E,M).  r, Jv may be null (cost only).
  double evaluate(const double* P, const double* S, const double* E, const double* M, double* r, double* Jv);
  void plus(const double* delta, std::vector<double>& P, std::vector<double>& S, std::vector<double>& E,
            std::vector<double>& M) const;
  // Schur solve of (J^T J + diag(D)^2) y = J^T r; false on Cholesky failure / non-finite result.
  bool schur_solve(const double* Jv, const double* r, const double* D, double* y);
};

}  // namespace oko
