// TEST INFRASTRUCTURE ONLY -- CPU restatement of one marginalisation step on an oracle Problem, in the reference's
// literal order of operations (dense H including one 3x3 block per marginalised landmark appended after the dense
// blocks, explicit diagonal preconditioning of the whole system per stage):
//   MarginalizationError::addResidualBlock        okvis_ceres/src/MarginalizationError.cpp:127-435
//       residuals are evaluated at the linearisation points (first-estimate Jacobians, :292-310), loss-corrected
//       (:313-365), H += J^T J, b0 -= J^T r (:367-420)
//   MarginalizationError::marginalizeOut          :507-802   (landmark stage, then dense stage; oracle_marg.hpp)
//   MarginalizationError::updateErrorComputation  :806-846
// The job format is the product's okb_marg_job (include/okvis_b200.h); which blocks / residuals go into a job is the
// bookkeeping of Estimator::applyMarginalizationStrategy (Estimator.cpp:434-773) and is done by the caller.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle_marg.hpp"
#include "oracle_solver.hpp"

namespace oko {

struct MargOut {
  int n = 0, rank = 0;
  std::vector<int32_t> kind;
  std::vector<uint32_t> idx;
  std::vector<double> x0, J, e0, H, b0;
};

// H_prev / b0_prev: the H_ / b0_ members kept from the previous call (null: J^T J, -J^T e0 of the problem's prior).
inline int marginalize_problem(Problem& P, const okb_marg_job& job, const double* H_prev, const double* b0_prev, MargOut& out) {
  const int NB = job.n_blocks;
  std::vector<int> col(NB + 1, 0), dim(NB);
  std::vector<int> pose_blk(P.K, -1), sb_blk(P.NSB, -1);
  for (int b = 0; b < NB; ++b) {
    dim[b] = block_min_dim(job.block_kind[b]);
    col[b + 1] = col[b] + dim[b];
    (job.block_kind[b] == OKB_BLOCK_POSE ? pose_blk : sb_blk)[job.block_idx[b]] = b;
  }
  const int Nd = col[NB], NL = job.n_landmarks, N = Nd + 3 * NL;
  // linearisation points
  std::vector<double> xlin((size_t)9 * NB, 0.0);
  std::vector<int> old_off(P.marg_kind.size() + 1, 0), old_col(P.marg_kind.size() + 1, 0);
  for (size_t i = 0; i < P.marg_kind.size(); ++i) {
    old_off[i + 1] = old_off[i] + block_dim(P.marg_kind[i]);
    old_col[i + 1] = old_col[i] + (P.marg_fixed[i] ? 0 : block_min_dim(P.marg_kind[i]));
  }
  for (int b = 0; b < NB; ++b) {
    const int pv = job.block_prev[b], w = block_dim(job.block_kind[b]);
    const double* src = pv >= 0 ? P.marg_x0.data() + old_off[pv]
                                : (job.block_kind[b] == OKB_BLOCK_POSE ? P.poses.data() + 7 * job.block_idx[b] : P.sb.data() + 9 * job.block_idx[b]);
    std::memcpy(&xlin[(size_t)9 * b], src, sizeof(double) * w);
  }
  std::vector<double> H((size_t)N * N, 0.0), b0(N, 0.0);
  // the current prior in the new ordering
  if (P.has_marg) {
    const int n_old = P.marg.n;
    std::vector<double> Ho, bo(n_old, 0.0);
    if (H_prev) { Ho.assign(H_prev, H_prev + (size_t)n_old * n_old); bo.assign(b0_prev, b0_prev + n_old); }
    else {
      Ho.assign((size_t)n_old * n_old, 0.0);
      for (int i = 0; i < n_old; ++i) {
        for (int j = 0; j < n_old; ++j) { double s = 0; for (int r = 0; r < n_old; ++r) s += P.marg_J[(size_t)r * n_old + i] * P.marg_J[(size_t)r * n_old + j]; Ho[(size_t)i * n_old + j] = s; }
        double s = 0; for (int r = 0; r < n_old; ++r) s += P.marg_J[(size_t)r * n_old + i] * P.marg_e0[r];
        bo[i] = -s;
      }
    }
    for (int bi = 0; bi < NB; ++bi) {
      if (job.block_prev[bi] < 0) continue;
      for (int a = 0; a < dim[bi]; ++a) {
        const int ro = old_col[job.block_prev[bi]] + a;
        b0[col[bi] + a] = bo[ro];
        for (int bj = 0; bj < NB; ++bj) {
          if (job.block_prev[bj] < 0) continue;
          for (int c = 0; c < dim[bj]; ++c) H[(size_t)(col[bi] + a) * N + col[bj] + c] = Ho[(size_t)ro * n_old + old_col[job.block_prev[bj]] + c];
        }
      }
    }
  }
  auto add = [&](int m, int nblk, const int* offs, const int* widths, const double* const* Jm, const double* r) {
    for (int i = 0; i < nblk; ++i)
      for (int a = 0; a < widths[i]; ++a) {
        double s = 0; for (int k = 0; k < m; ++k) s += Jm[i][k * widths[i] + a] * r[k];
        b0[offs[i] + a] -= s;
        for (int j = 0; j < nblk; ++j)
          for (int c = 0; c < widths[j]; ++c) {
            double h = 0; for (int k = 0; k < m; ++k) h += Jm[i][k * widths[i] + a] * Jm[j][k * widths[j] + c];
            H[(size_t)(offs[i] + a) * N + offs[j] + c] += h;
          }
      }
  };
  for (int i = 0; i < job.n_sb_priors; ++i) {
    const okb_sb_prior& pr = P.sb_priors[job.sb_priors[i]];
    const int blk = sb_blk[pr.sb_idx];
    if (blk < 0) return -1;
    double r[9], Jm[81];
    speed_bias_error(pr.meas, pr.sqrt_info, &xlin[(size_t)9 * blk], r, Jm);
    const int offs[1] = {col[blk]}, widths[1] = {9};
    const double* Js[1] = {Jm};
    add(9, 1, offs, widths, Js, r);
  }
  for (int i = 0; i < job.n_imu_terms; ++i) {
    const int t = (int)job.imu_terms[i];
    const okb_imu_term& T = P.imu_terms[t];
    const int blks[4] = {pose_blk[T.pose0], sb_blk[T.sb0], pose_blk[T.pose1], sb_blk[T.sb1]};
    for (int k = 0; k < 4; ++k) if (blks[k] < 0) return -1;
    double r[15], J0[90], J1[135], J2[90], J3[135];
    imu_error(P.samples.data() + T.sample_offset, (int)T.sample_count, P.imu_params, T.t0_ns, T.t1_ns, &xlin[(size_t)9 * blks[0]], &xlin[(size_t)9 * blks[1]],
              &xlin[(size_t)9 * blks[2]], &xlin[(size_t)9 * blks[3]], P.imu_cache[t], r, J0, J1, J2, J3);
    const int offs[4] = {col[blks[0]], col[blks[1]], col[blks[2]], col[blks[3]]}, widths[4] = {6, 9, 6, 9};
    const double* Js[4] = {J0, J1, J2, J3};
    add(15, 4, offs, widths, Js, r);
  }
  for (int j = 0; j < NL; ++j) {
    const uint32_t l = job.landmarks[j];
    for (const okb_observation& ob : P.obs) {
      if (ob.lm_idx != l) continue;
      const int blk = pose_blk[ob.pose_idx];
      if (blk < 0) return -1;
      double r[2], J0[12], J1[6];
      reprojection_error(P.cams[ob.cam_idx], &xlin[(size_t)9 * blk], P.lms.data() + 4 * (size_t)l, P.ext.data() + 7 * ob.ext_idx, ob.z, ob.sqrt_info, r, J0, J1, nullptr);
      if (P.use_cauchy) {       // Corrector as restated at MarginalizationError.cpp:325-365 with CauchyLoss(1): rho'' < 0 branch
        const double sq = r[0] * r[0] + r[1] * r[1];
        const double rho1 = 1.0 / (1.0 + sq), rho2 = -rho1 * rho1;
        const double sqrt_rho1 = std::sqrt(rho1);
        double residual_scaling, alpha_sq_norm;
        if (sq == 0.0 || rho2 <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
        else { const double D = 1.0 + 2.0 * sq * rho2 / rho1; const double alpha = 1.0 - std::sqrt(D); residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / sq; }
        auto correct = [&](double* Jm, int w) {
          for (int c = 0; c < w; ++c) {
            const double rtJ = r[0] * Jm[c] + r[1] * Jm[w + c];
            Jm[c] = sqrt_rho1 * (Jm[c] - alpha_sq_norm * r[0] * rtJ);
            Jm[w + c] = sqrt_rho1 * (Jm[w + c] - alpha_sq_norm * r[1] * rtJ);
          }
        };
        correct(J0, 6); correct(J1, 3);
        r[0] *= residual_scaling; r[1] *= residual_scaling;
      }
      const int offs[2] = {col[blk], Nd + 3 * j}, widths[2] = {6, 3};
      const double* Js[2] = {J0, J1};
      add(2, 2, offs, widths, Js, r);
    }
  }
  // marginalizeOut: landmark stage, then dense stage
  int n = N;
  if (NL > 0) {
    std::vector<std::pair<int, int>> rl{{Nd, 3 * NL}};
    marginalize_stage(H, b0, n, rl, true);
  }
  std::vector<std::pair<int, int>> rd;
  for (int b = 0; b < NB; ++b)
    if (job.block_marginalize[b]) {
      if (!rd.empty() && rd.back().first + rd.back().second == col[b]) rd.back().second += dim[b];
      else rd.emplace_back(col[b], dim[b]);
    }
  if (!rd.empty()) marginalize_stage(H, b0, n, rd, false);
  out.n = n; out.H = H; out.b0 = b0;
  out.H.resize((size_t)n * n); out.b0.resize(n);
  out.rank = n ? marg_update_error_computation(out.H, out.b0, n, out.J, out.e0) : 0;
  out.kind.clear(); out.idx.clear(); out.x0.clear();
  for (int b = 0; b < NB; ++b) {
    if (job.block_marginalize[b]) continue;
    out.kind.push_back(job.block_kind[b]); out.idx.push_back(job.block_idx[b]);
    out.x0.insert(out.x0.end(), &xlin[(size_t)9 * b], &xlin[(size_t)9 * b] + block_dim(job.block_kind[b]));
  }
  return n;
}

}  // namespace oko
