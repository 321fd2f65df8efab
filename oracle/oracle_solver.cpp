// TEST INFRASTRUCTURE ONLY -- see oracle_solver.hpp for scope and the parity-unpinned notice.
#include "oracle_solver.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oko {

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Problem::Problem(const okb_window_desc& w) {
  K = w.n_poses; NSB = w.n_speed_bias; NE = w.n_extrinsics; L = w.n_landmarks;
  poses.assign(w.poses, w.poses + 7 * K);
  sb.assign(w.speed_bias, w.speed_bias + 9 * NSB);
  ext.assign(w.extrinsics, w.extrinsics + 7 * NE);
  ext_fixed.assign(NE, 1);
  if (w.extrinsics_fixed) ext_fixed.assign(w.extrinsics_fixed, w.extrinsics_fixed + NE);
  lms.assign(w.landmarks, w.landmarks + 4 * L);
  cams.assign(w.cameras, w.cameras + w.n_cameras);
  obs.assign(w.obs, w.obs + w.n_obs);
  imu_terms.assign(w.imu_terms, w.imu_terms + w.n_imu_terms);
  samples.assign(w.imu_samples, w.imu_samples + w.n_imu_samples);
  imu_params = w.imu_params;
  pose_priors.assign(w.pose_priors, w.pose_priors + w.n_pose_priors);
  sb_priors.assign(w.sb_priors, w.sb_priors + w.n_sb_priors);
  relpose.assign(w.relpose_terms, w.relpose_terms + w.n_relpose_terms);
  if (w.marg && w.marg->n > 0) {
    has_marg = true;
    const okb_marg_prior& m = *w.marg;
    marg_kind.assign(m.block_kind, m.block_kind + m.n_blocks);
    marg_idx.assign(m.block_idx, m.block_idx + m.n_blocks);
    int xdim = 0;
    for (int i = 0; i < m.n_blocks; ++i) xdim += block_dim(m.block_kind[i]);
    marg_x0.assign(m.x0, m.x0 + xdim);
    marg_J.assign(m.J, m.J + (size_t)m.n * m.n);
    marg_e0.assign(m.e0, m.e0 + m.n);
    marg_fixed.assign(m.n_blocks, 0);
    for (int i = 0; i < m.n_blocks; ++i)
      if (marg_kind[i] == OKB_BLOCK_EXTRINSICS && ext_fixed[marg_idx[i]]) marg_fixed[i] = 1;
    marg.n = m.n; marg.n_blocks = m.n_blocks;
    marg.block_kind = marg_kind.data(); marg.block_idx = marg_idx.data();
    marg.x0 = marg_x0.data(); marg.J = marg_J.data(); marg.e0 = marg_e0.data();
  }
  imu_cache.resize(imu_terms.size());
  build_structure();
}

void Problem::build_structure() {
  pose_off.resize(K); ext_off.assign(NE, -1); sb_off.resize(NSB);
  int o = 0;
  for (int k = 0; k < K; ++k) { pose_off[k] = o; o += 6; }
  for (int e = 0; e < NE; ++e) if (!ext_fixed[e]) { ext_off[e] = o; o += 6; }
  for (int j = 0; j < NSB; ++j) { sb_off[j] = o; o += 9; }
  d = o;
  n_tan = d + 3 * L;
  rblocks.clear(); jblocks.clear();
  int row = 0, data = 0;
  auto add_block = [&](int col, int w, int m) { jblocks.push_back({col, w, data}); data += m * w; };
  for (size_t i = 0; i < obs.size(); ++i) {
    RBlock rb{row, 2, (int)jblocks.size(), 0};
    add_block(pose_off[obs[i].pose_idx], 6, 2);
    add_block(lm_off(obs[i].lm_idx), 3, 2);
    if (ext_off[obs[i].ext_idx] >= 0) add_block(ext_off[obs[i].ext_idx], 6, 2);
    rb.nb = (int)jblocks.size() - rb.b0;
    rblocks.push_back(rb); row += 2;
  }
  for (size_t i = 0; i < imu_terms.size(); ++i) {
    RBlock rb{row, 15, (int)jblocks.size(), 4};
    add_block(pose_off[imu_terms[i].pose0], 6, 15);
    add_block(sb_off[imu_terms[i].sb0], 9, 15);
    add_block(pose_off[imu_terms[i].pose1], 6, 15);
    add_block(sb_off[imu_terms[i].sb1], 9, 15);
    rblocks.push_back(rb); row += 15;
  }
  for (size_t i = 0; i < pose_priors.size(); ++i) {
    RBlock rb{row, 6, (int)jblocks.size(), 1};
    add_block(pose_off[pose_priors[i].pose_idx], 6, 6);
    rblocks.push_back(rb); row += 6;
  }
  for (size_t i = 0; i < sb_priors.size(); ++i) {
    RBlock rb{row, 9, (int)jblocks.size(), 1};
    add_block(sb_off[sb_priors[i].sb_idx], 9, 9);
    rblocks.push_back(rb); row += 9;
  }
  for (size_t i = 0; i < relpose.size(); ++i) {
    RBlock rb{row, 6, (int)jblocks.size(), 0};
    if (ext_off[relpose[i].ext0] >= 0) add_block(ext_off[relpose[i].ext0], 6, 6);
    if (ext_off[relpose[i].ext1] >= 0) add_block(ext_off[relpose[i].ext1], 6, 6);
    rb.nb = (int)jblocks.size() - rb.b0;
    rblocks.push_back(rb); row += 6;
  }
  if (has_marg) {
    RBlock rb{row, marg.n, (int)jblocks.size(), 0};
    for (int i = 0; i < marg.n_blocks; ++i) {
      if (marg_fixed[i]) continue;
      const int kind = marg_kind[i];
      const int off = kind == OKB_BLOCK_POSE ? pose_off[marg_idx[i]]
                    : kind == OKB_BLOCK_SPEED_BIAS ? sb_off[marg_idx[i]] : ext_off[marg_idx[i]];
      add_block(off, block_min_dim(kind), marg.n);
    }
    rb.nb = (int)jblocks.size() - rb.b0;
    rblocks.push_back(rb); row += marg.n;
  }
  n_rows = row; n_jvals = data;
  // landmark -> obs CSR
  lm_ptr.assign(L + 1, 0);
  for (auto& ob : obs) lm_ptr[ob.lm_idx + 1]++;
  for (int l = 0; l < L; ++l) lm_ptr[l + 1] += lm_ptr[l];
  lm_obs.resize(obs.size());
  std::vector<int> fill(lm_ptr.begin(), lm_ptr.end() - 1);
  for (size_t i = 0; i < obs.size(); ++i) lm_obs[fill[obs[i].lm_idx]++] = (int)i;
}

// ResidualBlock::Evaluate + Corrector (Ceres 1.9 residual_block.cc / corrector.cc; in-tree mirror
// MarginalizationError.cpp:325-365).  For CauchyLoss(1) rho'' < 0 => residual and Jacobian scaled by
// sqrt(rho'), cost contribution 0.5*rho(s).
double Problem::evaluate(const double* P, const double* S, const double* E, const double* M, double* r, double* Jv) {
  double cost = 0.0;
  const int nobs = (int)obs.size();
  const bool wantJ = (Jv != nullptr);
  std::vector<double> rtmp;
  if (!r) { rtmp.resize(n_rows); r = rtmp.data(); }
#pragma omp parallel for reduction(+ : cost) schedule(static) num_threads(num_threads)
  for (int i = 0; i < nobs; ++i) {
    const okb_observation& ob = obs[i];
    const RBlock& rb = rblocks[i];
    double res[2], J0[12], J1[6], J2[12];
    const bool extFree = rb.nb == 3;
    reprojection_error(cams[ob.cam_idx], P + 7 * ob.pose_idx, M + 4 * ob.lm_idx, E + 7 * ob.ext_idx, ob.z,
                       ob.sqrt_info, res, wantJ ? J0 : nullptr, wantJ ? J1 : nullptr,
                       (wantJ && extFree) ? J2 : nullptr);
    const double sq = res[0] * res[0] + res[1] * res[1];
    double scale = 1.0;
    if (use_cauchy) {
      const double sum = 1.0 + sq;
      const double inv = 1.0 / sum;
      const double rho1 = std::max(std::numeric_limits<double>::min(), inv);
      cost += 0.5 * std::log(sum);
      scale = std::sqrt(rho1);
    } else {
      cost += 0.5 * sq;
    }
    r[rb.row] = scale * res[0]; r[rb.row + 1] = scale * res[1];
    if (wantJ) {
      double* d0 = Jv + jblocks[rb.b0].data;
      double* d1 = Jv + jblocks[rb.b0 + 1].data;
      for (int k = 0; k < 12; ++k) d0[k] = scale * J0[k];
      for (int k = 0; k < 6; ++k) d1[k] = scale * J1[k];
      if (extFree) {
        double* d2 = Jv + jblocks[rb.b0 + 2].data;
        for (int k = 0; k < 12; ++k) d2[k] = scale * J2[k];
      }
    }
  }
  int bi = nobs;
  for (size_t i = 0; i < imu_terms.size(); ++i, ++bi) {
    const okb_imu_term& t = imu_terms[i];
    const RBlock& rb = rblocks[bi];
    double* j[4] = {nullptr, nullptr, nullptr, nullptr};
    if (wantJ) for (int b = 0; b < 4; ++b) j[b] = Jv + jblocks[rb.b0 + b].data;
    imu_error(samples.data() + t.sample_offset, t.sample_count, imu_params, t.t0_ns, t.t1_ns, P + 7 * t.pose0,
              S + 9 * t.sb0, P + 7 * t.pose1, S + 9 * t.sb1, imu_cache[i], r + rb.row, j[0], j[1], j[2], j[3]);
    double sq = 0; for (int k = 0; k < 15; ++k) sq += r[rb.row + k] * r[rb.row + k];
    cost += 0.5 * sq;
  }
  for (size_t i = 0; i < pose_priors.size(); ++i, ++bi) {
    const RBlock& rb = rblocks[bi];
    pose_error(pose_priors[i].meas, pose_priors[i].sqrt_info, P + 7 * pose_priors[i].pose_idx, r + rb.row,
               wantJ ? Jv + jblocks[rb.b0].data : nullptr);
    double sq = 0; for (int k = 0; k < 6; ++k) sq += r[rb.row + k] * r[rb.row + k];
    cost += 0.5 * sq;
  }
  for (size_t i = 0; i < sb_priors.size(); ++i, ++bi) {
    const RBlock& rb = rblocks[bi];
    speed_bias_error(sb_priors[i].meas, sb_priors[i].sqrt_info, S + 9 * sb_priors[i].sb_idx, r + rb.row,
                     wantJ ? Jv + jblocks[rb.b0].data : nullptr);
    double sq = 0; for (int k = 0; k < 9; ++k) sq += r[rb.row + k] * r[rb.row + k];
    cost += 0.5 * sq;
  }
  for (size_t i = 0; i < relpose.size(); ++i, ++bi) {
    const RBlock& rb = rblocks[bi];
    double J0[36], J1[36];
    relative_pose_error(relpose[i].sqrt_info, E + 7 * relpose[i].ext0, E + 7 * relpose[i].ext1, r + rb.row, J0, J1);
    if (rb.nb == 0) continue;  // all-constant block: Ceres drops it (fixed cost)
    if (wantJ) {
      int b = rb.b0;
      if (ext_off[relpose[i].ext0] >= 0) std::memcpy(Jv + jblocks[b++].data, J0, sizeof J0);
      if (ext_off[relpose[i].ext1] >= 0) std::memcpy(Jv + jblocks[b++].data, J1, sizeof J1);
    }
    double sq = 0; for (int k = 0; k < 6; ++k) sq += r[rb.row + k] * r[rb.row + k];
    cost += 0.5 * sq;
  }
  if (has_marg) {
    const RBlock& rb = rblocks[bi];
    std::vector<double> x;
    for (int i = 0; i < marg.n_blocks; ++i) {
      const int kind = marg_kind[i];
      const double* src = kind == OKB_BLOCK_POSE ? P + 7 * marg_idx[i]
                        : kind == OKB_BLOCK_SPEED_BIAS ? S + 9 * marg_idx[i] : E + 7 * marg_idx[i];
      x.insert(x.end(), src, src + block_dim(kind));
    }
    std::vector<double> Jeff;
    if (wantJ) Jeff.resize((size_t)marg.n * marg.n);
    marginalization_error(marg, marg_fixed.data(), x.data(), r + rb.row, wantJ ? Jeff.data() : nullptr);
    if (wantJ) {
      int col = 0;
      for (int b = 0; b < rb.nb; ++b) {
        const JBlock& jb = jblocks[rb.b0 + b];
        for (int rr = 0; rr < marg.n; ++rr)
          for (int cc = 0; cc < jb.w; ++cc) Jv[jb.data + rr * jb.w + cc] = Jeff[(size_t)rr * marg.n + col + cc];
        col += jb.w;
      }
    }
    double sq = 0; for (int k = 0; k < marg.n; ++k) sq += r[rb.row + k] * r[rb.row + k];
    cost += 0.5 * sq;
  }
  return cost;
}

double Problem::cost_only() {
  return evaluate(poses.data(), sb.data(), ext.data(), lms.data(), nullptr, nullptr);
}

// ProgramEvaluator::Plus: block-wise LocalParameterization::Plus
// (PoseLocalParameterization.cpp:60-87, HomogeneousPointLocalParameterization.cpp:59-72).
void Problem::plus(const double* delta, std::vector<double>& P, std::vector<double>& S, std::vector<double>& E,
                   std::vector<double>& M) const {
  P.resize(poses.size()); S.resize(sb.size()); E = ext; M.resize(lms.size());
  for (int k = 0; k < K; ++k) pose_plus(&poses[7 * k], delta + pose_off[k], &P[7 * k]);
  for (int e = 0; e < NE; ++e) if (ext_off[e] >= 0) pose_plus(&ext[7 * e], delta + ext_off[e], &E[7 * e]);
  for (int j = 0; j < NSB; ++j) for (int c = 0; c < 9; ++c) S[9 * j + c] = sb[9 * j + c] + delta[sb_off[j] + c];
  for (int l = 0; l < L; ++l) {
    for (int c = 0; c < 3; ++c) M[4 * l + c] = lms[4 * l + c] + delta[lm_off(l) + c];
    M[4 * l + 3] = lms[4 * l + 3];
  }
}

// 3x3 inverse via Cholesky, as SchurEliminator's InvertPSDMatrix (Ceres 1.9 small_blas / schur_eliminator_impl.h).
static bool invert_psd3(const double* A, double* Ainv) {
  double Lm[9];
  if (llt_lower_eigen(A, Lm, 3) >= 0) {
    for (int i = 0; i < 9; ++i) Ainv[i] = std::numeric_limits<double>::quiet_NaN();
    return false;
  }
  // inverse of L
  double Li[9] = {0};
  for (int c = 0; c < 3; ++c) {
    Li[c * 3 + c] = 1.0 / Lm[c * 3 + c];
    for (int r2 = c + 1; r2 < 3; ++r2) {
      double s = 0;
      for (int k = c; k < r2; ++k) s += Lm[r2 * 3 + k] * Li[k * 3 + c];
      Li[r2 * 3 + c] = -s / Lm[r2 * 3 + r2];
    }
  }
  matmul_tn(Li, Li, Ainv, 3, 3, 3);
  return true;
}

bool Problem::schur_solve(const double* Jv, const double* r, const double* D, double* y) {
  const double t0 = now_s();
  std::vector<double> lhs((size_t)d * d, 0.0), rhs(d, 0.0);
  // F^T F + D_c^2 and F^T r over every residual block (dense parts only)
  for (const RBlock& rb : rblocks) {
    for (int a = 0; a < rb.nb; ++a) {
      const JBlock& ja = jblocks[rb.b0 + a];
      if (ja.col >= d) continue;
      const double* A = Jv + ja.data;
      for (int i = 0; i < ja.w; ++i) {
        double s = 0;
        for (int k = 0; k < rb.m; ++k) s += A[k * ja.w + i] * r[rb.row + k];
        rhs[ja.col + i] += s;
      }
      for (int b = 0; b < rb.nb; ++b) {
        const JBlock& jb = jblocks[rb.b0 + b];
        if (jb.col >= d) continue;
        const double* B = Jv + jb.data;
        for (int i = 0; i < ja.w; ++i)
          for (int j = 0; j < jb.w; ++j) {
            double s = 0;
            for (int k = 0; k < rb.m; ++k) s += A[k * ja.w + i] * B[k * jb.w + j];
            lhs[(size_t)(ja.col + i) * d + jb.col + j] += s;
          }
      }
    }
  }
  for (int i = 0; i < d; ++i) lhs[(size_t)i * d + i] += D[i] * D[i];
  // eliminate landmarks
  std::vector<double> ete_inv((size_t)9 * L), gl((size_t)3 * L);
  bool ok = true;
  struct WB { int col; double w[18]; };  // 3 x 6 block of E^T F
  std::vector<std::vector<WB>> Wall(L);
#pragma omp parallel for schedule(dynamic, 16) num_threads(num_threads)
  for (int l = 0; l < L; ++l) {
    double ete[9] = {0}, g[3] = {0};
    std::vector<WB>& W = Wall[l];
    for (int p = lm_ptr[l]; p < lm_ptr[l + 1]; ++p) {
      const RBlock& rb = rblocks[lm_obs[p]];
      const double* Jl = Jv + jblocks[rb.b0 + 1].data;  // 2x3
      for (int i = 0; i < 3; ++i) {
        g[i] += Jl[i] * r[rb.row] + Jl[3 + i] * r[rb.row + 1];
        for (int j = 0; j < 3; ++j) ete[i * 3 + j] += Jl[i] * Jl[j] + Jl[3 + i] * Jl[3 + j];
      }
      for (int b = 0; b < rb.nb; ++b) {
        if (b == 1) continue;
        const JBlock& jb = jblocks[rb.b0 + b];
        const double* F = Jv + jb.data;  // 2x6
        WB* wb = nullptr;
        for (auto& x : W) if (x.col == jb.col) { wb = &x; break; }
        if (!wb) { W.push_back(WB{jb.col, {0}}); wb = &W.back(); }
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 6; ++j) wb->w[i * 6 + j] += Jl[i] * F[j] + Jl[3 + i] * F[6 + j];
      }
    }
    for (int i = 0; i < 3; ++i) ete[i * 3 + i] += D[lm_off(l) + i] * D[lm_off(l) + i];
    if (!invert_psd3(ete, &ete_inv[9 * l])) {
#pragma omp atomic write
      ok = false;
    }
    for (int i = 0; i < 3; ++i) gl[3 * l + i] = g[i];
  }
  if (ok) {
    for (int l = 0; l < L; ++l) {
      const std::vector<WB>& W = Wall[l];
      const double* Ei = &ete_inv[9 * l];
      double Eig[3];
      matmul(Ei, &gl[3 * l], Eig, 3, 3, 1);
      for (const WB& a : W) {
        double AtE[18];  // 6x3 = W_a^T * Einv
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) AtE[i * 3 + j] = a.w[0 * 6 + i] * Ei[0 * 3 + j] + a.w[1 * 6 + i] * Ei[1 * 3 + j] + a.w[2 * 6 + i] * Ei[2 * 3 + j];
        for (int i = 0; i < 6; ++i) rhs[a.col + i] -= a.w[0 * 6 + i] * Eig[0] + a.w[1 * 6 + i] * Eig[1] + a.w[2 * 6 + i] * Eig[2];
        for (const WB& b : W)
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
              lhs[(size_t)(a.col + i) * d + b.col + j] -= AtE[i * 3 + 0] * b.w[0 * 6 + j] + AtE[i * 3 + 1] * b.w[1 * 6 + j] + AtE[i * 3 + 2] * b.w[2 * 6 + j];
      }
    }
  }
  const double t1 = now_s();
  times.schur += t1 - t0;
  if (!ok) return false;
  // dense Cholesky of the reduced camera system (CHOLMOD in the reference)
  std::vector<double> Lc((size_t)d * d);
  {
    std::vector<double>& Mx = lhs;
    for (int k = 0; k < d; ++k) {
      double x = Mx[(size_t)k * d + k];
      for (int p = 0; p < k; ++p) x -= Lc[(size_t)k * d + p] * Lc[(size_t)k * d + p];
      if (!(x > 0.0)) { times.reduced_solve += now_s() - t1; return false; }
      x = std::sqrt(x);
      Lc[(size_t)k * d + k] = x;
      for (int i = k + 1; i < d; ++i) {
        double s = Mx[(size_t)i * d + k];
        for (int p = 0; p < k; ++p) s -= Lc[(size_t)i * d + p] * Lc[(size_t)k * d + p];
        Lc[(size_t)i * d + k] = s / x;
      }
    }
  }
  std::vector<double> z(d);
  for (int i = 0; i < d; ++i) {
    double s = rhs[i];
    for (int p = 0; p < i; ++p) s -= Lc[(size_t)i * d + p] * z[p];
    z[i] = s / Lc[(size_t)i * d + i];
  }
  for (int i = d - 1; i >= 0; --i) {
    double s = z[i];
    for (int p = i + 1; p < d; ++p) s -= Lc[(size_t)p * d + i] * y[p];
    y[i] = s / Lc[(size_t)i * d + i];
  }
  const double t2 = now_s();
  times.reduced_solve += t2 - t1;
  // back-substitution
  for (int l = 0; l < L; ++l) {
    double v[3] = {gl[3 * l], gl[3 * l + 1], gl[3 * l + 2]};
    for (const WB& a : Wall[l])
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) v[i] -= a.w[i * 6 + j] * y[a.col + j];
    matmul(&ete_inv[9 * l], v, y + lm_off(l), 3, 3, 1);
  }
  times.backsub += now_s() - t2;
  for (int i = 0; i < n_tan; ++i) if (!std::isfinite(y[i])) return false;
  return true;
}

// TrustRegionMinimizer::Minimize + DoglegStrategy (Ceres 1.9) with the options of Estimator.cpp:854-874.
okb_summary Problem::solve(const okb_solve_options& opt, std::vector<IterationRecord>* trace, int nthreads) {
  num_threads = std::max(1, nthreads);
  use_cauchy = opt.use_cauchy_loss != 0;
  times = PhaseTimes();
  const double t_start = now_s();
  okb_summary sum;
  std::memset(&sum, 0, sizeof sum);
  // Ceres defaults
  const double min_relative_decrease = 1e-3, function_tolerance = 1e-6, gradient_tolerance = 1e-10;
  const double parameter_tolerance = 1e-8, min_trust_region_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  const double min_mu = 1e-8, max_mu = 1.0, mu_increase_factor = 10.0;
  const int max_consecutive_invalid = 5;
  double radius = 1e4, mu = min_mu;
  (void)max_radius;

  std::vector<double> r(n_rows), Jv(n_jvals), scale(n_tan), gradient(n_tan);
  std::vector<double> diagonal(n_tan), dl_gradient(n_tan), gn_step(n_tan), step(n_tan), delta(n_tan), lm_diag(n_tan);
  std::vector<double> model_res(n_rows);
  std::vector<double> P2, S2, E2, M2;
  double alpha = 0, dogleg_step_norm = 0;
  bool reuse = false;

  auto x_norm_fn = [&]() {
    double s = 0;
    for (double v : poses) s += v * v;
    for (int e = 0; e < NE; ++e) if (ext_off[e] >= 0) for (int c = 0; c < 7; ++c) s += ext[7 * e + c] * ext[7 * e + c];
    for (double v : sb) s += v * v;
    for (double v : lms) s += v * v;
    return std::sqrt(s);
  };
  auto left_multiply = [&](const double* x, double* yv) {  // y += J^T x
    for (const RBlock& rb : rblocks)
      for (int b = 0; b < rb.nb; ++b) {
        const JBlock& jb = jblocks[rb.b0 + b];
        const double* A = &Jv[jb.data];
        for (int k = 0; k < rb.m; ++k) {
          const double xv = x[rb.row + k];
          for (int c = 0; c < jb.w; ++c) yv[jb.col + c] += A[k * jb.w + c] * xv;
        }
      }
  };
  auto right_multiply = [&](const double* x, double* yv) {  // y += J x
    for (const RBlock& rb : rblocks)
      for (int b = 0; b < rb.nb; ++b) {
        const JBlock& jb = jblocks[rb.b0 + b];
        const double* A = &Jv[jb.data];
        for (int k = 0; k < rb.m; ++k) {
          double s = 0;
          for (int c = 0; c < jb.w; ++c) s += A[k * jb.w + c] * x[jb.col + c];
          yv[rb.row + k] += s;
        }
      }
  };
  auto squared_column_norm = [&](double* out) {
    std::fill(out, out + n_tan, 0.0);
    for (const RBlock& rb : rblocks)
      for (int b = 0; b < rb.nb; ++b) {
        const JBlock& jb = jblocks[rb.b0 + b];
        const double* A = &Jv[jb.data];
        for (int k = 0; k < rb.m; ++k)
          for (int c = 0; c < jb.w; ++c) out[jb.col + c] += A[k * jb.w + c] * A[k * jb.w + c];
      }
  };
  auto scale_columns = [&]() {
    for (const RBlock& rb : rblocks)
      for (int b = 0; b < rb.nb; ++b) {
        const JBlock& jb = jblocks[rb.b0 + b];
        double* A = &Jv[jb.data];
        for (int k = 0; k < rb.m; ++k)
          for (int c = 0; c < jb.w; ++c) A[k * jb.w + c] *= scale[jb.col + c];
      }
  };
  auto gradient_max_norm = [&]() {
    // ||Plus(x, -g) - x||_inf (Ceres 1.9 projected-gradient form)
    std::vector<double> ng(n_tan);
    for (int i = 0; i < n_tan; ++i) ng[i] = -gradient[i];
    plus(ng.data(), P2, S2, E2, M2);
    double m = 0;
    for (size_t i = 0; i < poses.size(); ++i) m = std::max(m, std::fabs(P2[i] - poses[i]));
    for (int e = 0; e < NE; ++e) if (ext_off[e] >= 0) for (int c = 0; c < 7; ++c) m = std::max(m, std::fabs(E2[7 * e + c] - ext[7 * e + c]));
    for (size_t i = 0; i < sb.size(); ++i) m = std::max(m, std::fabs(S2[i] - sb[i]));
    for (size_t i = 0; i < lms.size(); ++i) m = std::max(m, std::fabs(M2[i] - lms[i]));
    return m;
  };

  double tt = now_s();
  double cost = evaluate(poses.data(), sb.data(), ext.data(), lms.data(), r.data(), Jv.data());
  times.evaluate_jac += now_s() - tt;
  std::fill(gradient.begin(), gradient.end(), 0.0);
  left_multiply(r.data(), gradient.data());
  sum.initial_cost = cost;
  double x_norm = x_norm_fn();
  int iteration = 0, num_successful = 0, num_invalid = 0;
  int termination = OKB_TERM_NO_CONVERGENCE;
  double last_iter_time = now_s() - t_start;

  if (gradient_max_norm() <= gradient_tolerance) {
    termination = OKB_TERM_GRADIENT_TOL;
  } else {
    // jacobi_scaling = true: scale = 1/(1+sqrt(colnorm^2)), computed once
    squared_column_norm(scale.data());
    for (int i = 0; i < n_tan; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i]));
    scale_columns();
    while (true) {
      // IterationCallback (CeresIterationCallback.hpp:78-87) on the previous iteration's summary
      if (opt.time_limit_s >= 0 && iteration >= opt.min_iterations &&
          (now_s() - t_start) + last_iter_time > opt.time_limit_s) {
        termination = OKB_TERM_TIME_LIMIT;
        break;
      }
      const double iter_start = now_s();
      if (iteration >= opt.max_iterations) { termination = OKB_TERM_NO_CONVERGENCE; break; }
      ++iteration;
      IterationRecord rec{cost, 0, radius, 0, 0, 0};
      // ---- DoglegStrategy::ComputeStep
      bool solver_ok = true;
      if (!reuse) {
        reuse = true;
        squared_column_norm(diagonal.data());
        for (int i = 0; i < n_tan; ++i) diagonal[i] = std::sqrt(std::min(std::max(diagonal[i], min_diag), max_diag));
        std::fill(dl_gradient.begin(), dl_gradient.end(), 0.0);
        left_multiply(r.data(), dl_gradient.data());
        for (int i = 0; i < n_tan; ++i) dl_gradient[i] /= diagonal[i];
        {  // Cauchy point
          std::vector<double> sg(n_tan), Jg(n_rows, 0.0);
          for (int i = 0; i < n_tan; ++i) sg[i] = dl_gradient[i] / diagonal[i];
          right_multiply(sg.data(), Jg.data());
          double g2 = 0, j2 = 0;
          for (double v : dl_gradient) g2 += v * v;
          for (double v : Jg) j2 += v * v;
          alpha = g2 / j2;
        }
        solver_ok = false;
        while (mu < max_mu) {
          for (int i = 0; i < n_tan; ++i) lm_diag[i] = diagonal[i] * std::sqrt(mu);
          if (schur_solve(Jv.data(), r.data(), lm_diag.data(), gn_step.data())) { solver_ok = true; break; }
          mu *= mu_increase_factor;
        }
        if (solver_ok) for (int i = 0; i < n_tan; ++i) gn_step[i] *= -diagonal[i];
      }
      bool step_valid = false;
      double model_cost_change = 0;
      if (solver_ok) {
        // ComputeTraditionalDoglegStep
        double gradient_norm = 0, gn_norm = 0;
        for (double v : dl_gradient) gradient_norm += v * v;
        for (double v : gn_step) gn_norm += v * v;
        gradient_norm = std::sqrt(gradient_norm); gn_norm = std::sqrt(gn_norm);
        if (gn_norm <= radius) {
          step = gn_step; dogleg_step_norm = gn_norm;
        } else if (gradient_norm * alpha >= radius) {
          for (int i = 0; i < n_tan; ++i) step[i] = -(radius / gradient_norm) * dl_gradient[i];
          dogleg_step_norm = radius;
        } else {
          double gdot = 0;
          for (int i = 0; i < n_tan; ++i) gdot += dl_gradient[i] * gn_step[i];
          const double b_dot_a = -alpha * gdot;
          const double a_squared_norm = std::pow(alpha * gradient_norm, 2.0);
          const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + std::pow(gn_norm, 2);
          const double c = b_dot_a - a_squared_norm;
          const double dd = std::sqrt(c * c + b_minus_a_squared_norm * (std::pow(radius, 2.0) - a_squared_norm));
          const double beta = (c <= 0) ? (dd - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (dd + c);
          double n2 = 0;
          for (int i = 0; i < n_tan; ++i) {
            step[i] = (-alpha * (1.0 - beta)) * dl_gradient[i] + beta * gn_step[i];
            n2 += step[i] * step[i];
          }
          dogleg_step_norm = std::sqrt(n2);
        }
        for (int i = 0; i < n_tan; ++i) step[i] /= diagonal[i];
        // model cost change
        std::fill(model_res.begin(), model_res.end(), 0.0);
        right_multiply(step.data(), model_res.data());
        for (int i = 0; i < n_rows; ++i) model_cost_change -= model_res[i] * (r[i] + model_res[i] / 2.0);
        step_valid = model_cost_change >= 0.0;
      }
      bool step_successful = false;
      if (!step_valid) {
        if (++num_invalid >= max_consecutive_invalid) { termination = OKB_TERM_FAILURE; rec.accepted = -1; if (trace) trace->push_back(rec); break; }
        rec.accepted = -1;
      } else {
        num_invalid = 0;
        for (int i = 0; i < n_tan; ++i) delta[i] = step[i] * scale[i];
        plus(delta.data(), P2, S2, E2, M2);
        tt = now_s();
        const double new_cost = evaluate(P2.data(), S2.data(), E2.data(), M2.data(), nullptr, nullptr);
        times.evaluate_cost += now_s() - tt;
        double sn = 0;
        for (size_t i = 0; i < poses.size(); ++i) sn += (P2[i] - poses[i]) * (P2[i] - poses[i]);
        for (int e = 0; e < NE; ++e) if (ext_off[e] >= 0) for (int c = 0; c < 7; ++c) sn += (E2[7 * e + c] - ext[7 * e + c]) * (E2[7 * e + c] - ext[7 * e + c]);
        for (size_t i = 0; i < sb.size(); ++i) sn += (S2[i] - sb[i]) * (S2[i] - sb[i]);
        for (size_t i = 0; i < lms.size(); ++i) sn += (M2[i] - lms[i]) * (M2[i] - lms[i]);
        rec.step_norm = std::sqrt(sn);
        if (rec.step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
          termination = OKB_TERM_PARAMETER_TOL; if (trace) trace->push_back(rec); break;
        }
        rec.cost_change = cost - new_cost;
        if (std::fabs(rec.cost_change) < function_tolerance * cost) {
          termination = OKB_TERM_FUNCTION_TOL; if (trace) trace->push_back(rec); break;
        }
        rec.relative_decrease = rec.cost_change / model_cost_change;
        step_successful = rec.relative_decrease > min_relative_decrease;
      }
      if (step_successful) {
        ++num_successful;
        // StepAccepted
        if (rec.relative_decrease < 0.25) radius *= 0.5;
        if (rec.relative_decrease > 0.75) radius = std::max(radius, 3.0 * dogleg_step_norm);
        mu = std::max(min_mu, 2.0 * mu / mu_increase_factor);
        reuse = false;
        poses = P2; sb = S2; ext = E2; lms = M2;
        x_norm = x_norm_fn();
        tt = now_s();
        cost = evaluate(poses.data(), sb.data(), ext.data(), lms.data(), r.data(), Jv.data());
        times.evaluate_jac += now_s() - tt;
        std::fill(gradient.begin(), gradient.end(), 0.0);
        left_multiply(r.data(), gradient.data());
        rec.accepted = 1; rec.cost = cost;
        if (gradient_max_norm() <= gradient_tolerance) {
          termination = OKB_TERM_GRADIENT_TOL; rec.radius = radius; if (trace) trace->push_back(rec); break;
        }
        scale_columns();
      } else {
        if (step_valid) { radius *= 0.5; reuse = true; }       // StepRejected
        else { mu *= mu_increase_factor; reuse = false; }      // StepIsInvalid
      }
      rec.radius = radius;
      if (trace) trace->push_back(rec);
      if (radius < min_trust_region_radius) { termination = OKB_TERM_MIN_RADIUS; break; }
      last_iter_time = now_s() - iter_start;
    }
  }
  sum.final_cost = cost;
  sum.iterations = iteration;
  sum.num_successful_steps = num_successful;
  sum.termination = termination;
  sum.final_radius = radius;
  int redo = 0;
  for (auto& c : imu_cache) redo += c.redoCounter;
  sum.imu_redo_count = redo;
  sum.solve_time_s = now_s() - t_start;
  times.other = sum.solve_time_s - (times.evaluate_jac + times.schur + times.reduced_solve + times.backsub + times.evaluate_cost);
  return sum;
}

// symmetric 3x3 eigenvalues (Jacobi sweeps), ascending -- stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>
static void eig3(const double* A, double* ev) {
  double a[9];
  std::memcpy(a, A, sizeof a);
  for (int sweep = 0; sweep < 50; ++sweep) {
    const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p * 3 + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - s * akq; a[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - s * aqk; a[q * 3 + k] = s * apk + c * aqk;
        }
      }
  }
  ev[0] = a[0]; ev[1] = a[4]; ev[2] = a[8];
  std::sort(ev, ev + 3);
}

// Estimator.cpp:880-900 with Map::getLhs (Map.cpp:101-156): H = sum J_lm^T J_lm over the landmark's
// residual blocks (sqrt-information weighted, NOT loss weighted), quality = sqrt(min)/sqrt(max).
void Problem::landmark_quality(std::vector<double>& quality) const {
  quality.assign(L, 0.0);
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int l = 0; l < L; ++l) {
    double H[9] = {0};
    for (int p = lm_ptr[l]; p < lm_ptr[l + 1]; ++p) {
      const okb_observation& ob = obs[lm_obs[p]];
      double res[2], J0[12], J1[6], J2[12];
      reprojection_error(cams[ob.cam_idx], &poses[7 * ob.pose_idx], &lms[4 * ob.lm_idx], &ext[7 * ob.ext_idx], ob.z,
                         ob.sqrt_info, res, J0, J1, J2);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H[i * 3 + j] += J1[i] * J1[j] + J1[3 + i] * J1[3 + j];
    }
    double ev[3];
    eig3(H, ev);
    quality[l] = (ev[0] < 1.0e-12) ? 0.0 : std::sqrt(ev[0]) / std::sqrt(ev[2]);
  }
}

}  // namespace oko
