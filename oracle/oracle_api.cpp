// TEST INFRASTRUCTURE ONLY -- extern "C" surface of the CPU oracle (loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product).
#include <cstring>
#include <vector>

#include "oracle_brisk.hpp"
#include "oracle_errors.hpp"
#include "oracle_gate.hpp"
#include "oracle_marg.hpp"
#include "oracle_marg_apply.hpp"
#include "oracle_matcher.hpp"
#include "oracle_solver.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oko;

extern "C" {

int oko_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void* oko_problem_create(const okb_window_desc* d) { return new Problem(*d); }
void oko_problem_destroy(void* p) { delete static_cast<Problem*>(p); }

// trace: [cap][6] = cost, cost_change, radius, step_norm, relative_decrease, accepted
// phase_times: [7] = evaluate_jac, schur, reduced_solve, backsub, evaluate_cost, quality, other
int oko_solve(void* p, const okb_solve_options* opt, int num_threads, okb_summary* out, double* trace, int cap,
              int* n_trace, double* phase_times) {
  Problem* pb = static_cast<Problem*>(p);
  std::vector<IterationRecord> tr;
  okb_summary s = pb->solve(*opt, &tr, num_threads);
  if (out) *out = s;
  if (trace) {
    const int n = std::min<int>(cap, (int)tr.size());
    for (int i = 0; i < n; ++i) {
      trace[i * 6 + 0] = tr[i].cost; trace[i * 6 + 1] = tr[i].cost_change; trace[i * 6 + 2] = tr[i].radius;
      trace[i * 6 + 3] = tr[i].step_norm; trace[i * 6 + 4] = tr[i].relative_decrease; trace[i * 6 + 5] = tr[i].accepted;
    }
    if (n_trace) *n_trace = n;
  }
  if (phase_times) {
    phase_times[0] = pb->times.evaluate_jac; phase_times[1] = pb->times.schur; phase_times[2] = pb->times.reduced_solve;
    phase_times[3] = pb->times.backsub; phase_times[4] = pb->times.evaluate_cost; phase_times[5] = pb->times.quality;
    phase_times[6] = pb->times.other;
  }
  return 0;
}

// full Estimator::optimize: solve + the per-landmark quality pass (Estimator.cpp:880-900)
void oko_get_state(void* p, double* poses, double* sb, double* lms, double* quality) {
  Problem* pb = static_cast<Problem*>(p);
  if (poses) std::memcpy(poses, pb->poses.data(), pb->poses.size() * 8);
  if (sb) std::memcpy(sb, pb->sb.data(), pb->sb.size() * 8);
  if (lms) std::memcpy(lms, pb->lms.data(), pb->lms.size() * 8);
  if (quality) {
    std::vector<double> q;
    pb->landmark_quality(q);
    std::memcpy(quality, q.data(), q.size() * 8);
  }
}
// Test plumbing for chained marginalisation checks: overwrite the estimates (caches untouched) and bring an ImuError
// cache into the state "preintegrated at sb_ref" (what the device reports through okb_debug_imu_cache).
void oko_set_state(void* p, const double* poses, const double* sb, const double* lms) {
  Problem* pb = static_cast<Problem*>(p);
  if (poses) std::memcpy(pb->poses.data(), poses, pb->poses.size() * 8);
  if (sb) std::memcpy(pb->sb.data(), sb, pb->sb.size() * 8);
  if (lms) std::memcpy(pb->lms.data(), lms, pb->lms.size() * 8);
}
int oko_set_imu_cache(void* p, int term, const double* sb_ref, int valid) {
  Problem* pb = static_cast<Problem*>(p);
  if (term < 0 || term >= (int)pb->imu_terms.size()) return -1;
  ImuCache& c = pb->imu_cache[term];
  if (!valid) { c.redo = true; return 0; }
  const okb_imu_term& T = pb->imu_terms[term];
  imu_redo_preintegration(pb->samples.data() + T.sample_offset, (int)T.sample_count, pb->imu_params, T.t0_ns, T.t1_ns, sb_ref, c);
  c.redo = false;
  return 0;
}
double oko_cost(void* p) { return static_cast<Problem*>(p)->cost_only(); }

void oko_eval_reprojection(int n, const okb_camera* cam, const double* pose, const double* lm, const double* ext,
                           const double* z, const double* sqrt_info, double* r, double* J0, double* J1, double* J2) {
  for (int i = 0; i < n; ++i)
    reprojection_error(*cam, pose + 7 * i, lm + 4 * i, ext + 7 * i, z + 2 * i, sqrt_info[i], r + 2 * i,
                       J0 ? J0 + 12 * i : nullptr, J1 ? J1 + 6 * i : nullptr, J2 ? J2 + 12 * i : nullptr);
}

int oko_project(const okb_camera* cam, const double* pt3, double* ip, double* J23) {
  return project(*cam, pt3, ip, J23) ? 1 : 0;
}

// sb_ref == NULL: fresh functor (redo_ = true).  Otherwise the cache is first preintegrated at sb_ref.
int oko_eval_imu(const okb_imu_params* prm, const okb_imu_sample* s, int n, int64_t t0, int64_t t1,
                 const double* pose0, const double* sb0, const double* pose1, const double* sb1, const double* sb_ref,
                 double* r, double* J0, double* J1, double* J2, double* J3, double* sqrt_info_out) {
  ImuCache c;
  if (sb_ref) {
    imu_redo_preintegration(s, n, *prm, t0, t1, sb_ref, c);
    c.redo = false;
  }
  imu_error(s, n, *prm, t0, t1, pose0, sb0, pose1, sb1, c, r, J0, J1, J2, J3);
  if (sqrt_info_out) std::memcpy(sqrt_info_out, c.squareRootInformation, sizeof c.squareRootInformation);
  return c.redoCounter;
}

int oko_imu_propagate(const okb_imu_params* prm, const okb_imu_sample* s, int n, int64_t t0, int64_t t1, double* pose,
                      double* sb, double* covariance, double* jacobian) {
  return imu_propagation(s, n, *prm, pose, sb, t0, t1, covariance, jacobian);
}

void oko_eval_pose_error(const double* meas, const double* sqrt_info, const double* pose, double* r, double* J) {
  pose_error(meas, sqrt_info, pose, r, J);
}
void oko_eval_speed_bias_error(const double* meas, const double* sqrt_info, const double* sb, double* r, double* J) {
  speed_bias_error(meas, sqrt_info, sb, r, J);
}
void oko_eval_relative_pose(const double* sqrt_info, const double* p0, const double* p1, double* r, double* J0,
                            double* J1) {
  relative_pose_error(sqrt_info, p0, p1, r, J0, J1);
}
void oko_eval_marginalization(const okb_marg_prior* m, const double* x, double* r, double* J_eff) {
  marginalization_error(*m, nullptr, x, r, J_eff);
}

void oko_pose_plus(const double* x, const double* delta, double* out) { pose_plus(x, delta, out); }
void oko_pose_minus(const double* x, const double* xpd, double* delta) { pose_minus(x, xpd, delta); }
void oko_pose_lift_jacobian(const double* x, double* J67) { pose_lift_jacobian(x, J67); }
void oko_pose_plus_jacobian(const double* x, double* J76) { pose_plus_jacobian(x, J76); }
int oko_sqrt_information(const double* info, int n, double* out) { return sqrt_information(info, out, n); }

// DenseMatcher over an explicit float distance matrix [nA][nB] (for the reference's known-answer tests)
int oko_match_matrix(const float* D, int nA, int nB, const uint8_t* skipA, const uint8_t* skipB, float threshold,
                     int num_best, int use_ratio, float ratio_threshold, okb_pair* topk, okb_pair* pairs,
                     int32_t* matches /*[nB][2]*/, float* match_dist) {
  std::vector<Match> m;
  dense_match(nA, nB, [&](int a, int b) { return D[(size_t)a * nB + b]; }, skipA, skipB, threshold, num_best,
              use_ratio != 0, ratio_threshold, topk, pairs, &m);
  for (size_t i = 0; i < m.size(); ++i) {
    if (matches) { matches[2 * i] = m[i].a; matches[2 * i + 1] = m[i].b; }
    if (match_dist) match_dist[i] = m[i].d;
  }
  return (int)m.size();
}

int oko_match_hamming(const uint8_t* A, int nA, const uint8_t* B, int nB, int desc_bytes, const uint8_t* skipA,
                      const uint8_t* skipB, float threshold, int num_best, int use_ratio, float ratio_threshold,
                      okb_pair* topk, okb_pair* pairs, int32_t* matches, float* match_dist) {
  std::vector<Match> m;
  dense_match(nA, nB,
              [&](int a, int b) { return (float)hamming(A + (size_t)a * desc_bytes, B + (size_t)b * desc_bytes, desc_bytes); },
              skipA, skipB, threshold, num_best, use_ratio != 0, ratio_threshold, topk, pairs, &m);
  for (size_t i = 0; i < m.size(); ++i) {
    if (matches) { matches[2 * i] = m[i].a; matches[2 * i + 1] = m[i].b; }
    if (match_dist) match_dist[i] = m[i].d;
  }
  return (int)m.size();
}

// DenseMatcher over VioKeyframeWindowMatchingAlgorithm::distance: Hamming below the threshold AND verifyMatch, else FLT_MAX
int oko_match_hamming_gated(const uint8_t* A, int nA, const uint8_t* B, int nB, int desc_bytes, const uint8_t* skipA, const uint8_t* skipB,
                            float threshold, int num_best, int use_ratio, float ratio_threshold, const okb_match_gate* gate, okb_pair* topk,
                            okb_pair* pairs) {
  std::vector<Match> m;
  dense_match(nA, nB,
              [&](int a, int b) {
                const float d = (float)hamming(A + (size_t)a * desc_bytes, B + (size_t)b * desc_bytes, desc_bytes);
                if (d < threshold && verify_match(*gate, a, b)) return d;
                return std::numeric_limits<float>::max();
              },
              skipA, skipB, threshold, num_best, use_ratio != 0, ratio_threshold, topk, pairs, &m);
  return (int)m.size();
}

// candidate lists: every B with distance < threshold per A, ascending B (CSR); returns total count
int oko_hamming_candidates(const uint8_t* A, int nA, const uint8_t* B, int nB, int desc_bytes, float threshold,
                           uint32_t* row_ptr, uint32_t* col_idx, uint16_t* dist, int cap) {
  int n = 0;
  for (int a = 0; a < nA; ++a) {
    row_ptr[a] = n;
    for (int b = 0; b < nB; ++b) {
      const uint32_t d = hamming(A + (size_t)a * desc_bytes, B + (size_t)b * desc_bytes, desc_bytes);
      if ((float)d < threshold) {
        if (n < cap) { col_idx[n] = b; dist[n] = (uint16_t)d; }
        ++n;
      }
    }
  }
  row_ptr[nA] = n;
  return n;
}

int oko_detect_describe(const uint8_t* img, int W, int H, int stride, const okb_camera* cam, const double* R_CW,
                        const okb_detect_params* prm, okb_keypoint* kps, uint8_t* desc, int max_out) {
  return detect_describe(img, W, H, stride, *cam, R_CW, *prm, kps, desc, max_out);
}

// ---- marginalisation numeric core (NEXT TIER groundwork, oracle_marg.hpp)
// ranges: [n_ranges][2] = (start, length), sorted and non-overlapping.  H (n x n) / b are overwritten with the reduced
// system in their leading n_out * n_out / n_out entries (row stride n_out).  Returns n_out.
int oko_marginalize_stage(double* H, double* b, int n, const int* ranges, int n_ranges, int landmark_blocks) {
  std::vector<double> Hv(H, H + (size_t)n * n), bv(b, b + n);
  std::vector<std::pair<int, int>> r;
  for (int i = 0; i < n_ranges; ++i) r.emplace_back(ranges[2 * i], ranges[2 * i + 1]);
  int nn = n;
  marginalize_stage(Hv, bv, nn, r, landmark_blocks != 0);
  std::memcpy(H, Hv.data(), sizeof(double) * (size_t)nn * nn);
  std::memcpy(b, bv.data(), sizeof(double) * nn);
  return nn;
}
int oko_marg_update_error_computation(const double* H, const double* b, int n, double* J, double* e0) {
  std::vector<double> Hv(H, H + (size_t)n * n), bv(b, b + n), Jv, ev;
  const int rank = marg_update_error_computation(Hv, bv, n, Jv, ev);
  std::memcpy(J, Jv.data(), sizeof(double) * (size_t)n * n);
  std::memcpy(e0, ev.data(), sizeof(double) * n);
  return rank;
}
// One marginalisation step on the problem's graph (oracle_marg_apply.hpp).  Outputs sized for n <= 160 / 64 blocks.
// Returns the dimension of the new prior (-1: a residual's parameter block is missing from the job).
int oko_marginalize(void* p, const okb_marg_job* job, const double* H_prev, const double* b0_prev, int32_t* kind, uint32_t* idx, double* x0,
                    double* J, double* e0, double* H, double* b0, int* rank) {
  Problem* pb = static_cast<Problem*>(p);
  MargOut o;
  const int n = marginalize_problem(*pb, *job, H_prev, b0_prev, o);
  if (n < 0) return n;
  for (size_t i = 0; i < o.kind.size(); ++i) { kind[i] = o.kind[i]; idx[i] = o.idx[i]; }
  std::memcpy(x0, o.x0.data(), sizeof(double) * o.x0.size());
  std::memcpy(J, o.J.data(), sizeof(double) * (size_t)n * n);
  std::memcpy(e0, o.e0.data(), sizeof(double) * n);
  std::memcpy(H, o.H.data(), sizeof(double) * (size_t)n * n);
  std::memcpy(b0, o.b0.data(), sizeof(double) * n);
  if (rank) *rank = o.rank;
  return n;
}
int oko_sym_eig(const double* A, int n, double* evals, double* evecs) {
  std::vector<double> Av(A, A + (size_t)n * n), ev, V;
  sym_eig(Av, n, ev, V);
  std::memcpy(evals, ev.data(), sizeof(double) * n);
  std::memcpy(evecs, V.data(), sizeof(double) * (size_t)n * n);
  return 0;
}

}  // extern "C"
