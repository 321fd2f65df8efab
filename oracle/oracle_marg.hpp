// TEST INFRASTRUCTURE ONLY -- CPU restatement of the dense numeric core of OKVIS' marginalisation
// (NEXT TIER: no device path exists for it yet; this file is groundwork so that a
// future device implementation has an oracle to be checked against).
//
// Restated from
//   MarginalizationError::marginalizeOut        okvis_ceres/src/MarginalizationError.cpp:618-741
//       (the bookkeeping before and after -- which blocks, ordering indices, Map removal -- is host logic of the
//        reference and not restated here)
//   MarginalizationError::updateErrorComputation  okvis_ceres/src/MarginalizationError.cpp:806-846
//   pseudoInverseSymmSqrt                        okvis_ceres/include/okvis/ceres/implementation/MarginalizationError.hpp:215-243
// PARITY UNPINNED: the reference relies on Eigen::SelfAdjointEigenSolver; eigenvector signs / the basis inside
// degenerate eigenspaces are not unique, so J and e0 are only defined up to an orthogonal transform of the
// residual space.  Invariants that ARE unique (and what the tests check): H' , b', J^T J, J^T e0, the rank.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <utility>
#include <vector>

namespace oko {

// Symmetric eigen-decomposition by cyclic Jacobi rotations.  A (n x n, row-major, symmetric) is overwritten;
// on return evals[i] ascending and evecs column i (row-major n x n: evecs[r*n + i]) the matching unit vector.
inline void sym_eig(std::vector<double>& A, int n, std::vector<double>& evals, std::vector<double>& evecs) {
  evecs.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) evecs[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-32 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {   // columns p, q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {   // rows p, q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = evecs[(size_t)k * n + p], vkq = evecs[(size_t)k * n + q];
          evecs[(size_t)k * n + p] = c * vkp - s * vkq;
          evecs[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
  evals.resize(n);
  std::vector<double> V((size_t)n * n);
  for (int i = 0; i < n; ++i) {
    evals[i] = A[(size_t)order[i] * n + order[i]];
    for (int r = 0; r < n; ++r) V[(size_t)r * n + i] = evecs[(size_t)r * n + order[i]];
  }
  evecs.swap(V);
}

// result = V * diag(sqrt(pinv(lambda)))  (MarginalizationError.hpp(impl):215-243); eigenvalues <= epsilon * n * lambda_max
// count as zero.  Returns the rank.
inline int pseudo_inverse_symm_sqrt(const std::vector<double>& a, int n, std::vector<double>& result,
                                    double epsilon = std::numeric_limits<double>::epsilon()) {
  std::vector<double> A = a, ev, V;
  sym_eig(A, n, ev, V);
  double lmax = ev.empty() ? 0.0 : ev[0];
  for (double e : ev) lmax = std::max(lmax, e);
  const double tol = epsilon * n * lmax;
  result.assign((size_t)n * n, 0.0);
  int rank = 0;
  for (int i = 0; i < n; ++i) {
    const double w = (ev[i] > tol) ? std::sqrt(1.0 / ev[i]) : 0.0;
    if (ev[i] > tol) ++rank;
    for (int r = 0; r < n; ++r) result[(size_t)r * n + i] = V[(size_t)r * n + i] * w;
  }
  return rank;
}

// preconditioner p = sqrt(diag(H)) where diag > 1e-9, else 1e-3 (MarginalizationError.cpp:618, 695, 816)
inline std::vector<double> marg_preconditioner(const std::vector<double>& H, int n) {
  std::vector<double> p(n);
  for (int i = 0; i < n; ++i) p[i] = (H[(size_t)i * n + i] > 1.0e-9) ? std::sqrt(H[(size_t)i * n + i]) : 1.0e-3;
  return p;
}

// Index lists of the kept (a) and marginalised (b) coordinates for sorted, non-overlapping (start, length) ranges
// (splitVector / splitSymmetricMatrix of the reference keep the original order inside both parts).
inline void split_indices(const std::vector<std::pair<int, int>>& ranges, int n, std::vector<int>& ia, std::vector<int>& ib) {
  std::vector<char> is_b(n, 0);
  for (const auto& r : ranges)
    for (int k = 0; k < r.second; ++k) is_b[r.first + k] = 1;
  ia.clear(); ib.clear();
  for (int i = 0; i < n; ++i) (is_b[i] ? ib : ia).push_back(i);
}

// One marginalisation stage on (H, b): Schur complement of the coordinates in `ranges`, exactly in the reference's
// order of operations: precondition, split, pseudo-inverse (per 3x3 block when `landmark_blocks`, of the whole
// 0.5 (V + V^T) otherwise), Schur, un-precondition.  H (n x n) and b are replaced by the reduced system.
inline void marginalize_stage(std::vector<double>& H, std::vector<double>& b, int& n, const std::vector<std::pair<int, int>>& ranges,
                              bool landmark_blocks) {
  if (ranges.empty()) return;
  const std::vector<double> p = marg_preconditioner(H, n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) H[(size_t)i * n + j] = H[(size_t)i * n + j] / p[i] / p[j];
    b[i] /= p[i];
  }
  std::vector<int> ia, ib;
  split_indices(ranges, n, ia, ib);
  const int na = (int)ia.size(), nb = (int)ib.size();
  std::vector<double> U((size_t)na * na), W((size_t)na * nb), V((size_t)nb * nb), ba(na), bb(nb), pa(na);
  for (int i = 0; i < na; ++i) {
    ba[i] = b[ia[i]]; pa[i] = p[ia[i]];
    for (int j = 0; j < na; ++j) U[(size_t)i * na + j] = H[(size_t)ia[i] * n + ia[j]];
    for (int j = 0; j < nb; ++j) W[(size_t)i * nb + j] = H[(size_t)ia[i] * n + ib[j]];
  }
  for (int i = 0; i < nb; ++i) {
    bb[i] = b[ib[i]];
    for (int j = 0; j < nb; ++j) V[(size_t)i * nb + j] = H[(size_t)ib[i] * n + ib[j]];
  }
  std::vector<double> Hn = U, bn = ba;
  auto apply = [&](int c0, int m, const std::vector<double>& Vis) {   // columns c0..c0+m of W with V^-1/2 (m x m)
    std::vector<double> M((size_t)na * m, 0.0);
    for (int i = 0; i < na; ++i)
      for (int j = 0; j < m; ++j) {
        double s = 0;
        for (int k = 0; k < m; ++k) s += W[(size_t)i * nb + c0 + k] * Vis[(size_t)k * m + j];
        M[(size_t)i * m + j] = s;
      }
    std::vector<double> t(m, 0.0);      // V^-1/2^T b_b
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < m; ++k) t[j] += Vis[(size_t)k * m + j] * bb[c0 + k];
    for (int i = 0; i < na; ++i) {
      double s = 0;
      for (int j = 0; j < m; ++j) s += M[(size_t)i * m + j] * t[j];
      bn[i] -= s;
      for (int i2 = 0; i2 < na; ++i2) {
        double h = 0;
        for (int j = 0; j < m; ++j) h += M[(size_t)i * m + j] * M[(size_t)i2 * m + j];
        Hn[(size_t)i * na + i2] -= h;
      }
    }
  };
  if (landmark_blocks) {
    for (int c0 = 0; c0 + 3 <= nb; c0 += 3) {
      std::vector<double> V1(9), Vis;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V1[i * 3 + j] = V[(size_t)(c0 + i) * nb + c0 + j];
      pseudo_inverse_symm_sqrt(V1, 3, Vis);
      apply(c0, 3, Vis);
    }
  } else {
    std::vector<double> V1((size_t)nb * nb), Vis;
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j < nb; ++j) V1[(size_t)i * nb + j] = 0.5 * (V[(size_t)i * nb + j] + V[(size_t)j * nb + i]);
    pseudo_inverse_symm_sqrt(V1, nb, Vis);
    apply(0, nb, Vis);
  }
  for (int i = 0; i < na; ++i) {
    bn[i] *= pa[i];
    for (int j = 0; j < na; ++j) Hn[(size_t)i * na + j] *= pa[i] * pa[j];
  }
  H.swap(Hn); b.swap(bn); n = na;
}

// updateErrorComputation: H = J^T J, e0 = -pinv(J^T) b.  J is (n x n) row-major (rows = residuals), returns the rank.
inline int marg_update_error_computation(const std::vector<double>& H, const std::vector<double>& b, int n, std::vector<double>& J,
                                         std::vector<double>& e0) {
  const std::vector<double> p = marg_preconditioner(H, n);
  std::vector<double> A((size_t)n * n), ev, V;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = 0.5 * (H[(size_t)i * n + j] + H[(size_t)j * n + i]) / p[i] / p[j];
  sym_eig(A, n, ev, V);
  double lmax = ev[0];
  for (double e : ev) lmax = std::max(lmax, e);
  const double tol = std::numeric_limits<double>::epsilon() * n * lmax;
  J.assign((size_t)n * n, 0.0);
  e0.assign(n, 0.0);
  int rank = 0;
  for (int k = 0; k < n; ++k) {            // residual row k <- eigenpair k
    const bool ok = ev[k] > tol;
    if (ok) ++rank;
    const double s = ok ? std::sqrt(ev[k]) : 0.0, si = ok ? std::sqrt(1.0 / ev[k]) : 0.0;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      J[(size_t)k * n + i] = s * V[(size_t)i * n + k] * p[i];
      acc += V[(size_t)i * n + k] / p[i] * b[i];
    }
    e0[k] = -si * acc;
  }
  return rank;
}

}  // namespace oko
