// Dense Cholesky + triangular solves of the reduced camera system inside one CTA (k_solve).
// Blocked right-looking factorisation, panel width 16:
//   (1) the 16x16 diagonal block is factored by warp 0 with one matrix row per lane (registers +
//       shuffles), (2) the panel below it is solved row-per-thread and transposed into a k-major
//       shared buffer, (3) the trailing matrix gets a rank-16 update as a register-tiled SYRK
//       (4x4 micro-tiles).  ~3 block barriers per 16 columns instead of 3 per column.
// The matrix lives in shared memory when it fits (d <= ~150) and in global memory otherwise.
#pragma once
#include <cuda_runtime.h>

namespace okb {

constexpr int CH_NB = 16;

// In-place lower Cholesky of the d x d row-major matrix M (only the lower triangle is referenced
// and written).  `panel` is shared scratch of CH_NB * ld_p doubles (ld_p >= d rounded up to 4),
// `flag` a shared int.  Returns 0 on success, 1 on a non-positive pivot (uniform over the CTA).
__device__ inline int block_cholesky(double* M, int d, double* panel, int ld_p, double* rdiag, int* flag) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *flag = 0;
  __syncthreads();
  for (int kb = 0; kb < d; kb += CH_NB) {
    const int nb = min(CH_NB, d - kb);
    // ---- (1) diagonal block, warp 0, lane = row
    if (warp == 0) {
      double row[CH_NB];
      const int r = kb + lane;
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) row[c] = (lane < nb && c <= lane && c < nb) ? M[(size_t)r * d + kb + c] : ((c == lane) ? 1.0 : 0.0);
      bool bad = false;
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) {
        const double piv = __shfl_sync(0xffffffffu, row[j], j);
        if (!(piv > 0.0)) bad = true;
        const double rs = rsqrt(piv);            // one slow operation per column instead of sqrt + divide
        const double lij = (lane == j) ? piv * rs : row[j] * rs;
        row[j] = lij;
        if (lane == j && j < nb) rdiag[kb + j] = rs;
#pragma unroll
        for (int c = j + 1; c < CH_NB; ++c) {
          const double lcj = __shfl_sync(0xffffffffu, lij, c);
          if (lane >= c) row[c] -= lij * lcj;
        }
      }
      if (lane < nb) {
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) if (c <= lane && c < nb) M[(size_t)r * d + kb + c] = row[c];
      }
      if (bad && lane == 0) *flag = 1;
    }
    __syncthreads();
    if (*flag) return 1;
    const int r0 = kb + nb;            // first trailing row
    const int n = d - r0;
    if (n <= 0) break;
    // ---- (2) panel solve: X * L_d^T = A  (row per thread), result also stored k-major in `panel`
    for (int i = tid; i < n; i += nthr) {
      double x[CH_NB];
      double* arow = M + (size_t)(r0 + i) * d + kb;
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) x[c] = (c < nb) ? arow[c] : 0.0;
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) {
        if (j < nb) {
          const double* lrow = M + (size_t)(kb + j) * d + kb;
          double s = x[j];
#pragma unroll
          for (int c = 0; c < CH_NB; ++c) if (c < j) s -= x[c] * lrow[c];
          x[j] = s * rdiag[kb + j];
        }
      }
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) {
        if (c < nb) arow[c] = x[c];
        panel[(size_t)c * ld_p + i] = (c < nb) ? x[c] : 0.0;
      }
    }
    // zero the padding columns of the panel so that 4x4 tiles can read past n
    const int n4 = (n + 3) & ~3;
    for (int e = tid; e < CH_NB * (n4 - n); e += nthr) panel[(size_t)(e / (n4 - n)) * ld_p + n + e % (n4 - n)] = 0.0;
    __syncthreads();
    // ---- (3) trailing update: A[r0+i][r0+j] -= sum_c P[c][i] P[c][j], i >= j, 4x4 micro-tiles
    const int nt = n4 >> 2;
    const int ntt = nt * (nt + 1) / 2;
    for (int t = tid; t < ntt; t += nthr) {
      int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while (ti * (ti + 1) / 2 > t) --ti;
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      const int tj = t - ti * (ti + 1) / 2;
      double acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = 0.0;
#pragma unroll 4
      for (int c = 0; c < CH_NB; ++c) {
        const double* pr = panel + (size_t)c * ld_p;
        const double2 a01 = *reinterpret_cast<const double2*>(pr + 4 * ti);
        const double2 a23 = *reinterpret_cast<const double2*>(pr + 4 * ti + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(pr + 4 * tj);
        const double2 b23 = *reinterpret_cast<const double2*>(pr + 4 * tj + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[ii * 4 + jj] += a[ii] * b[jj];
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = 4 * ti + ii, j = 4 * tj + jj;
          if (i < n && j <= i) M[(size_t)(r0 + i) * d + r0 + j] -= acc[ii * 4 + jj];
        }
    }
    __syncthreads();
  }
  return 0;
}

// Solves L L^T x = b in place in `x` (shared memory vector of length d, initialised with b).
__device__ inline void block_cholesky_solve(const double* M, int d, const double* rdiag, double* x) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  // forward: L z = b
  for (int kb = 0; kb < d; kb += CH_NB) {
    const int nb = min(CH_NB, d - kb);
    if (warp == 0) {
      double t = (lane < nb) ? x[kb + lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        const double zj = __shfl_sync(0xffffffffu, t, j) * rdiag[kb + j];
        if (lane == j) t = zj;
        else if (lane > j && lane < nb) t -= M[(size_t)(kb + lane) * d + kb + j] * zj;
      }
      if (lane < nb) x[kb + lane] = t;
    }
    __syncthreads();
    for (int i = kb + nb + tid; i < d; i += nthr) {
      const double* lrow = M + (size_t)i * d + kb;
      double s = x[i];
      for (int c = 0; c < nb; ++c) s -= lrow[c] * x[kb + c];
      x[i] = s;
    }
    __syncthreads();
  }
  // backward: L^T u = z
  const int nblocks = (d + CH_NB - 1) / CH_NB;
  for (int bi = nblocks - 1; bi >= 0; --bi) {
    const int kb = bi * CH_NB;
    const int nb = min(CH_NB, d - kb);
    if (warp == 0) {
      double t = (lane < nb) ? x[kb + lane] : 0.0;
      for (int j = nb - 1; j >= 0; --j) {
        const double uj = __shfl_sync(0xffffffffu, t, j) * rdiag[kb + j];
        if (lane == j) t = uj;
        else if (lane < j) t -= M[(size_t)(kb + j) * d + kb + lane] * uj;
      }
      if (lane < nb) x[kb + lane] = t;
    }
    __syncthreads();
    for (int i = tid; i < kb; i += nthr) {
      double s = x[i];
      for (int c = 0; c < nb; ++c) s -= M[(size_t)(kb + c) * d + i] * x[kb + c];
      x[i] = s;
    }
    __syncthreads();
  }
}

}  // namespace okb
