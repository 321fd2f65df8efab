// Dense Cholesky + triangular solves of the reduced camera system inside one CTA (k_solve).
// Blocked right-looking factorisation, panel width CH_NB:
//   (1) the 16x16 diagonal block is factored by warp 0 with one matrix row per lane (registers +
//       shuffles), (2) the panel below it is solved row-per-thread and transposed into a k-major
//       shared buffer, (3) the trailing matrix gets the rank-8 update on the FP64 tensor cores
//       (mma.sync.m8n8k4.f64, 8x8 tiles, one warp per tile).  ~3 block barriers per 16 columns instead of 3 per column.
// The matrix lives in shared memory when it fits (d <= ~150) and in global memory otherwise.
#pragma once
#include <cuda_runtime.h>

namespace okb {

constexpr int CH_NB = 8;

// Packed lower-triangular storage: row r holds columns 0..r at offset r(r+1)/2.
__host__ __device__ inline size_t tri_row(int r) { return (size_t)r * (r + 1) / 2; }

// In-place lower Cholesky of the leading d x d block of the matrix M held in packed lower-triangular storage
// (tri_row).  M has nrows >= d rows: rows d..nrows-1 are carried
// along like any sub-diagonal row, so an appended right-hand side row g^T comes out as (L^-1 g)^T -- the
// forward substitution rides on the factorisation.  `panel` is shared scratch of CH_NB * ld_p doubles
// (ld_p >= nrows rounded up to 4), `flag` a shared int.  Returns 0 on success, 1 on a non-positive
// pivot (uniform over the CTA).
__device__ inline int block_cholesky(double* M, int d, int nrows, double* panel, int ld_p, double* rdiag, int* flag,
                                     unsigned long long* prof = nullptr, bool use_dmma = true) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *flag = 0;
  __syncthreads();
#ifdef OKB_CHOL_PROF   // cycles of thread 0 per sub-phase: diagonal block | panel solve | trailing update
  unsigned long long t_ph = clock64();
#define CHOL_MARK(i) do { if (tid == 0 && prof) { const unsigned long long n_ = clock64(); prof[i] += n_ - t_ph; t_ph = n_; } } while (0)
#else
#define CHOL_MARK(i) do { } while (0)
#endif
  for (int kb = 0; kb < d; kb += CH_NB) {
    const int nb = min(CH_NB, d - kb);
    // ---- (1) diagonal block, warp 0, lane = row
    if (warp == 0) {
      double row[CH_NB];
      const int r = kb + lane;
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) row[c] = (lane < nb && c <= lane && c < nb) ? M[tri_row(r) + kb + c] : ((c == lane) ? 1.0 : 0.0);
      // The pivot of column j+1 is final after the first update of step j: its shuffle and reciprocal square
      // root are issued right there, so that long dependent chain overlaps the remaining updates of step j.
      double piv = __shfl_sync(0xffffffffu, row[0], 0);
      bool bad = !(piv > 0.0);
      double rs = rsqrt(piv);                    // one slow operation per column instead of sqrt + divide
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) {
        const double lij = (lane == j) ? piv * rs : row[j] * rs;
        row[j] = lij;
        if (lane == j && j < nb) rdiag[kb + j] = rs;
        if (j + 1 < CH_NB) {
          const double l1 = __shfl_sync(0xffffffffu, lij, j + 1);
          if (lane >= j + 1) row[j + 1] -= lij * l1;
          piv = __shfl_sync(0xffffffffu, row[j + 1], j + 1);
          if (!(piv > 0.0)) bad = true;
          rs = rsqrt(piv);
        }
#pragma unroll
        for (int c = j + 2; c < CH_NB; ++c) {
          const double lcj = __shfl_sync(0xffffffffu, lij, c);
          if (lane >= c) row[c] -= lij * lcj;
        }
      }
      if (lane < nb) {
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) if (c <= lane && c < nb) M[tri_row(r) + kb + c] = row[c];
      }
      if (bad && lane == 0) *flag = 1;
    }
    __syncthreads();
    CHOL_MARK(0);
    if (*flag) return 1;
    const int r0 = kb + nb;            // first trailing row
    const int n = nrows - r0;          // trailing rows (including the appended ones)
    const int ncol = d - r0;           // trailing columns that still get factored
    if (n <= 0) break;
    // ---- (2) panel solve: X * L_d^T = A  (row per thread), result also stored k-major in `panel`
    for (int i = tid; i < n; i += nthr) {
      double x[CH_NB];
      double* arow = M + tri_row(r0 + i) + kb;
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) x[c] = (c < nb) ? arow[c] : 0.0;
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) {
        if (j < nb) {
          const double* lrow = M + tri_row(kb + j) + kb;
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
    CHOL_MARK(1);
    // ---- (3) trailing update: A[r0+i][r0+j] -= sum_c P[c][i] P[c][j], i >= j.  This is the one dense contraction of
    // the reduced solve, so it runs on the FP64 tensor cores: 8x8 output tiles, one warp per tile, the rank-8 panel
    // product as two mma.sync.m8n8k4.f64 (A = P^T tile rows, B = P tile columns, both read from the k-major panel in
    // shared memory).  Fragment layout (PTX ISA, m8n8k4 f64): lane = 4 g + t holds A[g][t], B[t][g], C[g][2t..2t+1].
    if (use_dmma) {
      const int nt8 = (n + 7) >> 3;
      const int ntt8 = nt8 * (nt8 + 1) / 2;
      const int nwarps = nthr >> 5;
      const int g = lane >> 2, t4 = lane & 3;
      const int n4m1 = n4 - 1;
      for (int t = warp; t < ntt8; t += nwarps) {
        int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while (ti * (ti + 1) / 2 > t) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        const int tj = t - ti * (ti + 1) / 2;
        const int ia = min(8 * ti + g, n4m1), jb = min(8 * tj + g, n4m1);     // rows beyond n are zero padding (masked below)
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < CH_NB; k0 += 4) {
          const double a = panel[(size_t)(k0 + t4) * ld_p + ia];
          const double b = panel[(size_t)(k0 + t4) * ld_p + jb];
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
        }
        const int i = 8 * ti + g, j = 8 * tj + 2 * t4;
        if (i < n) {
          double* mrow = M + tri_row(r0 + i) + r0;
          if (j <= i && j < ncol) mrow[j] -= d0;
          if (j + 1 <= i && j + 1 < ncol) mrow[j + 1] -= d1;
        }
      }
    } else {
      // matrix in global memory (d > ~150): register-tiled 4x4 micro-tiles, one per thread -- sixteen independent
      // read-modify-writes in flight per thread hide the L2 latency better than the warp-wide tensor-core tiles
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
            if (i < n && j <= i && j < ncol) M[tri_row(r0 + i) + r0 + j] -= acc[ii * 4 + jj];
          }
      }
    }
    __syncthreads();
    CHOL_MARK(2);
  }
#undef CHOL_MARK
  return 0;
}

// Solves L^T u = z in place in `x` (shared memory vector of length d holding z = L^-1 b, which the
// factorisation produced in the appended row).
__device__ inline void block_cholesky_backward(const double* M, int d, const double* rdiag, double* x) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nblocks = (d + CH_NB - 1) / CH_NB;
  for (int bi = nblocks - 1; bi >= 0; --bi) {
    const int kb = bi * CH_NB;
    const int nb = min(CH_NB, d - kb);
    if (warp == 0) {
      double t = (lane < nb) ? x[kb + lane] : 0.0;
      // the block's entries and reciprocal diagonal are fetched before the serial chain starts
      double m[CH_NB], rd[CH_NB];
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) {
        m[j] = (j < nb && lane < j) ? M[tri_row(kb + j) + kb + lane] : 0.0;
        rd[j] = (j < nb) ? rdiag[kb + j] : 0.0;
      }
#pragma unroll
      for (int j = CH_NB - 1; j >= 0; --j) {
        if (j < nb) {
          const double uj = __shfl_sync(0xffffffffu, t, j) * rd[j];
          if (lane == j) t = uj;
          else if (lane < j) t -= m[j] * uj;
        }
      }
      if (lane < nb) x[kb + lane] = t;
    }
    __syncthreads();
    for (int i = tid; i < kb; i += nthr) {
      double s = x[i];
      for (int c = 0; c < nb; ++c) s -= M[tri_row(kb + c) + i] * x[kb + c];
      x[i] = s;
    }
    __syncthreads();
  }
}

}  // namespace okb
