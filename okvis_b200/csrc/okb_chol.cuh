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

// ------------------------------------------------------------------------------------------------
// Speed/bias chain elimination.  The reduced system orders the unknowns [poses (dc) | speed/bias blocks (9 each)].
// The speed/bias part A of the matrix is block-tridiagonal (an IMU term couples consecutive frames, priors touch one
// block), so it is eliminated first:  A = L_A L_A^T block by block (one warp, an 18-row window in registers),
// Y = L_A^-1 B column by column (one thread per pose column -- column j of block row b only needs column j of block
// row b-1, so the recursion runs without a barrier), C' = C - Y^T Y on the FP64 tensor cores, and only the dc x dc
// pose system C' goes through the dense blocked factorisation.
// ChainView addresses A through one offset per speed/bias row i (block b = i / 9):  ab[ao[i] + k] is the entry in
// column 9 (b - 1) + k, k = 0..8 the previous block's columns and k = 9 + c this block's column c.  That covers both
// homes of A: in place in the packed system (ao[i] = tri_row(dc + i) + dc + 9 b - 9) and the banded shared-memory
// copy used when the system itself lives in global memory (ao[i] = 18 i).  yo[i] = tri_row(dc + i) is row i of Y.
// ------------------------------------------------------------------------------------------------
struct ChainView {
  double* ab;
  const int* ao;
  double* M;
  const int* yo;
  int dc;
};

// warp 0 only.  rd[i] receives 1 / L_ii.  Sets *flag on a non-positive pivot.  pipelined: arrives at the named barrier 1 + b after block b
// (producer side of the hand-off to chain_forward; needs nsb <= 15).
// Lanes 0..8 hold the rows of the diagonal block A_bb, lanes 9..17 the rows of [A_{b+1,b} | A_{b+1,b+1}]: nine
// right-looking column steps leave L_bb, L_{b+1,b} and the updated A_{b+1,b+1} in the registers.
__device__ inline void chain_factor(const ChainView& A, int nsb, double* rd, int* flag, bool pipelined) {
  const int lane = threadIdx.x & 31;
  bool bad = false;
  for (int b = 0; b < nsb; ++b) {
    const int i0 = 9 * b;
    const bool mine = lane < 9 || (lane < 18 && b + 1 < nsb);
    double* q = A.ab + (mine ? A.ao[i0 + lane] + (lane < 9 ? 9 : 0) : 0);
    double row[18];
#pragma unroll
    for (int c = 0; c < 18; ++c) row[c] = (mine && c <= lane) ? q[c] : 0.0;
    double piv = __shfl_sync(0xffffffffu, row[0], 0);
    if (!(piv > 0.0)) bad = true;
    double rs = rsqrt(piv);
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const double lij = (lane == j) ? piv * rs : row[j] * rs;
      row[j] = lij;
      if (lane == j) rd[i0 + j] = rs;
      if (j + 1 < 9) {     // the next pivot first: its reciprocal square root overlaps the remaining updates
        const double l1 = __shfl_sync(0xffffffffu, lij, j + 1);
        if (lane >= j + 1) row[j + 1] -= lij * l1;
        piv = __shfl_sync(0xffffffffu, row[j + 1], j + 1);
        if (!(piv > 0.0)) bad = true;
        rs = rsqrt(piv);
      }
#pragma unroll
      for (int c = (j + 1 < 9) ? j + 2 : j + 1; c < 18; ++c) {
        const double lcj = __shfl_sync(0xffffffffu, lij, c);
        if (lane >= c) row[c] -= lij * lcj;
      }
    }
    if (mine) {
#pragma unroll
      for (int c = 0; c < 18; ++c) if (c <= lane) q[c] = row[c];
    }
    __syncwarp();
    if (pipelined) asm volatile("bar.arrive %0, %1;" :: "r"(1 + b), "r"((int)blockDim.x) : "memory");     // releases chain_forward's block row b
  }
  if (bad && lane == 0) *flag = 1;
}

// Y = L_A^-1 [B | g_s] in place: one thread per pose column j of the speed/bias rows of M (j < dc) or the right-hand
// side (j == dc, the vector rhs_s of length 9 nsb).  No block barrier inside.
// pipelined: called by the warps >= 1 WHILE warp 0 runs chain_factor; block row b starts when the named barrier
// 1 + b completes (L_bb and L_{b,b-1} are final), so the recursion trails the factorisation by one block instead of
// following it.  Needs dc + 1 <= blockDim.x - 32 (one column per thread).  Otherwise all threads call it after a
// __syncthreads.
__device__ inline void chain_forward(const ChainView& A, int nsb, const double* rd, double* rhs_s, bool pipelined) {
  const int dc = A.dc, nthr = (int)blockDim.x, skip = pipelined ? 32 : 0;
  for (int j0 = 0; j0 <= dc; j0 += nthr - skip) {
    const int j = j0 + (int)threadIdx.x - skip;
    const bool act = j <= dc;
    double yp[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) yp[k] = 0.0;
    for (int b = 0; b < nsb; ++b) {
      const int i0 = 9 * b;
      if (pipelined) {
        __syncwarp();      // lanes without a column take part in the barrier, converged
        asm volatile("bar.sync %0, %1;" :: "r"(1 + b), "r"(nthr) : "memory");
      }
      if (!act) continue;
      double t[9];
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const double* lr = A.ab + A.ao[i0 + r];
        double s = (j < dc) ? A.M[A.yo[i0 + r] + j] : rhs_s[i0 + r];
        if (b > 0) {
#pragma unroll
          for (int k = 0; k < 9; ++k) s -= lr[k] * yp[k];
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) if (c < r) s -= lr[9 + c] * t[c];
        t[r] = s * rd[i0 + r];
      }
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        if (j < dc) A.M[A.yo[i0 + r] + j] = t[r]; else rhs_s[i0 + r] = t[r];
        yp[r] = t[r];
      }
    }
  }
}

// C' = C - Y^T Y (lower triangle, written to Cp -- which may be M itself) and g' = g_p - Y^T z (in place in rhs_p).
// Y = the first dc entries of the ns speed/bias rows of M; z = rhs_s.  8x8 tiles, one warp per tile, the contraction over
// the ns rows of Y as mma.sync.m8n8k4.f64 (fragment layout as in block_cholesky).
__device__ inline void chain_schur(const ChainView& A, double* Cp, int ns, double* rhs_p, const double* rhs_s) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int dc = A.dc;
  const double* M = A.M;
  const int nt8 = (dc + 7) >> 3;
  const int ntt8 = nt8 * (nt8 + 1) / 2;
  const int g = lane >> 2, t4 = lane & 3;
  for (int t = warp; t < ntt8; t += nwarps) {
    int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int ia = min(8 * ti + g, dc - 1), jb = min(8 * tj + g, dc - 1);     // rows beyond dc are masked at the store
    double d0 = 0.0, d1 = 0.0;
#pragma unroll 4
    for (int k0 = 0; k0 < ns; k0 += 4) {
      const int k = k0 + t4;
      const int yo = A.yo[min(k, ns - 1)];
      double a = M[yo + ia], bq = M[yo + jb];
      if (k >= ns) { a = 0.0; bq = 0.0; }
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(bq));
    }
    const int i = 8 * ti + g, j = 8 * tj + 2 * t4;
    if (i < dc) {
      if (j <= i) Cp[tri_row(i) + j] = M[tri_row(i) + j] - d0;
      if (j + 1 <= i) Cp[tri_row(i) + j + 1] = M[tri_row(i) + j + 1] - d1;
    }
  }
  for (int i = nthr - 1 - tid; i < dc; i += nthr) {     // right-hand side, on the warps that got the fewest tiles
    double s = rhs_p[i];
    for (int r = 0; r < ns; ++r) s -= M[A.yo[r] + i] * rhs_s[r];
    rhs_p[i] = s;
  }
}

// u_s = L_A^-T (z - Y u_p): x holds u_p in [0, dc); on entry x[dc + r] = z_r - (Y u_p)_r.  warp 0 only.
__device__ inline void chain_backward(const ChainView& A, int nsb, const double* rd, double* x) {
  const int lane = threadIdx.x & 31, dc = A.dc;
  const int l9 = min(lane, 8);
  for (int b = nsb - 1; b >= 0; --b) {
    const int i0 = 9 * b;
    double t = (lane < 9) ? x[dc + i0 + lane] : 0.0;
    if (b + 1 < nsb) {      // L_{b+1,b}^T u_{b+1}
#pragma unroll
      for (int k = 0; k < 9; ++k) t -= A.ab[A.ao[i0 + 9 + k] + l9] * x[dc + i0 + 9 + k];
    }
    double m[9], r9[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { m[j] = (lane < j) ? A.ab[A.ao[i0 + j] + 9 + lane] : 0.0; r9[j] = rd[i0 + j]; }
#pragma unroll
    for (int j = 8; j >= 0; --j) {
      const double uj = __shfl_sync(0xffffffffu, t, j) * r9[j];
      if (lane == j) t = uj;
      else if (lane < j) t -= m[j] * uj;
    }
    if (lane < 9) x[dc + i0 + lane] = t;
    __syncwarp();
  }
}

}  // namespace okb
