// Landmark-side kernels of the solver round (sm_100a):
//   k_linearize ("A1"): one thread per (landmark, frame).  Reprojection residuals + Jacobian factors of
//       the frame's cameras, Cauchy weighting, per-frame M_f = sum rho' A^T A and m_f = sum rho' A^T r
//       (stored, [K][L] layouts -> fully coalesced), plus the pose-block contributions
//       H_pp,f = G^T M_f G and g_p,f = G^T m_f reduced per CTA.  Light on registers: many resident
//       warps hide the FP64 dependency chains.
//   k_schur ("A2"): one CTA per (landmark chunk, window), tiles of 32 landmarks.  Landmark block
//       H_ll = sum_f M_f, (H_ll + mu E)^-1 by Cholesky, Y_f = W_f L^-T into a shared-memory tile, then the
//       Schur complement as a register-tiled dense SYRK  S += Y Y^T (4x4 micro-tiles, accumulators live
//       in registers for the whole chunk).
//   k_quality: post-solve landmark quality (Estimator.cpp:880-894), one thread per landmark.
#pragma once
#include "okb_estimator.cuh"

namespace okb {

constexpr int L1_THREADS = 128;          // k_linearize block = landmarks per CTA (x one frame)
constexpr int A2_THREADS = 288;          // k_schur block (9 warps: 2 k-splits of 136 micro-tiles at K=10)
constexpr int A2_TILE = 32;              // landmarks per Y tile
constexpr int kPartH = 32;               // doubles per (cx, frame) record: 27 H_pp/g_p + cost + stepnorm2 + pad

struct SlotCtx {
  SlotXf xf;
  CamIntr cam;
  int frame;
  int valid;
};

// ------------------------------------------------------------------------------------------------
// A1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(L1_THREADS) k_linearize(const WinDev* __restrict__ wins, int win_first) {
  const WinDev& W = wins[win_first + blockIdx.z];
  const SolverState* st = W.st;
  if (st->done) return;
  const int f = blockIdx.y;
  const int cx = blockIdx.x;
  if (f >= W.K || cx * L1_THREADS >= W.L) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = W.L, CP = W.CP;

  __shared__ SlotCtx slots[32];                 // the frame's cameras (CP <= 32)
  __shared__ double sred[kPartH][L1_THREADS + 1];
  __shared__ double s_tw[3];

  if (tid < CP) {
    const SlotInfo si = W.slots[f * CP + tid];
    SlotCtx& sc = slots[tid];
    sc.valid = si.valid; sc.frame = f;
    if (si.valid) { make_slot_xf(W.pose_c + 7 * si.pose_idx, W.ext + 7 * si.ext_idx, sc.xf); cam_load(W.cams[si.cam_idx], sc.cam); }
  }
  if (tid < 3) s_tw[tid] = W.pose_c[7 * f + tid];
  __syncthreads();

  const int mode = st->mode, cur = st->cur;
  const bool cauchy = W.use_cauchy != 0;
  const int l = cx * L1_THREADS + tid;
#pragma unroll
  for (int i = 0; i < 29; ++i) sred[i][tid] = 0.0;   // contributions go straight to shared memory (no live registers)

  if (l < L) {
    const bool vis = (W.lm_vis[l] >> f) & 1u;
    if (vis || f == 0) {
      double X[4];
      {
        const double4 x4 = *reinterpret_cast<const double4*>(W.lm + 4 * (size_t)l);
        X[0] = x4.x; X[1] = x4.y; X[2] = x4.z; X[3] = x4.w;
      }
      if (mode == MODE_STEP) {
        const double a = st->a, b = st->b;
        const double* g = W.lm_g[cur] + 3 * (size_t)l;
        const double* E = W.lm_E[cur] + 3 * (size_t)l;
        const double* gn = W.lm_gn + 3 * (size_t)l;
        double dn = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double dlt = a * g[c] / E[c] + b * gn[c];
          X[c] += dlt;
          dn += dlt * dlt;
        }
        if (f == 0) sred[28][tid] = dn;
      }
      if (f == 0) *reinterpret_cast<double4*>(W.lm_c + 4 * (size_t)l) = make_double4(X[0], X[1], X[2], X[3]);
      if (vis) {
        double M0 = 0, M1 = 0, M2 = 0, M3 = 0, M4 = 0, M5 = 0, m0 = 0, m1 = 0, m2 = 0, cost = 0;
        // the first two cameras' observations are fetched up front so that their latency overlaps
        double pw[2] = {0.0, 0.0};
        double2 pz[2] = {make_double2(0, 0), make_double2(0, 0)};
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < CP) { const size_t gi = (size_t)(f * CP + c) * L + l; pw[c] = W.obs_w[gi]; pz[c] = W.obs_z[gi]; }
        for (int c = 0; c < CP; ++c) {
          const SlotCtx& sc = slots[c];
          const size_t gi = (size_t)(f * CP + c) * L + l;
          const double wobs = (c < 2) ? pw[c & 1] : W.obs_w[gi];
          if (wobs > 0.0 && sc.valid) {
            const double2 z = (c < 2) ? pz[c & 1] : W.obs_z[gi];
            double r[2], A[6];
            reproj_slot<true>(sc.xf, sc.cam, X, z.x, z.y, wobs, r, A);
            const double sq = r[0] * r[0] + r[1] * r[1];
            double rho1 = 1.0;
            if (cauchy) { rho1 = 1.0 / (1.0 + sq); cost += 0.5 * log(1.0 + sq); }
            else cost += 0.5 * sq;
            M0 += rho1 * (A[0] * A[0] + A[3] * A[3]);
            M1 += rho1 * (A[0] * A[1] + A[3] * A[4]);
            M2 += rho1 * (A[0] * A[2] + A[3] * A[5]);
            M3 += rho1 * (A[1] * A[1] + A[4] * A[4]);
            M4 += rho1 * (A[1] * A[2] + A[4] * A[5]);
            M5 += rho1 * (A[2] * A[2] + A[5] * A[5]);
            m0 += rho1 * (A[0] * r[0] + A[3] * r[1]);
            m1 += rho1 * (A[1] * r[0] + A[4] * r[1]);
            m2 += rho1 * (A[2] * r[0] + A[5] * r[1]);
          }
        }
        double* Mo = W.lm_M + ((size_t)f * L + l) * 6;
        Mo[0] = M0; Mo[1] = M1; Mo[2] = M2; Mo[3] = M3; Mo[4] = M4; Mo[5] = M5;
        double* mo = W.lm_mf + ((size_t)f * L + l) * 3;
        mo[0] = m0; mo[1] = m1; mo[2] = m2;
        // pose-block contributions: G = [w I, -[p]x], p = X - t_WS w
        const double w = X[3];
        const double p0 = X[0] - s_tw[0] * w, p1 = X[1] - s_tw[1] * w, p2 = X[2] - s_tw[2] * w;
        const double Q00 = M1 * p2 - M2 * p1, Q01 = -M0 * p2 + M2 * p0, Q02 = M0 * p1 - M1 * p0;
        const double Q10 = M3 * p2 - M4 * p1, Q11 = -M1 * p2 + M4 * p0, Q12 = M1 * p1 - M3 * p0;
        const double Q20 = M4 * p2 - M5 * p1, Q21 = -M2 * p2 + M5 * p0, Q22 = M2 * p1 - M4 * p0;
        const double w2 = w * w;
        sred[0][tid] = w2 * M0; sred[1][tid] = w2 * M1; sred[2][tid] = w2 * M2; sred[3][tid] = w2 * M3; sred[4][tid] = w2 * M4; sred[5][tid] = w2 * M5;
        sred[6][tid] = -w * Q00; sred[7][tid] = -w * Q01; sred[8][tid] = -w * Q02;
        sred[9][tid] = -w * Q10; sred[10][tid] = -w * Q11; sred[11][tid] = -w * Q12;
        sred[12][tid] = -w * Q20; sred[13][tid] = -w * Q21; sred[14][tid] = -w * Q22;
        sred[15][tid] = p2 * Q10 - p1 * Q20; sred[16][tid] = p2 * Q11 - p1 * Q21; sred[17][tid] = p2 * Q12 - p1 * Q22;
        sred[18][tid] = -p2 * Q01 + p0 * Q21; sred[19][tid] = -p2 * Q02 + p0 * Q22;
        sred[20][tid] = p1 * Q02 - p0 * Q12;
        sred[21][tid] = w * m0; sred[22][tid] = w * m1; sred[23][tid] = w * m2;
        sred[24][tid] = p1 * m2 - p2 * m1; sred[25][tid] = p2 * m0 - p0 * m2; sred[26][tid] = p0 * m1 - p1 * m0;
        sred[27][tid] = cost;
      }
    }
  }
  // ---- CTA reduction of the 29 used values (fixed order -> deterministic)
  __syncthreads();
  for (int e = warp; e < 29; e += L1_THREADS / 32) {
    double s = sred[e][lane] + sred[e][lane + 32] + sred[e][lane + 64] + sred[e][lane + 96];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) W.partH[((size_t)cx * W.K + f) * kPartH + e] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Landmark blocks: one thread per landmark.  H_ll = sum_f M_f, g_l = -sum_f m_f, the metric E_l,
// R = H_ll + mu E and its Cholesky factor: stores R^-1 (back-substitution), L^-1 (Schur tile),
// g_l, E_l and z = L^-1 g_l.  [f][l] layouts make every load/store coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_lmblock(const WinDev* __restrict__ wins, int win_first) {
  const WinDev& W = wins[win_first + blockIdx.y];
  SolverState* st = W.st;
  if (st->done) return;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = W.L, K = W.K;
  if (l >= L) return;
  const int mode = st->mode, cur = st->cur;
  const double mu = (mode == MODE_STEP) ? fmax(kMinMu, 2.0 * st->mu / kMuIncrease) : st->mu;
  double H[6] = {0, 0, 0, 0, 0, 0}, gl[3] = {0, 0, 0};
  uint32_t vis = W.lm_vis[l];
  for (int f = 0; f < K; ++f) {
    if ((vis >> f) & 1u) {
      const double* Mo = W.lm_M + ((size_t)f * L + l) * 6;
      const double* mo = W.lm_mf + ((size_t)f * L + l) * 3;
#pragma unroll
      for (int i = 0; i < 6; ++i) H[i] += Mo[i];
      gl[0] -= mo[0]; gl[1] -= mo[1]; gl[2] -= mo[2];
    }
  }
  double sc3[3], E[3];
  const double hd[3] = {H[0], H[3], H[5]};
  if (mode == MODE_INIT) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { sc3[c] = 1.0 / (1.0 + sqrt(hd[c])); W.lm_scale[3 * (size_t)l + c] = sc3[c]; }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) sc3[c] = W.lm_scale[3 * (size_t)l + c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double s2 = sc3[c] * sc3[c];
    E[c] = fmin(fmax(s2 * hd[c], kMinDiag), kMaxDiag) / s2;
  }
  const double R[6] = {H[0] + mu * E[0], H[1], H[2], H[3] + mu * E[1], H[4], H[5] + mu * E[2]};
  double Lc[6], Li[6];
  if (chol3(R, Lc)) linv3(Lc, Li);
  else {
#pragma unroll
    for (int i = 0; i < 6; ++i) Li[i] = 0.0;
    st->numeric_fail = 1;
  }
  double* Ro = W.lm_Rinv + 6 * (size_t)l;
  Ro[0] = Li[0] * Li[0] + Li[1] * Li[1] + Li[3] * Li[3];
  Ro[1] = Li[1] * Li[2] + Li[3] * Li[4];
  Ro[2] = Li[3] * Li[5];
  Ro[3] = Li[2] * Li[2] + Li[4] * Li[4];
  Ro[4] = Li[4] * Li[5];
  Ro[5] = Li[5] * Li[5];
  double* gs = W.lm_g[cur ^ 1] + 3 * (size_t)l;
  double* Es = W.lm_E[cur ^ 1] + 3 * (size_t)l;
#pragma unroll
  for (int c = 0; c < 3; ++c) { gs[c] = gl[c]; Es[c] = E[c]; }
  double* Lo = W.lm_Li + 9 * (size_t)l;     // L^-1 (6) | z (3)
#pragma unroll
  for (int i = 0; i < 6; ++i) Lo[i] = Li[i];
  Lo[6] = Li[0] * gl[0];
  Lo[7] = Li[1] * gl[0] + Li[2] * gl[1];
  Lo[8] = Li[3] * gl[0] + Li[4] * gl[1] + Li[5] * gl[2];
}

// ------------------------------------------------------------------------------------------------
// A2
// ------------------------------------------------------------------------------------------------
constexpr int kMStride = 7;   // doubles per (landmark, frame) M block in shared memory (6 + 1 pad: 2-way bank conflicts at most)
constexpr int kLiStride = 10;  // L^-1 (6) | z (3) | pad

__host__ __device__ inline size_t smemA2_bytes(int K, int dcp) {
  size_t b = 0;
  b += (size_t)3 * A2_TILE * dcp * sizeof(double);                 // Y tile, k-major
  b += (size_t)2 * K * A2_TILE * kMStride * sizeof(double);        // M tiles (double buffered)
  b += (size_t)2 * A2_TILE * kLiStride * sizeof(double);           // L^-1 | z
  b += (size_t)2 * A2_TILE * 4 * sizeof(double);                   // X
  b += (size_t)K * 4 * sizeof(double);                             // frame translations
  return b;
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int TPT>
__global__ void __launch_bounds__(A2_THREADS, 2) k_schur(const WinDev* __restrict__ wins, int win_first) {
  const WinDev& W = wins[win_first + blockIdx.y];
  SolverState* st = W.st;
  if (st->done) return;
  const int chunk = blockIdx.x;
  if (chunk >= W.n_chunks) return;
  const int tid = threadIdx.x;
  const int K = W.K, dc = W.dc, dcp = W.dcp, L = W.L;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Yt = reinterpret_cast<double*>(smem_raw);
  double* sMb = Yt + (size_t)3 * A2_TILE * dcp;
  double* sLib = sMb + (size_t)2 * K * A2_TILE * kMStride;
  double* sXb = sLib + 2 * A2_TILE * kLiStride;
  double* tws = sXb + 2 * A2_TILE * 4;

  for (int f = tid; f < K; f += A2_THREADS) {
    tws[4 * f] = W.pose_c[7 * f]; tws[4 * f + 1] = W.pose_c[7 * f + 1]; tws[4 * f + 2] = W.pose_c[7 * f + 2];
  }

  // ---- SYRK thread mapping (4x4 micro-tiles of the lower triangle of the dcp x dcp matrix)
  const int NT = dcp >> 2;
  const int NTT = NT * (NT + 1) / 2;
  int KS = 1;
  if (TPT == 1) { KS = A2_THREADS / NTT; if (KS < 1) KS = 1; }
  const int ks = (TPT == 1) ? tid / NTT : 0;
  int ti[TPT], tj[TPT];
  bool syrk_on[TPT];
  double acc[TPT][16];
#pragma unroll
  for (int m = 0; m < TPT; ++m) {
    const int tt = (TPT == 1) ? tid % NTT : tid + m * A2_THREADS;
    syrk_on[m] = (TPT == 1) ? (ks < KS) : (tt < NTT);
    const int tq = syrk_on[m] ? tt : 0;
    int a = (int)((sqrt(8.0 * tq + 1.0) - 1.0) * 0.5);
    while (a * (a + 1) / 2 > tq) --a;
    while ((a + 1) * (a + 2) / 2 <= tq) ++a;
    ti[m] = a;
    tj[m] = tq - a * (a + 1) / 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[m][i] = 0.0;
  }

  const int lm_begin = chunk * W.lm_per_chunk;
  const int lm_end = min(L, lm_begin + W.lm_per_chunk);

  // asynchronous staging of one tile (cp.async, 8-byte elements, no registers held):
  //   M  [f][l][6] (global, zero where unobserved) -> [ll][f][kMStride]
  //   Li [l][9] -> [ll][kLiStride],  X = lm_c [l][4] -> [ll][4]
  auto stage = [&](int base, int buf) {
    double* sM = sMb + (size_t)buf * K * A2_TILE * kMStride;
    double* sLi = sLib + buf * A2_TILE * kLiStride;
    double* sX = sXb + buf * A2_TILE * 4;
    const int nl = min(A2_TILE, lm_end - base);
    for (int i = tid; i < K * A2_TILE * 6; i += A2_THREADS) {
      const int f = i / (A2_TILE * 6), r = i % (A2_TILE * 6);
      const int ll = r / 6, e = r % 6;
      if (ll < nl) cp_async8(sM + ((size_t)ll * K + f) * kMStride + e, W.lm_M + ((size_t)f * L + base) * 6 + r);
    }
    for (int i = tid; i < A2_TILE * 9; i += A2_THREADS) {
      const int ll = i / 9, e = i % 9;
      if (ll < nl) cp_async8(sLi + ll * kLiStride + e, W.lm_Li + 9 * (size_t)base + i);
    }
    for (int i = tid; i < A2_TILE * 4; i += A2_THREADS) {
      const int ll = i / 4;
      if (ll < nl) cp_async8(sX + i, W.lm_c + 4 * (size_t)base + i);
    }
    cp_async_commit();
  };

  int buf = 0;
  if (lm_begin < lm_end) stage(lm_begin, 0);
  for (int base = lm_begin; base < lm_end; base += A2_TILE) {
    const int nl = min(A2_TILE, lm_end - base);
    const double* sM = sMb + (size_t)buf * K * A2_TILE * kMStride;
    const double* sLi = sLib + buf * A2_TILE * kLiStride;
    const double* sX = sXb + buf * A2_TILE * 4;
    cp_async_wait_all();
    __syncthreads();                       // tile `buf` is complete; the previous SYRK has finished
    if (base + A2_TILE < lm_end) stage(base + A2_TILE, buf ^ 1);   // overlaps with this tile's math
    // ---- augmented row z and the padding rows of the Y tile
    if (tid < A2_TILE) {
      const int ll = tid;
      double* yz = Yt + (size_t)(3 * ll) * dcp + dc;
      if (ll < nl) { yz[0] = sLi[ll * kLiStride + 6]; yz[dcp] = sLi[ll * kLiStride + 7]; yz[2 * dcp] = sLi[ll * kLiStride + 8]; }
      else { yz[0] = 0.0; yz[dcp] = 0.0; yz[2 * dcp] = 0.0; }
      for (int r = dc + 1; r < dcp; ++r) { Yt[(size_t)(3 * ll) * dcp + r] = 0.0; Yt[(size_t)(3 * ll + 1) * dcp + r] = 0.0; Yt[(size_t)(3 * ll + 2) * dcp + r] = 0.0; }
    }
    // ---- Y_f = W_f L^-T for every (landmark, frame) pair of the tile
    for (int pidx = tid; pidx < A2_TILE * K; pidx += A2_THREADS) {
      const int ll = pidx / K, f = pidx % K;
      double* y0 = Yt + (size_t)(3 * ll) * dcp + 6 * f;
      double* y1 = y0 + dcp;
      double* y2 = y1 + dcp;
      double M0 = 0, M1 = 0, M2 = 0, M3 = 0, M4 = 0, M5 = 0;
      if (ll < nl) {
        const double* sp = sM + ((size_t)ll * K + f) * kMStride;
        M0 = sp[0]; M1 = sp[1]; M2 = sp[2]; M3 = sp[3]; M4 = sp[4]; M5 = sp[5];
      }
      if (M0 != 0.0 || M3 != 0.0 || M5 != 0.0) {
        const double* Li = sLi + ll * kLiStride;
        const double w = sX[ll * 4 + 3];
        const double p0 = sX[ll * 4] - tws[4 * f] * w, p1 = sX[ll * 4 + 1] - tws[4 * f + 1] * w, p2 = sX[ll * 4 + 2] - tws[4 * f + 2] * w;
        const double N00 = M0 * Li[0], N01 = M0 * Li[1] + M1 * Li[2], N02 = M0 * Li[3] + M1 * Li[4] + M2 * Li[5];
        const double N10 = M1 * Li[0], N11 = M1 * Li[1] + M3 * Li[2], N12 = M1 * Li[3] + M3 * Li[4] + M4 * Li[5];
        const double N20 = M2 * Li[0], N21 = M2 * Li[1] + M4 * Li[2], N22 = M2 * Li[3] + M4 * Li[4] + M5 * Li[5];
        y0[0] = -w * N00; y0[1] = -w * N10; y0[2] = -w * N20;
        y1[0] = -w * N01; y1[1] = -w * N11; y1[2] = -w * N21;
        y2[0] = -w * N02; y2[1] = -w * N12; y2[2] = -w * N22;
        y0[3] = -(p1 * N20 - p2 * N10); y0[4] = -(p2 * N00 - p0 * N20); y0[5] = -(p0 * N10 - p1 * N00);
        y1[3] = -(p1 * N21 - p2 * N11); y1[4] = -(p2 * N01 - p0 * N21); y1[5] = -(p0 * N11 - p1 * N01);
        y2[3] = -(p1 * N22 - p2 * N12); y2[4] = -(p2 * N02 - p0 * N22); y2[5] = -(p0 * N12 - p1 * N02);
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) { y0[i] = 0.0; y1[i] = 0.0; y2[i] = 0.0; }
      }
    }
    __syncthreads();
    // ---- SYRK over the tile
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      if (syrk_on[m]) {
        const int ncols = 3 * A2_TILE;
#pragma unroll 2
        for (int k = ks; k < ncols; k += KS) {
          const double* row = Yt + (size_t)k * dcp;
          const double2 a01 = *reinterpret_cast<const double2*>(row + 4 * ti[m]);
          const double2 a23 = *reinterpret_cast<const double2*>(row + 4 * ti[m] + 2);
          const double2 b01 = *reinterpret_cast<const double2*>(row + 4 * tj[m]);
          const double2 b23 = *reinterpret_cast<const double2*>(row + 4 * tj[m] + 2);
          const double a[4] = {a01.x, a01.y, a23.x, a23.y};
          const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[m][i * 4 + j] += a[i] * b[j];
        }
      }
    }
    buf ^= 1;
  }
  __syncthreads();

  // ---- epilogue: chunk partial of the Schur accumulator (deterministic k-split order)
  double* Sp = W.partA + (size_t)chunk * W.partA_stride;
  for (int s = 0; s < KS; ++s) {
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      if (syrk_on[m] && ks == s) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double* p = Sp + (size_t)(4 * ti[m] + i) * dcp + 4 * tj[m] + j;
            *p = (s == 0) ? acc[m][i * 4 + j] : (*p + acc[m][i * 4 + j]);
          }
      }
    }
    __syncthreads();
  }
}

// Sums the per-chunk Schur accumulators of a window into chunk 0 (fixed order => deterministic), in
// parallel over the matrix elements, so that k_solve reads one partial regardless of the chunk count.
__global__ void __launch_bounds__(256) k_reduce_partials(const WinDev* __restrict__ wins, int win_first) {
  const WinDev& W = wins[win_first + blockIdx.y];
  if (W.st->done || W.n_chunks <= 1) return;
  const int n = W.dcp * W.dcp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = W.partA[i];
    for (int c = 1; c < W.n_chunks; ++c) s += W.partA[(size_t)c * W.partA_stride + i];
    W.partA[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Post-solve landmark quality: H = sum J_lm^T J_lm (sqrt-information weighted, no robust weight)
// at the final estimate; quality = sqrt(lambda_min)/sqrt(lambda_max), 0 if lambda_min < 1e-12.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_quality(const WinDev* __restrict__ wins, int win_first) {
  const WinDev& W = wins[win_first + blockIdx.y];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotCtx* slots = reinterpret_cast<SlotCtx*>(smem_raw);
  const int tid = threadIdx.x;
  for (int s = tid; s < W.NS; s += blockDim.x) {
    const SlotInfo si = W.slots[s];
    SlotCtx& sc = slots[s];
    sc.valid = si.valid; sc.frame = si.valid ? si.pose_idx : 0;
    if (si.valid) { make_slot_xf(W.pose + 7 * si.pose_idx, W.ext + 7 * si.ext_idx, sc.xf); cam_load(W.cams[si.cam_idx], sc.cam); }
  }
  __syncthreads();
  const int L = W.L, CP = W.CP;
  for (int l = blockIdx.x * blockDim.x + tid; l < L; l += gridDim.x * blockDim.x) {
    const double4 x4 = *reinterpret_cast<const double4*>(W.lm + 4 * (size_t)l);
    const double X[4] = {x4.x, x4.y, x4.z, x4.w};
    double H[6] = {0, 0, 0, 0, 0, 0};
    uint32_t vis = W.lm_vis[l];
    while (vis) {
      const int f = __ffs(vis) - 1;
      vis &= vis - 1;
      for (int c = 0; c < CP; ++c) {
        const int s = f * CP + c;
        const double wobs = W.obs_w[(size_t)s * L + l];
        if (wobs > 0.0 && slots[s].valid) {
          const double2 z = W.obs_z[(size_t)s * L + l];
          double r[2], A[6];
          reproj_slot<true>(slots[s].xf, slots[s].cam, X, z.x, z.y, wobs, r, A);
          H[0] += A[0] * A[0] + A[3] * A[3]; H[1] += A[0] * A[1] + A[3] * A[4]; H[2] += A[0] * A[2] + A[3] * A[5];
          H[3] += A[1] * A[1] + A[4] * A[4]; H[4] += A[1] * A[2] + A[4] * A[5]; H[5] += A[2] * A[2] + A[5] * A[5];
        }
      }
    }
    double ev[3];
    eig3sym_closed(H, ev);
    W.quality[l] = (ev[0] < 1.0e-12) ? 0.0 : sqrt(ev[0]) / sqrt(ev[2]);
  }
}

}  // namespace okb
