// Device-side data model of one keyframe window and of the batched dogleg/Schur solver.
// See DESIGN.md ("Data layout in HBM", "Kernels") for the rationale.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "okb_imu.cuh"

namespace okb {

// ---- compile-time limits ------------------------------------------------------------------
constexpr int kMaxFrames = 32;        // frame visibility mask is one u32
constexpr int kMaxChunks = 64;        // landmark chunks (CTAs) per window in kernel A
constexpr int kMaxDense = 512;        // reduced-system dimension limit
constexpr int kMaxMarg = 160;         // marginalisation prior dimension limit
constexpr int kMaxMargBlocks = 64;    // parameter blocks connected to the marginalisation prior
constexpr int kMaxShard = 8;          // ranks of a landmark-sharded window (one NVSwitch domain)

// Ceres 1.9 defaults used by Estimator::optimize
constexpr double kInitialRadius = 1e4;
constexpr double kMinRadius = 1e-32;
constexpr double kMinRelativeDecrease = 1e-3;
constexpr double kFunctionTolerance = 1e-6;
constexpr double kGradientTolerance = 1e-10;
constexpr double kParameterTolerance = 1e-8;
constexpr double kMinDiag = 1e-6;
constexpr double kMaxDiag = 1e32;
constexpr double kMinMu = 1e-8;
constexpr double kMaxMu = 1.0;
constexpr double kMuIncrease = 10.0;
constexpr int kMaxConsecutiveInvalid = 5;

enum Mode { MODE_INIT = 0, MODE_STEP = 1, MODE_REBUILD = 2 };

struct SlotInfo { int pose_idx, ext_idx, cam_idx, valid; };

// Dimensions of the device-resident window graph as the command interpreter left them (okb_graph.cuh); the host
// mirrors the same arithmetic, k_compile_obs cross-checks.  err: first GERR_* raised by a command or by the compile.
struct GraphState { int K, NSB, L, n_obs, n_imu, n_samples, n_pp, n_sbp, marg_n, marg_nb, marg_xdim, NE, NC, err; };

// Per-window solver state (device resident, one per window; also read back as the summary).
struct SolverState {
  int mode;            // what kernel A must evaluate next / what kernel S must judge
  int done;            // 1 = terminated
  int cur;             // buffer index of the linearisation at the committed state (lm_g / lm_E / gd / Ed)
  int iteration, num_successful, num_invalid, termination;
  int reuse;           // dogleg reuse_ flag
  int numeric_fail;    // set by kernel A when a landmark block is not positive definite
  int imu_redo;
  int imu_redo_final;
  double radius, mu, mu_spec;
  double cost, initial_cost;
  double x_norm2;      // ||x||^2 of the committed state (ambient)
  // dogleg scalars of the current linearisation (metric E):
  double G2, VHV, GU, N2;
  double a, b;         // step = a * g/E + b * (-u)
  double model_cost_change, dogleg_step_norm;
  double cand_step_norm2_dense;   // ||x - x_cand||^2 over the dense blocks
  double grad_max;
  unsigned long long t_start_ns, t_last_iter_ns, t_iter_begin_ns;
  double solve_time_s;
  unsigned long long phase_ns[16];  // accumulated phase times (diagnostics): 0..7 k_solve, 8..13 k_schur (chunk 0, thread 0)
  // Landmark-sharded windows: exchange epoch (one per solver round; never reset, so stale mailbox flags of an earlier
  // optimize can not match) and the accumulated device time spent waiting for / summing the peers' partial systems.
  unsigned long long shard_epoch;
  unsigned long long shard_wait_ns, shard_rounds;
  int shard_fault;     // 1 = a peer did not arrive within the time-out (the window terminates with FAILURE)
  GraphState g;
};

// ---- landmark-sharded single window: per-window mailbox in every rank's device memory.
// Rank s PUSHES its reduced landmark partial (Schur accumulator, pose-block sums, cost) into box [parity][s] of every
// rank's mailbox with plain stores over NVLink, then releases flag [parity][s]; the consumer (k_solve) acquires the
// `world` flags in its own memory and adds the boxes in rank order, so every rank forms the bit-identical reduced
// system.  Boxes are double-buffered by the parity of the round's epoch ("Multi-GPU").
constexpr size_t kShardHeaderBytes = 4096;
constexpr size_t kShardFlags1 = 0;       // u64 [2][kMaxShard]   reduced-system boxes ready
constexpr size_t kShardFlags2 = 128;     // u64 [2][kMaxShard]   step scalars ready
constexpr size_t kShardCounter = 256;    // u32                  last-CTA counter of k_shard_push (local use)
constexpr size_t kShardScalars = 512;    // f64 [2][kMaxShard][16]
__host__ __device__ inline size_t shard_box_doubles(int K, int dcp) { return (size_t)dcp * dcp + (size_t)K * 32 + 8; }
__host__ __device__ inline size_t shard_win_bytes(int world, size_t box_cap) {
  return (kShardHeaderBytes + 2 * (size_t)world * box_cap * sizeof(double) + 255) & ~(size_t)255;
}

// Everything kernels need to know about one window.  All pointers are device pointers into the
// window's arena.
// lm_M / lm_mf are tile-major: [tile of 32 landmarks][frame][element][32] -- a warp's accesses are 256-byte rows and
// the frames [a, b] of one tile are ONE contiguous block (a single TMA bulk copy in k_schur).
__host__ __device__ inline size_t lm_M_index(int l, int f, int K) { return ((size_t)(l >> 5) * K + f) * 192 + (l & 31); }   // + 32 * element (0..5)
__host__ __device__ inline size_t lm_mf_index(int l, int f, int K) { return ((size_t)(l >> 5) * K + f) * 96 + (l & 31); }    // + 32 * element (0..2)

// packed download block: pose [K][7] | speed/bias [NSB][9] | pad to 4 doubles | landmarks [L][4] | quality [L]
// (the landmark part is written with 32-byte vector stores)
__host__ __device__ inline size_t out_lm_offset(int K, int NSB) { return ((size_t)7 * K + (size_t)9 * NSB + 3) & ~(size_t)3; }

struct SlotCtx;   // per (frame, camera) transform + intrinsics at the candidate state (okb_kernels_lm.cuh)

struct WinDev {
  // capacities the arena was planned for (okb_window_reserve / the largest upload); the current dimensions below
  // change with the graph commands
  int Kcap, Lcap, Ocap, Scap, Tcap, NEcap, NCcap, PPcap;
  // master graph in the caller's index space (okb_graph.cuh); pose / sb / ext below are master and working copy at once
  double* m_lm;                                // [Lcap][4] landmark estimates, caller's order (written back after every optimize)
  double* m_lm_init;                           // [Lcap][4] as of the last full upload (okb_window_reset)
  unsigned char* m_mark;                       // [Lcap] scratch marks of the command interpreter (all zero between commands)
  okb_observation* m_obs;                      // [Ocap] observation list; sqrt_info == 0: removed, compacted by k_compile_obs
  uint32_t* m_vis;                             // [Lcap] frame-visibility masks, caller's order (compile scratch)
  uint32_t* m_bitmap;                          // one bit per observation-grid cell (duplicate detection)
  uint32_t* perm;                              // [Lcap] internal (sorted) landmark index -> caller's index
  double* out;                                 // packed estimates for the download: pose [K][7] | sb [NSB][9] | landmarks [L][4] | quality [L]
  const unsigned char* cmd;                    // command buffer of the pending commit
  int cmd_bytes;
  int dirty;                                   // 1 = the graph changed: compile before the next optimize
  int full;                                    // 1 = the pending commit is a full upload: snapshot the initial state
  int K, NSB, NE, L, NC;
  int Lp;              // L rounded up to whole tiles of 32: leading dimension of lm_M / lm_mf, rows of lm_Li / lm_c (bulk copies stay aligned)
  int CP;              // cameras per frame padded to a power of two
  int NS;              // slots = K * CP  (slot = frame * CP + cam)
  int NG;              // slot groups of 32 lanes
  int NSP;             // NG * 32: padded slots per landmark in the observation grid
  int d, dc, dcp;      // reduced dims: dc = 6K (pose part), d = dc + 9 NSB, dcp = 4*ceil((dc+1)/4)
  int n_imu, n_samples, n_pp, n_sbp;
  int marg_n, marg_nb, marg_xdim;
  int n_chunks, lm_per_chunk;
  int use_cauchy;
  // state
  double *pose, *sb, *ext, *lm;                // committed  [K][7] [NSB][9] [NE][7] [L][4]
  double *pose_init, *sb_init;                 // as of the last full upload (okb_window_reset); landmarks: m_lm_init
  double *pose_c, *sb_c, *lm_c;                // candidate
  // graph
  SlotInfo* slots;                             // [NSP]
  okb_camera* cams;                            // [NC]
  double2* obs_z;                              // [NS][L]  slot-major: coalesced for thread-per-landmark kernels
  double* obs_w;                               // [NS][L] sqrt information, 0 = no observation
  uint32_t* lm_vis;                            // [L] bit f set = observed in frame f
  int n_obs;                                   // host's upper bound of the list length (launch sizing); exact: st->g.n_obs
  // Landmarks are stored sorted by (first, last) observing frame (okb_window_upload): tracks are runs of
  // consecutive frames, so neighbouring landmarks see nearly the same frames -> warp-coherent visibility
  // in k_linearize and block-sparse Schur tiles in k_schur.
  unsigned char* zero_ptr[3];                  // regions cleared by k_zero at every compile: observation grid (z, w), M blocks
  size_t zero_bytes[3];
  uint32_t* lm_inv;                            // [L] caller's landmark index -> internal (sorted) index
  uint32_t* tile_range;                  // [ceil(L/32)] frames seen by the tile: first | last << 8 (first > last: none)
  // per-landmark solver data
  double* lm_g[2];                             // [L][3] gradient block (double buffered: cur / speculative)
  double* lm_E[2];                             // [L][3] metric (Ceres diagonal^2 / scale^2)
  double* lm_Rinv;                             // [L][6] (H_ll + mu E)^-1, symmetric packed
  double* lm_M;                                // [Lp/32][K][6][32] per-frame sum of rho' A^T A (lm_M_index)
  double* lm_mf;                               // [Lp/32][K][3][32] per-frame sum of rho' A^T r (lm_mf_index)
  SlotCtx* slot_ctx;                           // [NS] built by k_reset / k_solve whenever the candidate poses change
  double* lm_Li;                               // [L][9] L^-1 of (H_ll + mu E) (6) and z = L^-1 g_l (3)
  double* lm_gn;                               // [L][3] Gauss-Newton step of the current linearisation
  double* lm_scale;                            // [L][3] Jacobi scale (fixed after the first linearisation)
  double* quality;                             // [L]
  // kernel A -> kernel S partial sums, one record per chunk
  double* partA;                               // [n_chunks][dcp*dcp] Schur accumulator per chunk (k_schur)
  int partA_stride;                            // dcp*dcp
  double* partH;                               // [ceil(L/128)][K][32] pose-block sums, cost, step norm (k_linearize)
  // dense part
  double* Hd;                                  // [d][d] J^T J restricted to dense blocks (speculative)
  double* gd[2];                               // [d]
  double* Ed[2];                               // [d]
  double* ud;                                  // [d] (H+mu E)^-1 g, dense part
  double* scale_d;                             // [d]
  double* chol;                                // [d][d] workspace (used when it does not fit in shared memory)
  // IMU
  okb_imu_term* imu_terms;
  okb_imu_sample* samples;
  okb_imu_params imu_params;
  ImuCache* imu_cache;                         // all zero (valid = 0) after an upload / okb_window_reset
  double* imu_out;                             // [n_imu][kImuOut] written by k_imu, consumed by k_solve
  int *sb_off;                                 // unused placeholder for alignment
  // priors
  okb_pose_prior* pp;
  okb_sb_prior* sbp;
  // marginalisation prior
  int32_t* marg_kind;
  uint32_t* marg_idx;
  int32_t* marg_col;                           // [marg_nb] first column of each block
  int32_t* marg_off;                           // [marg_nb] offset into x0
  double *marg_x0, *marg_J, *marg_e0, *marg_H0;  // H0 = J^T J
  double *marg_Hs, *marg_b0;                   // H / b0 as MarginalizationError keeps them between calls (okb_marg.cuh); host-supplied priors: J^T J, -J^T e0
  // landmark sharding: this rank holds the landmarks / observations of its shard and all dense blocks
  int shard_rank, shard_world;                 // world <= 1: not sharded
  int shard_box_cap;                           // doubles per box
  unsigned char* shard_mail[kMaxShard];        // this window's mailbox in every rank's memory, as mapped on this device
  SolverState* st;
};

}  // namespace okb
