// Opaque context behind the C-ABI handle (host side).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/okvis_b200.h"
#include "okb_estimator.cuh"

// Capacities an arena is planned for: frames, landmarks, observations, IMU samples / terms, extrinsics, cameras,
// priors per kind, marginalisation prior dimension.
struct WinCaps { int K = 0, L = 0, O = 0, S = 0, T = 0, NE = 0, NC = 0, PP = 0, MN = 0; };

struct WinStore {
  unsigned char* arena = nullptr;     // device
  size_t arena_bytes = 0;
  WinCaps caps, reserve;              // what the arena holds / the minimum asked for by okb_window_reserve
  unsigned char* staging = nullptr;   // pinned: the pending command stream of the slot (okb_graph.cuh)
  size_t staging_bytes = 0;
  size_t cmd_used = 0;                // bytes of commands not yet committed
  size_t cmd_cap_dev = 0;             // size of the device-side command buffer
  bool staging_busy = false;          // an H2D copy of `staging` may still be in flight (wait on `copied` before rewriting)
  bool full_pending = false;          // the pending stream is a full upload
  bool committed = false;
  int obs_bound = 0;                  // upper bound of the device-side observation list length
  int K_init = 0, NSB_init = 0, L_init = 0;   // shape at the last full upload (okb_window_reset)
  std::vector<okb_imu_term> terms;    // host mirror of the IMU terms / prior indices (dimension bookkeeping of remove_frame)
  std::vector<uint32_t> pp_idx, sbp_idx;
  size_t h2d_bytes = 0;
  bool uploaded = false;
  cudaEvent_t copied = nullptr;       // recorded after the H2D copy of `staging`
  unsigned char* out_staging = nullptr;  // pinned host buffer the estimates are downloaded into
  size_t out_bytes = 0;
  cudaEvent_t down = nullptr;         // recorded after the D2H copies of a download
  okb_solve_options opt{};            // options of the last okb_optimize_async on this slot
  int done_idx = -1;                  // index into okb_ctx::done_ring of the last solver work launched on this slot
  unsigned char* marg_scratch = nullptr;   // device scratch of okb_window_marginalize (okb_marg.cuh), allocated on first use
  size_t marg_scratch_bytes = 0;
  int marg_lm_cap = 0, marg_K_cap = 0, marg_L_cap = 0, marg_O_cap = 0;
};

struct okb_frontend_state;

struct okb_ctx {
  int device = 0;
  int max_windows = 0;
  int sm_count = 148;
  int smem_optin = 0;
  int smem_per_sm = 0;
  int chunk_cap = 1;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_imu = nullptr;     // k_imu runs beside the landmark kernels
  cudaEvent_t ev_round = nullptr, ev_imu = nullptr;
  // Transfers (window uploads / estimate downloads) run on their own stream so that they overlap solver
  // kernels of other windows.  ev_join orders the solver stream after all uploads issued so far;
  // done_ring[k] is recorded on the solver stream after each optimize / reset and transfers of the
  // windows it touched wait on it.
  cudaStream_t stream_xfer = nullptr;
  cudaEvent_t ev_join = nullptr;
  static constexpr int kDoneRing = 8;
  cudaEvent_t done_ring[kDoneRing] = {};
  int done_next = 0;
  std::vector<WinStore> wins;
  std::vector<okb::WinDev> host;      // host mirror of d_wins
  okb::WinDev* d_wins = nullptr;
  okb::SolverState* d_states = nullptr;
  okb::SolverState* h_states = nullptr;  // pinned
  void* hook_buf = nullptr;
  size_t hook_bytes = 0;
  okb_frontend_state* frontend = nullptr;
  // landmark-sharded windows (okb_shard_*): this rank's mailbox and the peers' mailboxes as mapped on this device
  int shard_rank = 0, shard_world = 1, shard_box_cap = 0;
  size_t shard_win_bytes = 0;
  unsigned char* shard_local = nullptr;                       // cudaMalloc'ed, exported through cudaIpc
  unsigned char* shard_peer[okb::kMaxShard] = {};             // [rank]; own entry = shard_local
  bool shard_peer_ipc[okb::kMaxShard] = {};                   // opened with cudaIpcOpenMemHandle (closed at destroy)
  int64_t launches = 0;
  // captured launch sequences of okb_optimize (run_rounds): key fields up to `exec`, compared bytewise
  struct GraphEntry {
    int first, count, rounds, with_quality, opt_iter, opt_min, opt_cauchy, _pad;
    double opt_time;
    unsigned char plan[96];
    cudaGraphExec_t exec;
    int64_t launches;
    uint64_t stamp;
  };
  std::vector<GraphEntry> graphs;
  uint64_t graph_clock = 0;
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;   // pairs (start, stop)
  std::vector<int> prof_kind;             // kernel id per pair
  size_t prof_used = 0;
  std::string error;
  std::mutex error_mu;                    // uploads of different slots may run on several host threads
  void set_error(const std::string& e) { std::lock_guard<std::mutex> g(error_mu); error = e; }
};

void okb_frontend_release(okb_ctx* c);
