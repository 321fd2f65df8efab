/*
 * okvis_b200.h -- C-ABI of libokvis_b200.so, the B200-native hot path of OKVIS.
 *
 * Every entry point is extern "C", takes plain pointers / sizes / POD structs and
 * returns an int status (0 = OKB_OK, <0 = error, see okb_status).  Host buffers are
 * borrowed for the duration of a call only; device memory is owned by the opaque
 * context handle.  No exception ever crosses this boundary.
 *
 * Reference interfaces replaced (paths relative to the OKVIS source tree):
 *   okb_optimize              <- okvis::Estimator::optimize           okvis_ceres/src/Estimator.cpp:843-906
 *                                (Map::solve -> ceres::Solve           okvis_ceres/include/okvis/ceres/Map.hpp:371-373)
 *   okb_window_upload/download<- the graph held by okvis::ceres::Map   okvis_ceres/include/okvis/ceres/Map.hpp:348-402
 *   okb_eval_reprojection     <- ReprojectionError::EvaluateWithMinimalJacobians
 *                                okvis_ceres/include/okvis/ceres/implementation/ReprojectionError.hpp:87-242
 *   okb_eval_imu              <- ImuError::EvaluateWithMinimalJacobians okvis_ceres/src/ImuError.cpp:514-685
 *   okb_imu_propagate         <- ImuError::propagation (static)        okvis_ceres/src/ImuError.cpp:287-504
 *   okb_eval_pose_error       <- PoseError::EvaluateWithMinimalJacobians         okvis_ceres/src/PoseError.cpp:86-136
 *   okb_eval_speed_bias_error <- SpeedAndBiasError::EvaluateWithMinimalJacobians okvis_ceres/src/SpeedAndBiasError.cpp:89-116
 *   okb_eval_relative_pose    <- RelativePoseError::EvaluateWithMinimalJacobians okvis_ceres/src/RelativePoseError.cpp:84-162
 *   okb_eval_marginalization  <- MarginalizationError::EvaluateWithMinimalJacobians okvis_ceres/src/MarginalizationError.cpp:893-946
 *   okb_window_marginalize    <- MarginalizationError::addResidualBlock / marginalizeOut / updateErrorComputation
 *                                okvis_ceres/src/MarginalizationError.cpp:127-435, 507-846 (driven by Estimator.cpp:434-773)
 *   okb_hamming_match         <- okvis::DenseMatcher::match            okvis_matcher/include/okvis/implementation/DenseMatcher.hpp:48-225
 *                                + assignbest                          okvis_matcher/src/DenseMatcher.cpp:69-110
 *   okb_hamming_match_gated   <- VioKeyframeWindowMatchingAlgorithm::distance + verifyMatch (hpp:132-144, cpp:304-339),
 *                                ProbabilisticStereoTriangulator::stereoTriangulate (cpp:168-227), triangulateFast
 *   okb_hamming_candidates    <- VioKeyframeWindowMatchingAlgorithm::specificDescriptorDistance
 *                                okvis_frontend/include/okvis/VioKeyframeWindowMatchingAlgorithm.hpp:246-254
 *   okb_detect_describe       <- okvis::Frontend::detectAndDescribe    okvis_frontend/src/Frontend.cpp:92-114
 *                                (Frame::detect / Frame::describe      okvis_cv/include/okvis/implementation/Frame.hpp:109-156)
 */
#ifndef OKVIS_B200_H_
#define OKVIS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
typedef enum {
  OKB_OK = 0,
  OKB_ERR_INVALID_ARG = -1,
  OKB_ERR_CUDA = -2,          /* a CUDA runtime call failed; see okb_last_error */
  OKB_ERR_UNSUPPORTED = -3,   /* e.g. non-fixed extrinsics in the device solver */
  OKB_ERR_CAPACITY = -4,      /* window larger than the compiled-in limits */
  OKB_ERR_NO_DEVICE = -5,
  OKB_ERR_NUMERIC = -6        /* solver failure (all windows report in summary) */
} okb_status;

/* --------------------------------------------------------------- basic PODs */

/* Distortion model ids follow okvis_cv/include/okvis/cameras/ (the four Distortion headers) */
enum { OKB_DIST_NONE = 0, OKB_DIST_RADTAN = 1, OKB_DIST_EQUIDISTANT = 2, OKB_DIST_RADTAN8 = 3 };

/* PinholeCamera<DISTORTION>: fu,fv,cu,cv + distortion coefficients
 * (okvis_cv/include/okvis/cameras/PinholeCamera.hpp:60-110). */
typedef struct okb_camera {
  int32_t model;
  int32_t width, height;
  int32_t _pad;
  double fu, fv, cu, cv;
  double dist[8];
} okb_camera;

/* One ReprojectionError residual block (okvis_ceres/include/okvis/implementation/Estimator.hpp:43-90):
 * parameter blocks (pose[pose_idx], landmark[lm_idx], extrinsics[ext_idx]); measurement z (pixels);
 * information = sqrt_info^2 * I2 (sqrt_info = 8/keypoint size); CauchyLoss(1) always attached. */
typedef struct okb_observation {
  uint32_t pose_idx, lm_idx, ext_idx, cam_idx;
  double z[2];
  double sqrt_info;
} okb_observation;

/* okvis::ImuMeasurement with the time stamp as int64 nanoseconds
 * (okvis_common/include/okvis/Measurements.hpp; Time::toSec = sec + 1e-9 nsec). */
typedef struct okb_imu_sample {
  int64_t t_ns;
  double gyro[3];
  double acc[3];
} okb_imu_sample;

/* okvis::ImuParameters (okvis_common/include/okvis/Parameters.hpp:105-120), numeric part. */
typedef struct okb_imu_params {
  double a_max, g_max;
  double sigma_g_c, sigma_a_c;
  double sigma_bg, sigma_ba;
  double sigma_gw_c, sigma_aw_c;
  double tau, g;
  double a0[3];
  int32_t rate;
  int32_t _pad;
} okb_imu_params;

/* One ImuError residual block (okvis_ceres/src/Estimator.cpp:288-306): parameter blocks
 * (pose0, sb0, pose1, sb1); samples [sample_offset, sample_offset+sample_count) of the window's
 * sample array cover [t0,t1] including one sample before t0 and one after t1. */
typedef struct okb_imu_term {
  uint32_t pose0, sb0, pose1, sb1;
  int64_t t0_ns, t1_ns;
  uint32_t sample_offset, sample_count;
} okb_imu_term;

/* PoseError on a T_WS block (okvis_ceres/src/PoseError.cpp:48-136).  sqrt_info is the 6x6
 * (row-major) matrix the reference obtains as LLT(information).matrixL().transpose(); the caller
 * supplies it explicitly because the reference's first-frame prior is singular. */
typedef struct okb_pose_prior {
  uint32_t pose_idx;
  uint32_t _pad;
  double meas[7];          /* [t, q_xyzw] */
  double sqrt_info[36];
} okb_pose_prior;

/* SpeedAndBiasError (okvis_ceres/src/SpeedAndBiasError.cpp:46-116). */
typedef struct okb_sb_prior {
  uint32_t sb_idx;
  uint32_t _pad;
  double meas[9];
  double sqrt_info[81];
} okb_sb_prior;

/* RelativePoseError between two extrinsics blocks (okvis_ceres/src/RelativePoseError.cpp:47-162). */
typedef struct okb_relpose_term {
  uint32_t ext0, ext1;
  double sqrt_info[36];
} okb_relpose_term;

/* Block kinds referenced by the marginalisation prior. */
enum { OKB_BLOCK_POSE = 0, OKB_BLOCK_SPEED_BIAS = 1, OKB_BLOCK_EXTRINSICS = 2 };

/* The linearised prior produced by MarginalizationError::updateErrorComputation
 * (okvis_ceres/src/MarginalizationError.cpp:806-846): e = e0 + J * DeltaChi, J is n x n row-major,
 * DeltaChi stacks minus(x0_i, x_i) of the connected blocks in list order (fixed blocks contribute
 * no columns).  x0 holds the linearisation points, 7 / 9 / 7 doubles per block, concatenated. */
typedef struct okb_marg_prior {
  int32_t n;               /* residual dimension = sum of minimal dims */
  int32_t n_blocks;
  const int32_t* block_kind;   /* [n_blocks] OKB_BLOCK_* */
  const uint32_t* block_idx;   /* [n_blocks] index into poses / speed_bias / extrinsics */
  const double* x0;            /* concatenated linearisation points */
  const double* J;             /* [n][n] row-major */
  const double* e0;            /* [n] */
} okb_marg_prior;

/* A whole keyframe window = what okvis::ceres::Map holds when optimize() is called. */
typedef struct okb_window_desc {
  int32_t n_poses, n_speed_bias, n_extrinsics, n_landmarks;
  int32_t n_cameras, n_obs, n_imu_terms, n_imu_samples;
  int32_t n_pose_priors, n_sb_priors, n_relpose_terms, _pad;
  const double* poses;            /* [n_poses][7]       [t, q_xyzw] (PoseParameterBlock.cpp:68-79) */
  const double* speed_bias;       /* [n_speed_bias][9]  [v_W, b_g, b_a] */
  const double* extrinsics;       /* [n_extrinsics][7]  T_SC */
  const uint8_t* extrinsics_fixed;/* [n_extrinsics] 1 = setParameterBlockConstant (Estimator.cpp:264-267) */
  const double* landmarks;        /* [n_landmarks][4]   homogeneous, w untouched by updates */
  const okb_camera* cameras;      /* [n_cameras] */
  const okb_observation* obs;     /* [n_obs] */
  const okb_imu_term* imu_terms;  /* [n_imu_terms] */
  const okb_imu_sample* imu_samples; /* [n_imu_samples] */
  okb_imu_params imu_params;
  const okb_pose_prior* pose_priors;
  const okb_sb_prior* sb_priors;
  const okb_relpose_term* relpose_terms;
  const okb_marg_prior* marg;     /* NULL if no marginalisation prior */
} okb_window_desc;

/* Solver options = the Ceres 1.9 options Estimator::optimize sets (Estimator.cpp:854-874)
 * plus the time-limit callback (Estimator.cpp:909-929, CeresIterationCallback.hpp:78-87). */
typedef struct okb_solve_options {
  int32_t max_iterations;      /* numIter */
  int32_t min_iterations;      /* CeresIterationCallback::iterationMinimum_ */
  double time_limit_s;         /* <0: no limit */
  int32_t use_cauchy_loss;     /* 1 = CauchyLoss(1) on reprojection blocks (reference default) */
  int32_t _pad;
} okb_solve_options;

/* Per-window result summary (subset of ceres::Solver::Summary). */
enum { OKB_TERM_NO_CONVERGENCE = 0, OKB_TERM_FUNCTION_TOL = 1, OKB_TERM_PARAMETER_TOL = 2,
       OKB_TERM_GRADIENT_TOL = 3, OKB_TERM_MIN_RADIUS = 4, OKB_TERM_TIME_LIMIT = 5,
       OKB_TERM_FAILURE = 6 };
typedef struct okb_summary {
  double initial_cost, final_cost;
  int32_t iterations;          /* accepted + rejected steps, as Ceres counts them */
  int32_t num_successful_steps;
  int32_t termination;
  int32_t imu_redo_count;
  double final_radius;
  double solve_time_s;         /* device time for this optimize call (whole batch) */
} okb_summary;

/* ------------------------------------------------------------ estimator path */
typedef struct okb_ctx okb_ctx;

/* Creates a context on CUDA device `device_id` able to hold `max_windows` resident windows. */
int okb_ctx_create(int device_id, int max_windows, okb_ctx** out);
void okb_ctx_destroy(okb_ctx* ctx);
const char* okb_last_error(const okb_ctx* ctx);   /* ctx may be NULL: returns the creation error */
/* Number of kernels this library has launched on this context since creation. */
int64_t okb_kernel_launches(const okb_ctx* ctx);
/* Raw CUDA stream (cudaStream_t) the context launches on, for event timing by the caller. */
void* okb_stream(const okb_ctx* ctx);

/* Packs and uploads one window into slot `win` (0 <= win < max_windows).  The library stores landmarks internally
 * sorted by observing-frame range; indices and results at this boundary always use the caller's order.  A slot must
 * not be uploaded while an okb_optimize_async on it is still in flight (call okb_optimize_finish first). */
int okb_window_upload(okb_ctx* ctx, int win, const okb_window_desc* desc);
/* Batched form of okb_window_upload (no reference counterpart: one okvis::Estimator owns one window; this serves
 * callers that keep many estimators / sessions on one GPU).  Uploads `count` windows into slots
 * [win_first, win_first+count); descs[i] describes slot win_first+i.
 * The packing of different slots is spread over up to `host_threads` host threads (<= 0: library default).
 * Uploads run on the context's transfer stream: they overlap solver work already launched on OTHER
 * slots, and every later okb_optimize* / okb_window_reset is ordered after them.  Uploads of different slots
 * may also be issued concurrently from several caller threads. */
int okb_window_upload_batch(okb_ctx* ctx, int win_first, int count, const okb_window_desc* descs, int host_threads);
/* ---- resident window: incremental graph updates -------------------------------------------------------------
 * Replace the reference's graph bookkeeping between two optimize() calls -- Estimator::addStates
 * (okvis_ceres/src/Estimator.cpp:110-343), addLandmark (:346-368), addObservation
 * (include/okvis/implementation/Estimator.hpp:43-90), removeObservation (Estimator.cpp:371-413), set_T_WS /
 * setSpeedAndBias / setLandmark (include/okvis/Estimator.hpp:343-409) and the frame / landmark removal part of
 * applyMarginalizationStrategy (Estimator.cpp:434-773) -- so that a window stays resident on the device and only
 * one frame's worth of data crosses PCIe per optimize.  Every call appends a command to the slot's pending stream
 * on the host (no device work); okb_window_commit -- or the next okb_optimize* / okb_window_download* /
 * okb_window_reset of the slot -- ships the stream with one copy and the device applies the commands in order and
 * re-compiles its internal layout.  Index conventions: pose / speed-bias indices are positions in the window
 * (removing position p moves every later frame down by one, exactly like erasing from the reference's ordered
 * statesMap_); landmark indices are stable slots chosen by the caller (0 <= idx < max_landmarks).
 * Indices are validated on the host against the mirrored dimensions (OKB_ERR_INVALID_ARG / OKB_ERR_CAPACITY
 * synchronously); duplicate observations and inconsistent extrinsics are detected by the device and reported by
 * the next okb_optimize_finish / okb_optimize. */

/* Capacities of slot `win` for later incremental growth; call before okb_window_upload of that slot (an upload
 * larger than the reservation simply enlarges it).  max_marg_dim <= 160. */
int okb_window_reserve(okb_ctx* ctx, int win, int max_frames, int max_landmarks, int max_observations,
                       int max_imu_samples, int max_marg_dim);
/* Appends one frame: pose (index n_poses), optionally its speed/bias block (index n_speed_bias; NULL: none) and the
 * ImuError term that links it (Estimator.cpp:288-306; NULL: none).  term->pose0/sb0/pose1/sb1 index the window
 * AFTER the append; term->sample_offset is relative to `samples`. */
int okb_window_add_frame(okb_ctx* ctx, int win, const double* pose /*7*/, const double* speed_bias /*9 or NULL*/,
                         const okb_imu_term* term, const okb_imu_sample* samples, int n_samples);
/* Removes a frame: its pose block, speed/bias block sb_idx (0xffffffff: none), every observation made in it, the IMU
 * terms and priors attached to it.  Later frames move down one position everywhere (observations, IMU terms,
 * priors, marginalisation-prior block list; a prior that still references the removed frame is an error). */
int okb_window_remove_frame(okb_ctx* ctx, int win, uint32_t pose_idx, uint32_t sb_idx);
/* Creates or overwrites landmark slots (Estimator::addLandmark / setLandmark). */
int okb_window_set_landmarks(okb_ctx* ctx, int win, int n, const uint32_t* idx, const double* xyzw /*[n][4]*/);
/* Removes landmarks and all their observations (Estimator::removeObservation loop of Estimator.cpp:625-725). */
int okb_window_remove_landmarks(okb_ctx* ctx, int win, int n, const uint32_t* idx);
int okb_window_add_observations(okb_ctx* ctx, int win, int n, const okb_observation* obs);
typedef struct okb_obs_key { uint32_t pose_idx, lm_idx, cam_idx, _pad; } okb_obs_key;
/* Estimator::removeObservation(landmarkId, poseId, camIdx, keypointIdx) (Estimator.cpp:371-413); unknown keys are ignored. */
int okb_window_remove_observations(okb_ctx* ctx, int win, int n, const okb_obs_key* keys);
/* Estimator::set_T_WS / setSpeedAndBias. */
int okb_window_set_states(okb_ctx* ctx, int win, int n_poses, const uint32_t* pose_idx, const double* poses /*[n][7]*/,
                          int n_sb, const uint32_t* sb_idx, const double* speed_bias /*[n][9]*/);
/* Replaces the pose priors and speed/bias priors; marg != NULL also replaces the marginalisation prior
 * (marg->n == 0 removes it), marg == NULL keeps the current one. */
int okb_window_set_priors(okb_ctx* ctx, int win, int n_pose_priors, const okb_pose_prior* pose_priors, int n_sb_priors,
                          const okb_sb_prior* sb_priors, const okb_marg_prior* marg);
/* Ships the pending commands of slots [win_first, win_first+count) to the device (transfer stream; overlaps solver
 * work on other slots).  Optional: the calls that need the result do it themselves. */
int okb_window_commit(okb_ctx* ctx, int win_first, int count);

/* Removes speed/bias block sb_idx together with the ImuError terms and SpeedAndBiasError priors attached to it; the
 * pose of that frame stays (the "removeAllButPose" frames of Estimator::applyMarginalizationStrategy,
 * Estimator.cpp:483-554).  Later speed/bias blocks move down one position. */
int okb_window_remove_speed_bias(okb_ctx* ctx, int win, uint32_t sb_idx);

/* ---- device-side marginalisation -----------------------------------------------------------------------------
 * The numeric half of Estimator::applyMarginalizationStrategy (okvis_ceres/src/Estimator.cpp:434-773):
 * MarginalizationError::addResidualBlock (okvis_ceres/src/MarginalizationError.cpp:127-435) for the listed residual
 * blocks -- evaluated at the first-estimate linearisation points, reprojection errors with the Cauchy correction --
 * marginalizeOut (:507-802) of the marked blocks and of the listed landmarks, and updateErrorComputation (:806-846).
 * The window's marginalisation prior is replaced on the device by the result (kept blocks in job order; their
 * linearisation points are kept when block_prev >= 0, else the current estimates become x0).  The bookkeeping half
 * stays with the caller, exactly like the reference: after this call it removes the linearised terms / blocks with
 * okb_window_remove_frame / okb_window_remove_speed_bias / okb_window_remove_landmarks and re-fixes the first pose
 * with okb_window_set_priors (Estimator.cpp:761-770).  H and b0 persist on the device between calls as the
 * reference's H_ / b0_ members do (a prior supplied by the host through okb_window_upload / okb_window_set_priors
 * starts from H = J^T J, b0 = -J^T e0).  Errors found by the device (a residual whose parameter blocks are not all
 * listed, inconsistent block_prev) surface at the next okb_optimize* / okb_window_download_marg. */
typedef struct okb_marg_job {
  int32_t n_blocks;                 /* dense parameter blocks of the linear system, in H order (<= 64) */
  int32_t n_imu_terms, n_sb_priors, n_landmarks;
  const int32_t* block_kind;        /* [n_blocks] OKB_BLOCK_POSE / OKB_BLOCK_SPEED_BIAS */
  const uint32_t* block_idx;        /* [n_blocks] position in the window NOW */
  const int32_t* block_prev;        /* [n_blocks] position in the current prior's block list, -1 = newly connected; every
                                       block of the current prior must appear exactly once */
  const uint8_t* block_marginalize; /* [n_blocks] 1 = marginalise this block out */
  const uint32_t* imu_terms;        /* ImuError terms to linearise (positions in the window's term list) */
  const uint32_t* sb_priors;        /* SpeedAndBiasError terms to linearise */
  const uint32_t* landmarks;        /* landmarks to marginalise: every observation they still have is linearised */
} okb_marg_job;
int okb_window_marginalize(okb_ctx* ctx, int win, const okb_marg_job* job);
/* Reads the window's current marginalisation prior back (tests, host-side persistence): J [n][n], e0 [n], the
 * H / b0 it was factored from, block list, linearisation points.  Any pointer may be NULL; arrays must hold the
 * compiled-in maxima (n <= 160, 64 blocks).  status[4] = {device error code of the last okb_window_marginalize,
 * rank of H, rank of the dense marginalised block, Jacobi sweeps}. */
int okb_window_download_marg(okb_ctx* ctx, int win, int32_t* n, int32_t* n_blocks, int32_t* block_kind, uint32_t* block_idx,
                             double* x0, double* J, double* e0, double* H, double* b0, int32_t* status);

/* Bytes the last commit (full upload or incremental commands) of this slot copied host -> device (0 if the slot is empty). */
int64_t okb_window_h2d_bytes(const okb_ctx* ctx, int win);
/* Restores the state of the last FULL upload of the slot on the device (no host traffic); used to repeat a solve
 * on resident data.  The window must have the shape (frames, landmarks) it had at that upload. */
int okb_window_reset(okb_ctx* ctx, int win_first, int win_count);
/* Runs the dogleg/Schur solver on windows [win_first, win_first+win_count) in one batch.
 * summaries may be NULL; otherwise [win_count]. */
int okb_optimize(okb_ctx* ctx, int win_first, int win_count, const okb_solve_options* opt,
                 okb_summary* summaries);
/* Same as okb_optimize but does not synchronise or read anything back (for device-side timing);
 * call okb_optimize_finish to collect the summaries. */
int okb_optimize_async(okb_ctx* ctx, int win_first, int win_count, const okb_solve_options* opt);
int okb_optimize_finish(okb_ctx* ctx, int win_first, int win_count, okb_summary* summaries);
/* Copies the estimates back.  Any pointer may be NULL.  quality[l] = sqrt(lambda_min)/sqrt(lambda_max)
 * of the landmark's 3x3 Hessian block as in Estimator.cpp:880-894 (0 if lambda_min < 1e-12). */
int okb_window_download(okb_ctx* ctx, int win, double* poses, double* speed_bias,
                        double* landmarks, double* quality);
/* Batched form of okb_window_download (no reference counterpart).  Downloads the estimates of `count` slots with one
 * synchronisation.  Each argument is an array of `count`
 * host pointers (or NULL to skip that quantity; individual entries may be NULL too).  Like the uploads this
 * runs on the transfer stream: it waits only for solver work launched on these slots, not for other slots. */
int okb_window_download_batch(okb_ctx* ctx, int win_first, int count, double* const* poses,
                              double* const* speed_bias, double* const* landmarks, double* const* quality);

/* ----------------------------------------------------- landmark-sharded single window (multi-GPU)
 * No reference counterpart (the reference is single-process CPU code); this is the partition SURVEY.md 8(e) row 2 /
 * BASELINE.json configs[4] name: the landmarks (and their observations) of ONE large window are dealt to `world`
 * ranks (lm_idx mod world), every rank holds all poses / speed-bias / IMU / prior blocks.  Per solver round each
 * rank builds the Schur complement of its shard, the partial reduced systems are all-reduced over NVLink peer
 * memory (push into every rank's mailbox + flags, fused into the chunk-reduction kernel and the reduced-solve
 * kernel's prologue), every rank solves the reduced system redundantly and back-substitutes its own landmarks.
 * A context is made a shard member BEFORE its windows are uploaded; every rank then uploads its shard into the
 * same slot and all ranks issue the same okb_optimize* calls.  time_limit_s uses rank 0's clock.
 *   multi-process (one process per GPU): okb_shard_export on every rank, exchange the OKB_SHARD_HANDLE_BYTES-byte
 *     handles out of band (e.g. torch.distributed.all_gather), okb_shard_connect with all `world` handles in rank order.
 *   single process driving several contexts (same or different devices): okb_shard_connect_local.
 * max_frames bounds the keyframes of any sharded window (mailbox sizing). */
#define OKB_SHARD_HANDLE_BYTES 64
int okb_shard_export(okb_ctx* ctx, int rank, int world, int max_frames, void* handle_out /* OKB_SHARD_HANDLE_BYTES */);
int okb_shard_connect(okb_ctx* ctx, const void* handles /* [world][OKB_SHARD_HANDLE_BYTES] */);
int okb_shard_connect_local(okb_ctx* const* ctxs /* [world], rank order */, int world, int max_frames);
/* After okb_optimize_finish: out[0] = exchange rounds of the last optimize of `win`, out[1] = microseconds this rank
 * spent waiting for and summing the peers' partial systems (device clock), out[2] = 1 if a peer timed out,
 * out[3] = exchange epoch. */
int okb_shard_stats(okb_ctx* ctx, int win, double out[4]);

/* Optional per-kernel device timing: CUDA events on the context stream around every solver kernel
 * launch.  okb_profile_enable(ctx,1) clears the counters; okb_profile_read synchronises and returns
 * out[0] = ms in k_landmarks, out[1] = its launches, out[2] = ms in k_solve, out[3] = launches,
 * out[4] = ms in k_quality, out[5] = launches. */
int okb_profile_enable(okb_ctx* ctx, int on);
int okb_profile_read(okb_ctx* ctx, double out[6]);
/* Diagnostics: nanoseconds the reduced-solve kernel spent per internal phase during the last optimize of
 * `win` (dense terms, partial gather, assembly, Cholesky, substitution, back-substitution, dogleg). */
int okb_debug_phase_ns(okb_ctx* ctx, int win, double out[16]);
/* Diagnostics: the mutable preintegration cache of ImuError term `term` (ImuError.hpp:251-276): the bias the
 * preintegration was done at, whether it is valid (0 = redo_), and how often it was redone. */
int okb_debug_imu_cache(okb_ctx* ctx, int win, int term, double sb_ref[9], int32_t* valid, int32_t* redo_count);

/* ------------------------------------------------- single-block test hooks
 * Mirror ErrorInterface::EvaluateWithMinimalJacobians (okvis_ceres/include/okvis/ceres/ErrorInterface.hpp:93-95).
 * All run on the device through the same device functions the solver uses.  Jacobians are the
 * MINIMAL ones, row-major.  n = batch size; arrays are [n][...]. */
int okb_eval_reprojection(okb_ctx* ctx, int n, const okb_camera* cam, const double* pose /*[n][7]*/,
                          const double* landmark /*[n][4]*/, const double* extrinsics /*[n][7]*/,
                          const double* z /*[n][2]*/, const double* sqrt_info /*[n]*/,
                          double* r /*[n][2]*/, double* J_pose /*[n][12]*/, double* J_lm /*[n][6]*/,
                          double* J_ext /*[n][12]*/);
int okb_eval_imu(okb_ctx* ctx, const okb_imu_params* params, const okb_imu_sample* samples, int n_samples,
                 int64_t t0_ns, int64_t t1_ns, const double* pose0, const double* sb0,
                 const double* pose1, const double* sb1, const double* sb_ref /* linearisation point or NULL */,
                 double* r /*15*/, double* J0 /*15x6*/, double* J1 /*15x9*/, double* J2 /*15x6*/,
                 double* J3 /*15x9*/, double* sqrt_info_out /*15x15 or NULL*/);
int okb_imu_propagate(okb_ctx* ctx, const okb_imu_params* params, const okb_imu_sample* samples, int n_samples,
                      int64_t t0_ns, int64_t t1_ns, double* pose /*7 in/out*/, double* sb /*9 in/out*/,
                      double* covariance /*15x15 or NULL*/, double* jacobian /*15x15 or NULL*/, int* n_used);
int okb_eval_pose_error(okb_ctx* ctx, const double* meas /*7*/, const double* sqrt_info /*36*/,
                        const double* pose /*7*/, double* r /*6*/, double* J /*6x6*/);
int okb_eval_speed_bias_error(okb_ctx* ctx, const double* meas /*9*/, const double* sqrt_info /*81*/,
                              const double* sb /*9*/, double* r /*9*/, double* J /*9x9*/);
int okb_eval_relative_pose(okb_ctx* ctx, const double* sqrt_info /*36*/, const double* pose0, const double* pose1,
                           double* r /*6*/, double* J0 /*6x6*/, double* J1 /*6x6*/);
/* x: current block values concatenated like marg->x0; J_eff is the n x n effective minimal Jacobian
 * after Ceres multiplies by the local parameterisation at x (J * lift(x0) * plus(x)). */
int okb_eval_marginalization(okb_ctx* ctx, const okb_marg_prior* marg, const double* x,
                             double* r /*n*/, double* J_eff /*n x n*/);

/* -------------------------------------------------------------- frontend path */
typedef struct okb_keypoint {      /* cv::KeyPoint fields the reference uses */
  float x, y;
  float size;
  float angle;                     /* degrees, set from the gravity direction (Frame.hpp(impl):128-156) */
  float response;
  int32_t octave;
} okb_keypoint;

/* brisk::ScaleSpaceFeatureDetector<HarrisScoreCalculator>(uniformityRadius, octaves=0,
 * absoluteThreshold, maxNumKpt) as constructed at okvis_frontend/src/Frontend.cpp:828-831. */
typedef struct okb_detect_params {
  double uniformity_radius;        /* "detection threshold" in the yaml (40) */
  double absolute_threshold;       /* 800 */
  int32_t max_keypoints;           /* 400 */
  int32_t desc_bytes;              /* 48 (reference) or 64 */
  int32_t rotation_invariance;     /* 1 */
  int32_t _pad;
} okb_detect_params;

int okb_detect_describe(okb_ctx* ctx, int cam_slot, const uint8_t* img, int width, int height, int stride,
                        const okb_camera* cam, const double R_CW[9], const okb_detect_params* params,
                        okb_keypoint* out_kp, uint8_t* out_desc, int max_out, int* n_out);

typedef struct okb_pair {          /* DenseMatcher::Pairing */
  int32_t index_a;                 /* -1 = B element unmatched */
  float distance;
} okb_pair;

/* DenseMatcher::match with the plain Hamming MatchingAlgorithm: per-A top-`num_best` lists
 * (out_topk [nA][num_best], may be NULL), greedy mutual assignment in sequential A-ascending order,
 * per-B winners (out_pairs [nB]).  use_ratio / ratio_threshold mirror useDistanceRatioThreshold_.
 * nA == 0 or nB == 0 is legal (no matches: every out_pairs entry is {-1, threshold}), like DenseMatcher::match on an
 * empty MatchingAlgorithm; the same holds for the gated and the candidate-list calls. */
int okb_hamming_match(okb_ctx* ctx, const uint8_t* A, int nA, const uint8_t* B, int nB, int desc_bytes,
                      const uint8_t* skipA, const uint8_t* skipB, float threshold, int num_best,
                      int use_ratio, float ratio_threshold, okb_pair* out_topk, okb_pair* out_pairs);

/* ---- geometry-gated matching --------------------------------------------------------------
 * VioKeyframeWindowMatchingAlgorithm::distance (okvis_frontend/include/okvis/VioKeyframeWindowMatchingAlgorithm.hpp:132-144)
 * returns the Hamming distance only if it is below the threshold AND verifyMatch passes
 * (okvis_frontend/src/VioKeyframeWindowMatchingAlgorithm.cpp:304-339), else FLT_MAX; the top-k lists and assignbest of
 * DenseMatcher then run over these gated distances.  okb_hamming_match_gated evaluates the gate on the device inside
 * the list kernel, so no candidate list has to travel back to the host.
 *   OKB_GATE_3D2D : chi-square test of landmark A's projection into frame B against keypoint B (:317-338);
 *                   proj_into_b / proj_uncertainty are what doSetup computes (:159-208)
 *   OKB_GATE_2D2D : ProbabilisticStereoTriangulator::stereoTriangulate (src/ProbabilisticStereoTriangulator.cpp:168-227):
 *                   triangulateFast (src/stereo_triangulation.cpp:51-125) of the two back-projected rays with
 *                   sigma = max(raySigmasA[a], raySigmasB[b]) and the two reprojection checks (:359-384).  The
 *                   bearing vectors are the cameras' backProject() of the keypoints (the caller has the camera
 *                   objects; the reference recomputes them inside every call). */
enum { OKB_GATE_NONE = 0, OKB_GATE_3D2D = 1, OKB_GATE_2D2D = 2 };
typedef struct okb_match_gate {
  int32_t mode, _pad;
  const double* kp_b;              /* [nB][2] keypoint coordinates in B (both modes) */
  const double* kp_size_b;         /* [nB]    keypoint size (both modes) */
  const double* proj_into_b;       /* 3D-2D: [nA][2] projectionsIntoB_ */
  const double* proj_uncertainty;  /* 3D-2D: [nA][4] projectionsIntoBUncertainties_, 2x2 row-major */
  const double* kp_a;              /* 2D-2D: [nA][2] */
  const double* kp_size_a;         /* 2D-2D: [nA] */
  const double* bearing_a;         /* 2D-2D: [nA][3] backProject(kp_a), camera A frame, any length */
  const double* bearing_b;         /* 2D-2D: [nB][3] backProject(kp_b), camera B frame */
  const double* ray_sigma_a;       /* 2D-2D: [nA] raySigmasA_ */
  const double* ray_sigma_b;       /* 2D-2D: [nB] raySigmasB_ */
  okb_camera cam_a, cam_b;         /* 2D-2D: geometries for the reprojection checks */
  double T_AB[7];                  /* 2D-2D: pose of camera B in camera A [t, q_xyzw] */
} okb_match_gate;
int okb_hamming_match_gated(okb_ctx* ctx, const uint8_t* A, int nA, const uint8_t* B, int nB, int desc_bytes,
                            const uint8_t* skipA, const uint8_t* skipB, float threshold, int num_best, int use_ratio,
                            float ratio_threshold, const okb_match_gate* gate, okb_pair* out_topk, okb_pair* out_pairs);

/* Candidate-list mode for the production (geometry-gated) matching algorithm: for every A, every B
 * with Hamming distance < threshold in ascending B order, CSR layout.  row_ptr [nA+1];
 * col_idx / dist have capacity `cap`; returns OKB_ERR_CAPACITY if more candidates exist
 * (row_ptr[nA] then holds the required capacity). */
int okb_hamming_candidates(okb_ctx* ctx, const uint8_t* A, int nA, const uint8_t* B, int nB, int desc_bytes,
                           float threshold, uint32_t* row_ptr, uint32_t* col_idx, uint16_t* dist, int cap);

#ifdef __cplusplus
}
#endif
#endif  /* OKVIS_B200_H_ */
