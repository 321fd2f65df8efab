"""Thin Python binding of the C-ABI (include/okvis_b200.h) -- the same calls a C++ host (the OKVIS
Estimator/Frontend shims) makes.  No numerics live here."""
import ctypes as C

import numpy as np

from . import abi

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = abi.load_library()
    return _lib


class OkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("okb error %d: %s" % (code, msg))
        self.code = code


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """okb_ctx: device memory for `max_windows` resident keyframe windows + the frontend buffers."""

    def __init__(self, device=0, max_windows=1):
        self._h = C.c_void_p()
        rc = lib().okb_ctx_create(int(device), int(max_windows), C.byref(self._h))
        if rc != 0:
            raise OkbError(rc, lib().okb_last_error(None).decode())
        self.max_windows = max_windows
        self._windows = {}

    def close(self):
        if self._h:
            lib().okb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OkbError(rc, lib().okb_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return int(lib().okb_kernel_launches(self._h))

    @property
    def stream(self):
        return lib().okb_stream(self._h)

    def profile_enable(self, on=True):
        self._check(lib().okb_profile_enable(self._h, int(on)))

    def profile_read(self):
        out = np.zeros(6)
        self._check(lib().okb_profile_read(self._h, _p(out)))
        return dict(landmarks_ms=out[0], landmarks_launches=int(out[1]), solve_ms=out[2], solve_launches=int(out[3]),
                    quality_ms=out[4], quality_launches=int(out[5]))

    # ---------------------------------------------------------------- landmark-sharded window (multi-GPU)
    def shard_export(self, rank, world, max_frames=32):
        """okb_shard_export: allocates this rank's mailbox; returns the 64-byte IPC handle (numpy uint8)."""
        h = np.zeros(64, np.uint8)
        self._check(lib().okb_shard_export(self._h, int(rank), int(world), int(max_frames), _p(h)))
        return h

    def shard_connect(self, handles):
        h = np.ascontiguousarray(handles, dtype=np.uint8)
        self._check(lib().okb_shard_connect(self._h, _p(h)))

    @staticmethod
    def shard_connect_local(ctxs, max_frames=32):
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        rc = lib().okb_shard_connect_local(arr, len(ctxs), int(max_frames))
        if rc != 0:
            raise OkbError(rc, "; ".join(lib().okb_last_error(c._h).decode() for c in ctxs))

    def shard_stats(self, win=0):
        out = np.zeros(4)
        self._check(lib().okb_shard_stats(self._h, int(win), _p(out)))
        return dict(rounds=int(out[0]), wait_us=float(out[1]), fault=int(out[2]), epoch=int(out[3]))

    # ---------------------------------------------------------------- estimator path
    def upload(self, win, window):
        d = window.desc()
        self._check(lib().okb_window_upload(self._h, int(win), C.byref(d)))
        self._windows[win] = window

    # ---- resident window: incremental graph updates (no device work until commit / optimize / download)
    def reserve(self, win, max_frames, max_landmarks, max_observations, max_imu_samples, max_marg_dim=0):
        self._check(lib().okb_window_reserve(self._h, int(win), int(max_frames), int(max_landmarks), int(max_observations),
                                             int(max_imu_samples), int(max_marg_dim)))

    def add_frame(self, win, pose, speed_bias=None, term=None, samples=None):
        pose = _f64(pose)
        sb = _f64(speed_bias) if speed_bias is not None else None
        t = np.ascontiguousarray(term) if term is not None else None
        smp = np.ascontiguousarray(samples) if samples is not None else None
        self._check(lib().okb_window_add_frame(self._h, int(win), _p(pose), _p(sb), _p(t), _p(smp), 0 if smp is None else len(smp)))

    def remove_frame(self, win, pose_idx, sb_idx=None):
        self._check(lib().okb_window_remove_frame(self._h, int(win), C.c_uint32(int(pose_idx)),
                                                  C.c_uint32(0xffffffff if sb_idx is None else int(sb_idx))))

    def set_landmarks(self, win, idx, xyzw):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        x = _f64(xyzw)
        self._check(lib().okb_window_set_landmarks(self._h, int(win), len(idx), _p(idx), _p(x)))

    def remove_landmarks(self, win, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        self._check(lib().okb_window_remove_landmarks(self._h, int(win), len(idx), _p(idx)))

    def add_observations(self, win, obs):
        obs = np.ascontiguousarray(obs, dtype=abi.observation_dtype)
        self._check(lib().okb_window_add_observations(self._h, int(win), len(obs), _p(obs)))

    def remove_observations(self, win, keys):
        k = np.zeros((len(keys), 4), np.uint32)
        k[:, :3] = np.asarray(keys, dtype=np.uint32).reshape(-1, 3)
        self._check(lib().okb_window_remove_observations(self._h, int(win), len(k), _p(k)))

    def set_states(self, win, pose_idx=(), poses=None, sb_idx=(), speed_bias=None):
        pi = np.ascontiguousarray(pose_idx, dtype=np.uint32)
        si = np.ascontiguousarray(sb_idx, dtype=np.uint32)
        self._check(lib().okb_window_set_states(self._h, int(win), len(pi), _p(pi) if len(pi) else None,
                                                _p(_f64(poses)) if len(pi) else None, len(si), _p(si) if len(si) else None,
                                                _p(_f64(speed_bias)) if len(si) else None))

    def set_priors(self, win, pose_priors, sb_priors, marg=None):
        pp = np.ascontiguousarray(pose_priors, dtype=abi.pose_prior_dtype)
        sp = np.ascontiguousarray(sb_priors, dtype=abi.sb_prior_dtype)
        m = None
        if marg is not None:
            m = abi.MargPrior()
            m.n, m.n_blocks = int(marg["J"].shape[0]), len(marg["block_kind"])
            m.block_kind = marg["block_kind"].ctypes.data_as(C.POINTER(C.c_int32))
            m.block_idx = marg["block_idx"].ctypes.data_as(C.POINTER(C.c_uint32))
            m.x0, m.J, m.e0 = abi.dptr(marg["x0"]), abi.dptr(marg["J"]), abi.dptr(marg["e0"])
        self._check(lib().okb_window_set_priors(self._h, int(win), len(pp), _p(pp) if len(pp) else None, len(sp),
                                                _p(sp) if len(sp) else None, C.byref(m) if m is not None else None))

    def remove_speed_bias(self, win, sb_idx):
        self._check(lib().okb_window_remove_speed_bias(self._h, int(win), C.c_uint32(int(sb_idx))))

    def marginalize(self, win, job):
        """okb_window_marginalize; job from abi.make_marg_job."""
        self._check(lib().okb_window_marginalize(self._h, int(win), C.byref(job)))

    def download_marg(self, win):
        n, nb = C.c_int32(0), C.c_int32(0)
        kind, idx = np.zeros(64, np.int32), np.zeros(64, np.uint32)
        x0, J, e0, H, b0 = np.zeros(9 * 64), np.zeros(160 * 160), np.zeros(160), np.zeros(160 * 160), np.zeros(160)
        st = np.zeros(4, np.int32)
        self._check(lib().okb_window_download_marg(self._h, int(win), C.byref(n), C.byref(nb), _p(kind), _p(idx), _p(x0), _p(J),
                                                   _p(e0), _p(H), _p(b0), _p(st)))
        n, nb = n.value, nb.value
        kind, idx = kind[:nb].copy(), idx[:nb].copy()
        xdim = int(sum(9 if k == abi.BLOCK_SPEED_BIAS else 7 for k in kind))
        return dict(n=n, block_kind=kind, block_idx=idx, x0=x0[:xdim].copy(), J=J[:n * n].reshape(n, n).copy(), e0=e0[:n].copy(),
                    H=H[:n * n].reshape(n, n).copy(), b0=b0[:n].copy(), status=st.copy())

    def commit(self, first=0, count=1):
        self._check(lib().okb_window_commit(self._h, int(first), int(count)))

    def prepare_readd_newest(self, win, w):
        """Pre-marshalled arguments of the streaming pattern `drop the newest frame, add it again with its IMU term and
        observations` (what a VIO host does once per camera frame: Estimator::addStates + addObservation of that frame,
        okvis_ceres/src/Estimator.cpp:110-343, implementation/Estimator.hpp:43-90) for slot `win` holding window `w`."""
        K = len(w.poses)
        t = w.imu_terms[K - 2:K - 1].copy()
        lo, n = int(t["sample_offset"][0]), int(t["sample_count"][0])
        t["sample_offset"] = 0
        smp = np.ascontiguousarray(w.imu_samples[lo:lo + n])
        obs = np.ascontiguousarray(w.obs[w.obs["pose_idx"] == K - 1])
        pose, sb = _f64(w.poses[K - 1]).copy(), _f64(w.speed_bias[K - 1]).copy()
        keep = (pose, sb, t, smp, obs)
        return (C.c_int(int(win)), C.c_uint32(K - 1), _p(pose), _p(sb), _p(t), _p(smp), C.c_int(n), C.c_int(len(obs)), _p(obs), keep)

    def readd_newest(self, prepared):
        """Issues remove_frame / add_frame / add_observations for every prepared slot (host-side command appends only)."""
        L_, h = lib(), self._h
        rm, af, ao = L_.okb_window_remove_frame, L_.okb_window_add_frame, L_.okb_window_add_observations
        for win, last, pose, sb, t, smp, n, n_obs, obs, _ in prepared:
            rc = rm(h, win, last, last) or af(h, win, pose, sb, t, smp, n) or ao(h, win, n_obs, obs)
            if rc:
                self._check(rc)

    def debug_phase_us(self, win):
        out = np.zeros(16)
        self._check(lib().okb_debug_phase_ns(self._h, int(win), _p(out)))
        names = ["dense_terms", "gather", "assemble", "cholesky", "substitution", "backsub", "dogleg", "_",
                 "p8", "p9", "p10", "p11", "p12", "p13", "p14", "p15"]   # OKB_SCHUR_PROF / OKB_CHOL_PROF builds: cycles
        return {n: v * 1e-3 for n, v in zip(names, out)}

    def debug_imu_cache(self, win, term):
        ref = np.zeros(9)
        v, r = C.c_int32(0), C.c_int32(0)
        self._check(lib().okb_debug_imu_cache(self._h, int(win), int(term), _p(ref), C.byref(v), C.byref(r)))
        return ref, v.value, r.value

    def h2d_bytes(self, win):
        f = lib().okb_window_h2d_bytes
        f.restype = C.c_int64
        return int(f(self._h, int(win)))

    def reset(self, first=0, count=1):
        self._check(lib().okb_window_reset(self._h, int(first), int(count)))

    @staticmethod
    def _options(max_iterations, min_iterations, time_limit_s, use_cauchy_loss):
        o = abi.SolveOptions()
        o.max_iterations, o.min_iterations, o.time_limit_s, o.use_cauchy_loss = (max_iterations, min_iterations,
                                                                                time_limit_s, use_cauchy_loss)
        return o

    def optimize(self, first=0, count=1, max_iterations=10, min_iterations=0, time_limit_s=-1.0, use_cauchy_loss=1):
        o = self._options(max_iterations, min_iterations, time_limit_s, use_cauchy_loss)
        out = (abi.Summary * count)()
        self._check(lib().okb_optimize(self._h, int(first), int(count), C.byref(o), out))
        return [s.as_dict() for s in out]

    def optimize_async(self, first=0, count=1, max_iterations=10, min_iterations=0, time_limit_s=-1.0,
                       use_cauchy_loss=1):
        o = self._options(max_iterations, min_iterations, time_limit_s, use_cauchy_loss)
        self._check(lib().okb_optimize_async(self._h, int(first), int(count), C.byref(o)))

    def optimize_finish(self, first=0, count=1):
        out = (abi.Summary * count)()
        self._check(lib().okb_optimize_finish(self._h, int(first), int(count), out))
        return [s.as_dict() for s in out]

    def download(self, win, with_quality=True, dims=None):
        """dims = (n_poses, n_speed_bias, n_landmarks) when the window was changed incrementally since its upload."""
        if dims is None:
            w = self._windows[win]
            dims = (len(w.poses), len(w.speed_bias), len(w.landmarks))
        poses, sb, lms = np.zeros((dims[0], 7)), np.zeros((dims[1], 9)), np.zeros((dims[2], 4))
        q = np.zeros(dims[2]) if with_quality else None
        self._check(lib().okb_window_download(self._h, int(win), _p(poses), _p(sb), _p(lms), _p(q)))
        return dict(poses=poses, speed_bias=sb, landmarks=lms, quality=q)

    def upload_batch(self, first, windows, host_threads=0, descs=None):
        """okb_window_upload_batch; `descs` (from make_descs) may be cached by the caller while the host arrays live."""
        if descs is None:
            descs = self.make_descs(windows)
        self._check(lib().okb_window_upload_batch(self._h, int(first), len(windows), descs, int(host_threads)))
        for i, w in enumerate(windows):
            self._windows[first + i] = w

    @staticmethod
    def make_descs(windows):
        arr = (abi.WindowDesc * len(windows))()
        for i, w in enumerate(windows):
            arr[i] = w.desc()
        arr._keep = list(windows)
        return arr

    def alloc_outputs(self, first, count):
        """Host buffers for download_batch plus the pointer tables over them."""
        outs = []
        for i in range(first, first + count):
            w = self._windows[i]
            outs.append(dict(poses=np.zeros_like(w.poses), speed_bias=np.zeros_like(w.speed_bias),
                             landmarks=np.zeros_like(w.landmarks), quality=np.zeros(len(w.landmarks))))
        PP = C.POINTER(C.c_double) * count
        tables = [PP(*[o[k].ctypes.data_as(C.POINTER(C.c_double)) for o in outs])
                  for k in ("poses", "speed_bias", "landmarks", "quality")]
        return outs, tables

    def download_batch(self, first, count, out=None):
        outs, tables = out if out is not None else self.alloc_outputs(first, count)
        self._check(lib().okb_window_download_batch(self._h, int(first), int(count), *tables))
        return outs

    # ---------------------------------------------------------------- single-block hooks
    def eval_reprojection(self, cam, pose, lm, ext, z, sqrt_info):
        pose, lm, ext, z, sqrt_info = map(_f64, (pose, lm, ext, z, sqrt_info))
        n = len(pose)
        cam_arr = np.array([cam], dtype=abi.camera_dtype)
        r, J0, J1, J2 = np.zeros((n, 2)), np.zeros((n, 2, 6)), np.zeros((n, 2, 3)), np.zeros((n, 2, 6))
        self._check(lib().okb_eval_reprojection(self._h, n, _p(cam_arr), _p(pose), _p(lm), _p(ext), _p(z),
                                                _p(sqrt_info), _p(r), _p(J0), _p(J1), _p(J2)))
        return r, J0, J1, J2

    def eval_imu(self, params, samples, t0_ns, t1_ns, pose0, sb0, pose1, sb1, sb_ref=None):
        a = [_f64(x) for x in (pose0, sb0, pose1, sb1)]
        ref = _f64(sb_ref) if sb_ref is not None else None
        samples = np.ascontiguousarray(samples)
        r = np.zeros(15)
        J = [np.zeros((15, 6)), np.zeros((15, 9)), np.zeros((15, 6)), np.zeros((15, 9))]
        sq = np.zeros((15, 15))
        rc = lib().okb_eval_imu(self._h, C.byref(params), _p(samples), len(samples), C.c_int64(int(t0_ns)),
                                C.c_int64(int(t1_ns)), *[_p(x) for x in a], _p(ref), _p(r), *[_p(j) for j in J], _p(sq))
        if rc < 0:
            self._check(rc)
        return r, J, sq, rc

    def imu_propagate(self, params, samples, t0_ns, t1_ns, pose, sb, want_cov=True):
        pose, sb = np.array(pose, dtype=np.float64), np.array(sb, dtype=np.float64)
        samples = np.ascontiguousarray(samples)
        P, F = np.zeros((15, 15)), np.zeros((15, 15))
        n = C.c_int(0)
        self._check(lib().okb_imu_propagate(self._h, C.byref(params), _p(samples), len(samples), C.c_int64(int(t0_ns)),
                                            C.c_int64(int(t1_ns)), _p(pose), _p(sb), _p(P) if want_cov else None, _p(F),
                                            C.byref(n)))
        return n.value, pose, sb, P, F

    def eval_pose_error(self, meas, sqrt_info, pose):
        r, J = np.zeros(6), np.zeros((6, 6))
        self._check(lib().okb_eval_pose_error(self._h, _p(_f64(meas)), _p(_f64(sqrt_info)), _p(_f64(pose)), _p(r), _p(J)))
        return r, J

    def eval_speed_bias_error(self, meas, sqrt_info, sb):
        r, J = np.zeros(9), np.zeros((9, 9))
        self._check(lib().okb_eval_speed_bias_error(self._h, _p(_f64(meas)), _p(_f64(sqrt_info)), _p(_f64(sb)), _p(r),
                                                    _p(J)))
        return r, J

    def eval_relative_pose(self, sqrt_info, pose0, pose1):
        r, J0, J1 = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
        self._check(lib().okb_eval_relative_pose(self._h, _p(_f64(sqrt_info)), _p(_f64(pose0)), _p(_f64(pose1)), _p(r),
                                                 _p(J0), _p(J1)))
        return r, J0, J1

    def eval_marginalization(self, marg, x):
        m = abi.MargPrior()
        m.n, m.n_blocks = int(marg["J"].shape[0]), len(marg["block_kind"])
        m.block_kind = marg["block_kind"].ctypes.data_as(C.POINTER(C.c_int32))
        m.block_idx = marg["block_idx"].ctypes.data_as(C.POINTER(C.c_uint32))
        m.x0, m.J, m.e0 = abi.dptr(marg["x0"]), abi.dptr(marg["J"]), abi.dptr(marg["e0"])
        r, J = np.zeros(m.n), np.zeros((m.n, m.n))
        self._check(lib().okb_eval_marginalization(self._h, C.byref(m), _p(_f64(x)), _p(r), _p(J)))
        return r, J

    # ---------------------------------------------------------------- frontend path
    def hamming_match(self, A, B, skipA=None, skipB=None, threshold=60.0, num_best=4, use_ratio=False,
                      ratio_threshold=3.0):
        A = np.ascontiguousarray(A, dtype=np.uint8)
        B = np.ascontiguousarray(B, dtype=np.uint8)
        nA, nB, nbytes = A.shape[0], B.shape[0], A.shape[1]
        topk = np.zeros((nA, num_best), abi.pair_dtype)
        pairs = np.zeros(nB, abi.pair_dtype)
        sa = np.ascontiguousarray(skipA, dtype=np.uint8) if skipA is not None else None
        sb = np.ascontiguousarray(skipB, dtype=np.uint8) if skipB is not None else None
        self._check(lib().okb_hamming_match(self._h, _p(A), nA, _p(B), nB, nbytes, _p(sa), _p(sb), C.c_float(threshold),
                                            num_best, int(use_ratio), C.c_float(ratio_threshold), _p(topk), _p(pairs)))
        return dict(topk=topk, pairs=pairs)

    def hamming_match_gated(self, A, B, gate, skipA=None, skipB=None, threshold=60.0, num_best=4, use_ratio=False, ratio_threshold=3.0):
        """okb_hamming_match_gated; gate from abi.make_match_gate."""
        A = np.ascontiguousarray(A, dtype=np.uint8)
        B = np.ascontiguousarray(B, dtype=np.uint8)
        nA, nB, nbytes = A.shape[0], B.shape[0], A.shape[1]
        topk = np.zeros((nA, num_best), abi.pair_dtype)
        pairs = np.zeros(nB, abi.pair_dtype)
        sa = np.ascontiguousarray(skipA, dtype=np.uint8) if skipA is not None else None
        sb = np.ascontiguousarray(skipB, dtype=np.uint8) if skipB is not None else None
        self._check(lib().okb_hamming_match_gated(self._h, _p(A), nA, _p(B), nB, nbytes, _p(sa), _p(sb), C.c_float(threshold), num_best,
                                                  int(use_ratio), C.c_float(ratio_threshold), C.byref(gate), _p(topk), _p(pairs)))
        return dict(topk=topk, pairs=pairs)

    def hamming_candidates(self, A, B, threshold=60.0, cap=None):
        A = np.ascontiguousarray(A, dtype=np.uint8)
        B = np.ascontiguousarray(B, dtype=np.uint8)
        nA, nB, nbytes = A.shape[0], B.shape[0], A.shape[1]
        cap = cap or max(nA * nB, 1)
        row_ptr = np.zeros(nA + 1, np.uint32)
        col = np.zeros(cap, np.uint32)
        dist = np.zeros(cap, np.uint16)
        rc = lib().okb_hamming_candidates(self._h, _p(A), nA, _p(B), nB, nbytes, C.c_float(threshold), _p(row_ptr),
                                          _p(col), _p(dist), cap)
        self._check(rc)
        n = int(row_ptr[-1])
        return row_ptr, col[:n].copy(), dist[:n].copy()

    def detect_describe(self, img, cam, R_CW, uniformity_radius=40.0, absolute_threshold=800.0, max_keypoints=400,
                        desc_bytes=48, rotation_invariance=True, cam_slot=0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        prm = abi.DetectParams()
        prm.uniformity_radius, prm.absolute_threshold = uniformity_radius, absolute_threshold
        prm.max_keypoints, prm.desc_bytes, prm.rotation_invariance = max_keypoints, desc_bytes, int(rotation_invariance)
        cam_arr = np.array([cam], dtype=abi.camera_dtype)
        R = _f64(np.asarray(R_CW).reshape(9))
        kps = np.zeros(max_keypoints, abi.keypoint_dtype)
        desc = np.zeros((max_keypoints, desc_bytes), np.uint8)
        n = C.c_int(0)
        self._check(lib().okb_detect_describe(self._h, int(cam_slot), _p(img), w, h, img.strides[0], _p(cam_arr), _p(R),
                                              C.byref(prm), _p(kps), _p(desc), max_keypoints, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()


def matches_from_pairs(pairs, topk, threshold, use_ratio=False, ratio_threshold=3.0):
    """The serial epilogue of DenseMatcher::matchBody (DenseMatcher.hpp(impl):97-122): turns the per-B
    winners into the ordered (A, B, distance) callback list.  Pure bookkeeping on the host side of the
    boundary, exactly where the reference runs setBestMatch."""
    out = []
    for b in range(len(pairs)):
        a, dist = int(pairs[b]["index_a"]), float(pairs[b]["distance"])
        if not dist < threshold:
            continue
        if use_ratio:
            bl = topk[a]
            if bl[1]["index_a"] != -1:
                d0, d1 = float(bl[0]["distance"]), float(bl[1]["distance"])
                if not (d0 == 0 or d1 / d0 > ratio_threshold):
                    continue
        out.append((a, b, dist))
    return out
