"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the CPU oracle (oracle/liboracle.so).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (okvis_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from okvis_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oko_problem_create.restype = C.c_void_p
        _lib.oko_problem_create.argtypes = [C.POINTER(abi.WindowDesc)]
        _lib.oko_problem_destroy.argtypes = [C.c_void_p]
        _lib.oko_problem_destroy.restype = None
        _lib.oko_solve.argtypes = [C.c_void_p, C.POINTER(abi.SolveOptions), C.c_int, C.POINTER(abi.Summary),
                                   C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        _lib.oko_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.oko_get_state.restype = None
        _lib.oko_cost.argtypes = [C.c_void_p]
        _lib.oko_cost.restype = C.c_double
    return _lib


def max_threads():
    return lib().oko_max_threads()


def options(max_iterations=10, min_iterations=0, time_limit_s=-1.0, use_cauchy_loss=1):
    o = abi.SolveOptions()
    o.max_iterations, o.min_iterations, o.time_limit_s, o.use_cauchy_loss = (max_iterations, min_iterations,
                                                                            time_limit_s, use_cauchy_loss)
    return o


class OracleProblem:
    """The oracle's okvis::Estimator stand-in for one window."""

    def __init__(self, window):
        self.window = window
        self._desc = window.desc()
        self._p = lib().oko_problem_create(C.byref(self._desc))

    def close(self):
        if self._p:
            lib().oko_problem_destroy(self._p)
            self._p = None

    def __del__(self):
        self.close()

    def cost(self):
        return lib().oko_cost(self._p)

    def solve(self, max_iterations=10, num_threads=1, use_cauchy_loss=1, min_iterations=0, time_limit_s=-1.0):
        opt = options(max_iterations, min_iterations, time_limit_s, use_cauchy_loss)
        s = abi.Summary()
        trace = np.zeros((max(max_iterations, 1) + 2, 6))
        n = C.c_int(0)
        phases = np.zeros(7)
        lib().oko_solve(self._p, C.byref(opt), num_threads, C.byref(s), trace.ctypes.data, len(trace), C.byref(n),
                        phases.ctypes.data)
        out = s.as_dict()
        out["trace"] = trace[:n.value].copy()
        out["phase_times"] = dict(zip(["evaluate_jac", "schur", "reduced_solve", "backsub", "evaluate_cost",
                                       "quality", "other"], phases.tolist()))
        return out

    def set_state(self, poses=None, speed_bias=None, landmarks=None):
        a = [np.ascontiguousarray(x, np.float64) if x is not None else None for x in (poses, speed_bias, landmarks)]
        f = lib().oko_set_state
        f.restype = None
        f(C.c_void_p(self._p), *[C.c_void_p(x.ctypes.data) if x is not None else None for x in a])

    def set_imu_cache(self, term, sb_ref, valid=True):
        r = np.ascontiguousarray(sb_ref, np.float64)
        assert lib().oko_set_imu_cache(C.c_void_p(self._p), int(term), C.c_void_p(r.ctypes.data), int(bool(valid))) == 0

    def marginalize(self, job, H_prev=None, b0_prev=None):
        """One marginalisation step (oracle_marg_apply.hpp) on the problem's graph at its current estimates."""
        kind, idx = np.zeros(64, np.int32), np.zeros(64, np.uint32)
        x0, J, e0, H, b0 = np.zeros(9 * 64), np.zeros(160 * 160), np.zeros(160), np.zeros(160 * 160), np.zeros(160)
        rank = C.c_int(0)
        f = lib().oko_marginalize
        f.restype = C.c_int
        hp = np.ascontiguousarray(H_prev, np.float64) if H_prev is not None else None
        bp = np.ascontiguousarray(b0_prev, np.float64) if b0_prev is not None else None
        n = f(C.c_void_p(self._p), C.byref(job), C.c_void_p(hp.ctypes.data) if hp is not None else None,
              C.c_void_p(bp.ctypes.data) if bp is not None else None, *[C.c_void_p(a.ctypes.data) for a in (kind, idx, x0, J, e0, H, b0)],
              C.byref(rank))
        if n < 0:
            raise RuntimeError("oracle marginalize: a residual's parameter block is missing from the job")
        nb = sum(1 for m in job._keep[3] if not m)
        kind, idx = kind[:nb].copy(), idx[:nb].copy()
        xdim = int(sum(9 if k == abi.BLOCK_SPEED_BIAS else 7 for k in kind))
        return dict(n=n, block_kind=kind, block_idx=idx, x0=x0[:xdim].copy(), J=J[:n * n].reshape(n, n).copy(), e0=e0[:n].copy(),
                    H=H[:n * n].reshape(n, n).copy(), b0=b0[:n].copy(), rank=rank.value)

    def state(self, with_quality=True):
        w = self.window
        poses = np.zeros_like(w.poses)
        sb = np.zeros_like(w.speed_bias)
        lms = np.zeros_like(w.landmarks)
        q = np.zeros(len(w.landmarks)) if with_quality else None
        lib().oko_get_state(self._p, poses.ctypes.data, sb.ctypes.data, lms.ctypes.data,
                            q.ctypes.data if with_quality else None)
        return dict(poses=poses, speed_bias=sb, landmarks=lms, quality=q)


# ---- single-functor hooks (numpy in / numpy out) -------------------------------------------------
def eval_reprojection(cam, pose, lm, ext, z, sqrt_info):
    n = len(pose)
    cam_arr = np.array([cam], dtype=abi.camera_dtype)
    r, J0, J1, J2 = np.zeros((n, 2)), np.zeros((n, 2, 6)), np.zeros((n, 2, 3)), np.zeros((n, 2, 6))
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pose, lm, ext, z, sqrt_info)]
    f = lib().oko_eval_reprojection
    f.restype = None
    f(C.c_int(n), C.c_void_p(cam_arr.ctypes.data), *[C.c_void_p(a.ctypes.data) for a in args],
      C.c_void_p(r.ctypes.data), C.c_void_p(J0.ctypes.data), C.c_void_p(J1.ctypes.data), C.c_void_p(J2.ctypes.data))
    return r, J0, J1, J2


def eval_imu(params, samples, t0_ns, t1_ns, pose0, sb0, pose1, sb1, sb_ref=None):
    r = np.zeros(15)
    J = [np.zeros((15, 6)), np.zeros((15, 9)), np.zeros((15, 6)), np.zeros((15, 9))]
    sq = np.zeros((15, 15))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pose0, sb0, pose1, sb1)]
    ref = np.ascontiguousarray(sb_ref, dtype=np.float64) if sb_ref is not None else None
    samples = np.ascontiguousarray(samples)
    f = lib().oko_eval_imu
    f.restype = C.c_int
    redo = f(C.byref(params), C.c_void_p(samples.ctypes.data), C.c_int(len(samples)), C.c_int64(int(t0_ns)),
             C.c_int64(int(t1_ns)), *[C.c_void_p(x.ctypes.data) for x in a],
             C.c_void_p(ref.ctypes.data) if ref is not None else None, C.c_void_p(r.ctypes.data),
             *[C.c_void_p(j.ctypes.data) for j in J], C.c_void_p(sq.ctypes.data))
    return r, J, sq, redo


def imu_propagate(params, samples, t0_ns, t1_ns, pose, sb, want_cov=True):
    pose = np.array(pose, dtype=np.float64)
    sb = np.array(sb, dtype=np.float64)
    P, F = np.zeros((15, 15)), np.zeros((15, 15))
    samples = np.ascontiguousarray(samples)
    f = lib().oko_imu_propagate
    f.restype = C.c_int
    n = f(C.byref(params), C.c_void_p(samples.ctypes.data), C.c_int(len(samples)), C.c_int64(int(t0_ns)),
          C.c_int64(int(t1_ns)), C.c_void_p(pose.ctypes.data), C.c_void_p(sb.ctypes.data),
          C.c_void_p(P.ctypes.data) if want_cov else None, C.c_void_p(F.ctypes.data))
    return n, pose, sb, P, F


def eval_pose_error(meas, sqrt_info, pose):
    r, J = np.zeros(6), np.zeros((6, 6))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (meas, sqrt_info, pose)]
    f = lib().oko_eval_pose_error
    f.restype = None
    f(*[C.c_void_p(x.ctypes.data) for x in a], C.c_void_p(r.ctypes.data), C.c_void_p(J.ctypes.data))
    return r, J


def eval_speed_bias_error(meas, sqrt_info, sb):
    r, J = np.zeros(9), np.zeros((9, 9))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (meas, sqrt_info, sb)]
    f = lib().oko_eval_speed_bias_error
    f.restype = None
    f(*[C.c_void_p(x.ctypes.data) for x in a], C.c_void_p(r.ctypes.data), C.c_void_p(J.ctypes.data))
    return r, J


def eval_relative_pose(sqrt_info, pose0, pose1):
    r, J0, J1 = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (sqrt_info, pose0, pose1)]
    f = lib().oko_eval_relative_pose
    f.restype = None
    f(*[C.c_void_p(x.ctypes.data) for x in a], C.c_void_p(r.ctypes.data), C.c_void_p(J0.ctypes.data),
      C.c_void_p(J1.ctypes.data))
    return r, J0, J1


def make_marg_struct(marg):
    m = abi.MargPrior()
    m.n, m.n_blocks = int(marg["J"].shape[0]), len(marg["block_kind"])
    m.block_kind = marg["block_kind"].ctypes.data_as(C.POINTER(C.c_int32))
    m.block_idx = marg["block_idx"].ctypes.data_as(C.POINTER(C.c_uint32))
    m.x0, m.J, m.e0 = abi.dptr(marg["x0"]), abi.dptr(marg["J"]), abi.dptr(marg["e0"])
    return m


def eval_marginalization(marg, x):
    m = make_marg_struct(marg)
    n = m.n
    r, J = np.zeros(n), np.zeros((n, n))
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = lib().oko_eval_marginalization
    f.restype = None
    f(C.byref(m), C.c_void_p(x.ctypes.data), C.c_void_p(r.ctypes.data), C.c_void_p(J.ctypes.data))
    return r, J


def pose_plus(x, delta):
    out = np.zeros(7)
    x = np.ascontiguousarray(x, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    f = lib().oko_pose_plus
    f.restype = None
    f(C.c_void_p(x.ctypes.data), C.c_void_p(delta.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def pose_minus(x, xpd):
    out = np.zeros(6)
    x = np.ascontiguousarray(x, dtype=np.float64)
    xpd = np.ascontiguousarray(xpd, dtype=np.float64)
    f = lib().oko_pose_minus
    f.restype = None
    f(C.c_void_p(x.ctypes.data), C.c_void_p(xpd.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def pose_lift_jacobian(x):
    J = np.zeros((6, 7))
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = lib().oko_pose_lift_jacobian
    f.restype = None
    f(C.c_void_p(x.ctypes.data), C.c_void_p(J.ctypes.data))
    return J


def pose_plus_jacobian(x):
    J = np.zeros((7, 6))
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = lib().oko_pose_plus_jacobian
    f.restype = None
    f(C.c_void_p(x.ctypes.data), C.c_void_p(J.ctypes.data))
    return J


def sqrt_information(info):
    info = np.ascontiguousarray(info, dtype=np.float64)
    n = info.shape[0]
    out = np.zeros((n, n))
    f = lib().oko_sqrt_information
    f.restype = C.c_int
    fail = f(C.c_void_p(info.ctypes.data), C.c_int(n), C.c_void_p(out.ctypes.data))
    return out, fail


def match_matrix(D, skipA=None, skipB=None, threshold=4.0, num_best=4, use_ratio=False, ratio_threshold=3.0):
    D = np.ascontiguousarray(D, dtype=np.float32)
    nA, nB = D.shape
    topk = np.zeros((nA, num_best), abi.pair_dtype)
    pairs = np.zeros(nB, abi.pair_dtype)
    matches = np.zeros((nB, 2), np.int32)
    md = np.zeros(nB, np.float32)
    sa = np.ascontiguousarray(skipA, dtype=np.uint8) if skipA is not None else None
    sb = np.ascontiguousarray(skipB, dtype=np.uint8) if skipB is not None else None
    f = lib().oko_match_matrix
    f.restype = C.c_int
    n = f(C.c_void_p(D.ctypes.data), C.c_int(nA), C.c_int(nB), C.c_void_p(sa.ctypes.data) if sa is not None else None,
          C.c_void_p(sb.ctypes.data) if sb is not None else None, C.c_float(threshold), C.c_int(num_best),
          C.c_int(int(use_ratio)), C.c_float(ratio_threshold), C.c_void_p(topk.ctypes.data),
          C.c_void_p(pairs.ctypes.data), C.c_void_p(matches.ctypes.data), C.c_void_p(md.ctypes.data))
    return dict(topk=topk, pairs=pairs, matches=matches[:n].copy(), distances=md[:n].copy())


def match_hamming(A, B, skipA=None, skipB=None, threshold=60.0, num_best=4, use_ratio=False, ratio_threshold=3.0):
    A = np.ascontiguousarray(A, dtype=np.uint8)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    nA, nB, nbytes = A.shape[0], B.shape[0], A.shape[1]
    topk = np.zeros((nA, num_best), abi.pair_dtype)
    pairs = np.zeros(nB, abi.pair_dtype)
    matches = np.zeros((nB, 2), np.int32)
    md = np.zeros(nB, np.float32)
    sa = np.ascontiguousarray(skipA, dtype=np.uint8) if skipA is not None else None
    sb = np.ascontiguousarray(skipB, dtype=np.uint8) if skipB is not None else None
    f = lib().oko_match_hamming
    f.restype = C.c_int
    n = f(C.c_void_p(A.ctypes.data), C.c_int(nA), C.c_void_p(B.ctypes.data), C.c_int(nB), C.c_int(nbytes),
          C.c_void_p(sa.ctypes.data) if sa is not None else None,
          C.c_void_p(sb.ctypes.data) if sb is not None else None, C.c_float(threshold), C.c_int(num_best),
          C.c_int(int(use_ratio)), C.c_float(ratio_threshold), C.c_void_p(topk.ctypes.data),
          C.c_void_p(pairs.ctypes.data), C.c_void_p(matches.ctypes.data), C.c_void_p(md.ctypes.data))
    return dict(topk=topk, pairs=pairs, matches=matches[:n].copy(), distances=md[:n].copy())


def match_hamming_gated(A, B, gate, skipA=None, skipB=None, threshold=60.0, num_best=4, use_ratio=False, ratio_threshold=3.0):
    """DenseMatcher over VioKeyframeWindowMatchingAlgorithm::distance (Hamming + verifyMatch), oracle_gate.hpp."""
    A = np.ascontiguousarray(A, dtype=np.uint8)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    nA, nB, nbytes = A.shape[0], B.shape[0], A.shape[1]
    topk = np.zeros((nA, num_best), abi.pair_dtype)
    pairs = np.zeros(nB, abi.pair_dtype)
    sa = np.ascontiguousarray(skipA, dtype=np.uint8) if skipA is not None else None
    sb = np.ascontiguousarray(skipB, dtype=np.uint8) if skipB is not None else None
    f = lib().oko_match_hamming_gated
    f.restype = C.c_int
    f(C.c_void_p(A.ctypes.data), C.c_int(nA), C.c_void_p(B.ctypes.data), C.c_int(nB), C.c_int(nbytes),
      C.c_void_p(sa.ctypes.data) if sa is not None else None, C.c_void_p(sb.ctypes.data) if sb is not None else None,
      C.c_float(threshold), C.c_int(num_best), C.c_int(int(use_ratio)), C.c_float(ratio_threshold), C.byref(gate),
      C.c_void_p(topk.ctypes.data), C.c_void_p(pairs.ctypes.data))
    return dict(topk=topk, pairs=pairs)


def hamming_candidates(A, B, threshold=60.0, cap=None):
    A = np.ascontiguousarray(A, dtype=np.uint8)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    nA, nB, nbytes = A.shape[0], B.shape[0], A.shape[1]
    cap = cap or nA * nB
    row_ptr = np.zeros(nA + 1, np.uint32)
    col = np.zeros(cap, np.uint32)
    dist = np.zeros(cap, np.uint16)
    f = lib().oko_hamming_candidates
    f.restype = C.c_int
    n = f(C.c_void_p(A.ctypes.data), C.c_int(nA), C.c_void_p(B.ctypes.data), C.c_int(nB), C.c_int(nbytes),
          C.c_float(threshold), C.c_void_p(row_ptr.ctypes.data), C.c_void_p(col.ctypes.data),
          C.c_void_p(dist.ctypes.data), C.c_int(cap))
    return row_ptr, col[:n].copy(), dist[:n].copy()


def detect_describe(img, cam, R_CW, uniformity_radius=40.0, absolute_threshold=800.0, max_keypoints=400, desc_bytes=48,
                    rotation_invariance=True):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    prm = abi.DetectParams()
    prm.uniformity_radius, prm.absolute_threshold = uniformity_radius, absolute_threshold
    prm.max_keypoints, prm.desc_bytes, prm.rotation_invariance = max_keypoints, desc_bytes, int(rotation_invariance)
    cam_arr = np.array([cam], dtype=abi.camera_dtype)
    R = np.ascontiguousarray(np.asarray(R_CW, dtype=np.float64).reshape(9))
    kps = np.zeros(max_keypoints, abi.keypoint_dtype)
    desc = np.zeros((max_keypoints, desc_bytes), np.uint8)
    f = lib().oko_detect_describe
    f.restype = C.c_int
    n = f(C.c_void_p(img.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(img.strides[0]), C.c_void_p(cam_arr.ctypes.data),
          C.c_void_p(R.ctypes.data), C.byref(prm), C.c_void_p(kps.ctypes.data), C.c_void_p(desc.ctypes.data),
          C.c_int(max_keypoints))
    return kps[:n].copy(), desc[:n].copy()


# ---- marginalisation numeric core (NEXT TIER groundwork: oracle only, no device path yet)
def _vp(a):
    return C.c_void_p(a.ctypes.data)


def marginalize_stage(H, b, ranges, landmark_blocks):
    """One stage of MarginalizationError::marginalizeOut (MarginalizationError.cpp:618-741) on a copy of (H, b)."""
    H = np.ascontiguousarray(H, dtype=np.float64).copy()
    b = np.ascontiguousarray(b, dtype=np.float64).copy()
    n = H.shape[0]
    r = np.ascontiguousarray(np.asarray(ranges, dtype=np.int32).reshape(-1, 2))
    f = lib().oko_marginalize_stage
    f.restype = C.c_int
    m = f(_vp(H), _vp(b), C.c_int(n), _vp(r), C.c_int(len(r)), C.c_int(int(bool(landmark_blocks))))
    return H.reshape(-1)[:m * m].reshape(m, m).copy(), b[:m].copy()


def marg_update_error_computation(H, b):
    """MarginalizationError::updateErrorComputation (MarginalizationError.cpp:806-846): returns J, e0, rank."""
    H = np.ascontiguousarray(H, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = H.shape[0]
    J, e0 = np.zeros((n, n)), np.zeros(n)
    f = lib().oko_marg_update_error_computation
    f.restype = C.c_int
    rank = f(_vp(H), _vp(b), C.c_int(n), _vp(J), _vp(e0))
    return J, e0, rank


def sym_eig(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    w, V = np.zeros(n), np.zeros((n, n))
    lib().oko_sym_eig(_vp(A), C.c_int(n), _vp(w), _vp(V))
    return w, V
