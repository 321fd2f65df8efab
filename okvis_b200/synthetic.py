"""Seeded synthetic keyframe windows (structure after the reference's
okvis_ceres/test/TestEstimator.cpp:60-203): smooth trajectory, 200 Hz IMU with EuRoC noise values
(config/config_fpga_p2_euroc.yaml:35-46), EuRoC stereo calibration (:2-23), landmarks in a box with
per-landmark visibility runs, pixel noise + outliers, perturbed initial guess, first-frame priors
(okvis_ceres/src/Estimator.cpp:238-285).

Everything is numpy; the arrays use the C layouts of include/okvis_b200.h so a window can be handed to
the C-ABI (and, in tests, to the oracle) without conversion.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi

# EuRoC calibration (config/config_fpga_p2_euroc.yaml:2-23 of the reference)
_T_SC0 = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
                   [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
                   [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
                   [0.0, 0.0, 0.0, 1.0]])
_T_SC1 = np.array([[0.0125552670891, -0.999755099723, 0.0182237714554, -0.0198435579556],
                   [0.999598781151, 0.0130119051815, 0.0251588363115, 0.0453689425024],
                   [-0.0253898008918, 0.0179005838253, 0.999517347078, 0.00786212447038],
                   [0.0, 0.0, 0.0, 1.0]])
_CAM0 = dict(fu=458.654880721, fv=457.296696463, cu=367.215803962, cv=248.37534061,
             dist=(-0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05))
_CAM1 = dict(fu=457.587426604, fv=456.13442556, cu=379.99944652, cv=255.238185386,
             dist=(-0.283683654496, 0.0745128430929, -0.000104738949098, -3.55590700274e-05))


# ------------------------------------------------------------------ small SO(3) helpers (data generation only)
def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def quat_from_R(R):
    """Rotation matrix -> unit quaternion [x,y,z,w]."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def R_from_quat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def delta_q(v):
    a = np.linalg.norm(v)
    if a < 1e-12:
        return np.array([0.5 * v[0], 0.5 * v[1], 0.5 * v[2], 1.0])
    return np.concatenate([np.sin(a / 2) * v / a, [np.cos(a / 2)]])


def pose_oplus(pose, d6):
    """Left-multiplicative world-frame perturbation, as Transformation::oplus."""
    q = quat_mul(delta_q(d6[3:]), pose[3:])
    return np.concatenate([pose[:3] + d6[:3], q / np.linalg.norm(q)])


def T_to_pose7(T):
    return np.concatenate([T[:3, 3], quat_from_R(T[:3, :3])])


# ------------------------------------------------------------------ camera projection (generation only)
def _distort_radtan(u, k):
    k1, k2, p1, p2 = k[:4]
    mx, my, mxy = u[..., 0] ** 2, u[..., 1] ** 2, u[..., 0] * u[..., 1]
    rho = mx + my
    rad = k1 * rho + k2 * rho * rho
    return np.stack([u[..., 0] + u[..., 0] * rad + 2 * p1 * mxy + p2 * (rho + 2 * mx),
                     u[..., 1] + u[..., 1] * rad + 2 * p2 * mxy + p1 * (rho + 2 * my)], -1)


def _distort_equi(u, k):
    r = np.sqrt(u[..., 0] ** 2 + u[..., 1] ** 2)
    th = np.arctan(r)
    thd = th * (1 + k[0] * th ** 2 + k[1] * th ** 4 + k[2] * th ** 6 + k[3] * th ** 8)
    s = np.where(r > 1e-8, thd / np.maximum(r, 1e-300), 1.0)
    return u * s[..., None]


def project_points(cam, p_C):
    """p_C (...,3) -> pixel (...,2), success mask (in front, inside the image)."""
    z = p_C[..., 2]
    zs = np.where(np.abs(z) < 1e-12, 1e-12, z)
    u = p_C[..., :2] / zs[..., None]
    model = int(cam["model"])
    if model == abi.DIST_RADTAN:
        dd = _distort_radtan(u, cam["dist"])
    elif model == abi.DIST_EQUIDISTANT:
        dd = _distort_equi(u, cam["dist"])
    elif model == abi.DIST_NONE:
        dd = u
    else:
        raise ValueError("generator supports none/radtan/equidistant")
    px = np.stack([cam["fu"] * dd[..., 0] + cam["cu"], cam["fv"] * dd[..., 1] + cam["cv"]], -1)
    ok = (z > 0.0) & (px[..., 0] >= 0) & (px[..., 1] >= 0) & (px[..., 0] < cam["width"]) & (px[..., 1] < cam["height"])
    # the radtan polynomial folds back far outside the field of view; keep only the monotone region
    ok &= (u[..., 0] ** 2 + u[..., 1] ** 2) < 1.2
    return px, ok


# ------------------------------------------------------------------ window container
@dataclass
class Window:
    poses: np.ndarray
    speed_bias: np.ndarray
    extrinsics: np.ndarray
    extrinsics_fixed: np.ndarray
    landmarks: np.ndarray
    cameras: np.ndarray
    obs: np.ndarray
    imu_terms: np.ndarray
    imu_samples: np.ndarray
    imu_params: abi.ImuParams
    pose_priors: np.ndarray
    sb_priors: np.ndarray
    relpose_terms: np.ndarray
    marg: dict = None           # optional {"block_kind","block_idx","x0","J","e0"}
    truth: dict = field(default_factory=dict)
    name: str = ""

    def desc(self):
        """ctypes okb_window_desc borrowing this window's arrays (keep `self` alive while in use)."""
        d = abi.WindowDesc()
        d.n_poses, d.n_speed_bias, d.n_extrinsics = len(self.poses), len(self.speed_bias), len(self.extrinsics)
        d.n_landmarks, d.n_cameras, d.n_obs = len(self.landmarks), len(self.cameras), len(self.obs)
        d.n_imu_terms, d.n_imu_samples = len(self.imu_terms), len(self.imu_samples)
        d.n_pose_priors, d.n_sb_priors, d.n_relpose_terms = len(self.pose_priors), len(self.sb_priors), len(
            self.relpose_terms)
        for name in ("poses", "speed_bias", "extrinsics", "extrinsics_fixed", "landmarks", "cameras", "obs",
                     "imu_terms", "imu_samples", "pose_priors", "sb_priors", "relpose_terms"):
            arr = getattr(self, name)
            setattr(d, name, arr.ctypes.data if len(arr) else None)
        d.imu_params = self.imu_params
        if self.marg is not None:
            m = getattr(self, "_marg_c", None)
            if m is None:       # built once: descriptors copied elsewhere keep pointing at this struct
                m = abi.MargPrior()
                mk = self.marg
                m.n, m.n_blocks = int(mk["J"].shape[0]), len(mk["block_kind"])
                m.block_kind = mk["block_kind"].ctypes.data_as(C.POINTER(C.c_int32))
                m.block_idx = mk["block_idx"].ctypes.data_as(C.POINTER(C.c_uint32))
                m.x0, m.J, m.e0 = abi.dptr(mk["x0"]), abi.dptr(mk["J"]), abi.dptr(mk["e0"])
                self._marg_c = m
            d.marg = C.pointer(m)
        return d

    # algorithmic bytes / flops per window-iteration, SURVEY.md 8(d)
    def algorithmic_bytes_per_iteration(self):
        K, L, N = len(self.poses), len(self.landmarks), len(self.obs)
        dd = 6 * K + 9 * len(self.speed_bias)
        return 36 * N + 32 * L + 96 * L + 128 * K + 56 * len(self.extrinsics) + 8 * (dd * dd + dd) + 2400 * len(
            self.imu_terms)

    def algorithmic_flops_per_iteration(self):
        K, L, N = len(self.poses), len(self.landmarks), len(self.obs)
        dd = 6 * K + 9 * len(self.speed_bias)
        f = np.zeros(L)
        seen = set(zip(self.obs["lm_idx"].tolist(), self.obs["pose_idx"].tolist()))
        for l, _ in seen:
            f[l] += 1
        return 1200 * N + 300 * N + 324 * float(np.sum(f * f)) + dd ** 3 / 3 + 14000 * len(self.imu_terms)


@dataclass
class WindowConfig:
    n_frames: int = 10
    n_cams: int = 2
    n_landmarks: int = 2000
    frame_dt: float = 0.25
    imu_rate: int = 200
    distortion: int = abi.DIST_RADTAN
    pixel_sigma: float = 0.8
    outlier_fraction: float = 0.02
    frame_time_offset_ns: int = 1_700_000   # frames do not coincide with IMU samples (exercises interpolation)
    pose_prior_quirk: bool = False          # True: reference's failed-LLT sqrt-information
    perturb: bool = True
    with_marg_prior: bool = False


CONFIGS = {
    1: WindowConfig(n_frames=5, n_cams=1, n_landmarks=300, frame_dt=0.125),
    2: WindowConfig(n_frames=10, n_cams=2, n_landmarks=2000, frame_dt=0.25),
    4: WindowConfig(n_frames=10, n_cams=2, n_landmarks=2000, frame_dt=0.25),
    5: WindowConfig(n_frames=20, n_cams=4, n_landmarks=8000, frame_dt=0.25),
}

_R_BASE = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]])  # S-z (optical axis) -> W-y (travel)


def _traj(t, phases):
    """World pose / velocity / acceleration of the sensor frame at times t (array)."""
    t = np.asarray(t, dtype=np.float64)
    v = 1.0
    r = np.stack([0.2 * np.sin(0.7 * t), v * t, 0.2 * np.sin(1.1 * t)], -1)
    vel = np.stack([0.2 * 0.7 * np.cos(0.7 * t), v * np.ones_like(t), 0.2 * 1.1 * np.cos(1.1 * t)], -1)
    acc = np.stack([-0.2 * 0.49 * np.sin(0.7 * t), np.zeros_like(t), -0.2 * 1.21 * np.sin(1.1 * t)], -1)
    ang = 0.1 * np.sin(0.5 * t[..., None] + phases)  # yaw, pitch, roll
    cy, sy, cp, sp, cr, sr = (np.cos(ang[..., 0]), np.sin(ang[..., 0]), np.cos(ang[..., 1]), np.sin(ang[..., 1]),
                              np.cos(ang[..., 2]), np.sin(ang[..., 2]))
    Rz = np.zeros(t.shape + (3, 3)); Ry = np.zeros_like(Rz); Rx = np.zeros_like(Rz)
    Rz[..., 0, 0], Rz[..., 0, 1], Rz[..., 1, 0], Rz[..., 1, 1], Rz[..., 2, 2] = cy, -sy, sy, cy, 1
    Ry[..., 0, 0], Ry[..., 0, 2], Ry[..., 2, 0], Ry[..., 2, 2], Ry[..., 1, 1] = cp, sp, -sp, cp, 1
    Rx[..., 1, 1], Rx[..., 1, 2], Rx[..., 2, 1], Rx[..., 2, 2], Rx[..., 0, 0] = cr, -sr, sr, cr, 1
    R = Rz @ Ry @ Rx @ _R_BASE
    return r, vel, acc, R


def make_window(config_id=2, window_idx=0, cfg=None, seed=None):
    """Builds one seeded synthetic window.  Seed = 0x0B200 + 1000*config_id + window_idx."""
    cfg = cfg or CONFIGS[config_id]
    if seed is None:
        seed = 0x0B200 + 1000 * config_id + window_idx
    rng = np.random.Generator(np.random.PCG64(seed))
    K, Cn, L = cfg.n_frames, cfg.n_cams, cfg.n_landmarks
    phases = rng.uniform(0, 2 * np.pi, 3)
    imu = abi.make_imu_params(rate=cfg.imu_rate)
    g = imu.g

    # ---- cameras + extrinsics
    cams = np.zeros(Cn, abi.camera_dtype)
    ext = np.zeros((Cn, 7))
    base = [(_CAM0, _T_SC0), (_CAM1, _T_SC1)]
    for c in range(Cn):
        intr, T = base[c % 2]
        T = T.copy()
        if c >= 2:  # cfg-5: two more cameras rotated +-90 deg about S-z
            a = np.pi / 2 if c == 2 else -np.pi / 2
            Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
            T[:3, :3] = Rz @ T[:3, :3]
            T[:3, 3] = Rz @ T[:3, 3]
        if cfg.distortion == abi.DIST_RADTAN:
            dist = intr["dist"]
        elif cfg.distortion == abi.DIST_EQUIDISTANT:
            dist = (-0.21, 0.14, 0.0006, 0.0003)  # EquidistantDistortion.hpp:104-107 test coefficients
        else:
            dist = ()
        cams[c] = abi.make_camera(cfg.distortion, 752, 480, intr["fu"], intr["fv"], intr["cu"], intr["cv"], dist)
        ext[c] = T_to_pose7(T)
    ext_fixed = np.ones(Cn, np.uint8)

    # ---- trajectory at frame times
    t_base = 1.0
    t_frames_ns = (np.round(t_base * 1e9).astype(np.int64) + cfg.frame_time_offset_ns +
                   np.round(np.arange(K) * cfg.frame_dt * 1e9).astype(np.int64))
    t_frames = t_frames_ns * 1e-9
    r_true, v_true, _, R_true = _traj(t_frames, phases)
    poses_true = np.stack([np.concatenate([r_true[k], quat_from_R(R_true[k])]) for k in range(K)])
    bg_true = rng.normal(0, 0.01, 3)
    ba_true = rng.normal(0, 0.05, 3)
    sb_true = np.concatenate([v_true, np.tile(bg_true, (K, 1)), np.tile(ba_true, (K, 1))], 1)

    # ---- IMU samples on the global 1/rate grid, covering [t_0 - dt, t_{K-1} + dt]
    dt_ns = int(round(1e9 / cfg.imu_rate))
    n0 = int(t_frames_ns[0] // dt_ns) - 1
    n1 = int(-(-t_frames_ns[-1] // dt_ns)) + 1
    ts_ns = np.arange(n0, n1 + 1, dtype=np.int64) * dt_ns
    ts = ts_ns * 1e-9
    _, _, a_W, R_s = _traj(ts, phases)
    h = 1e-5
    _, _, _, R_p = _traj(ts + h, phases)
    _, _, _, R_m = _traj(ts - h, phases)
    dR = np.einsum("nji,njk->nik", R_s, (R_p - R_m) / (2 * h))  # R^T Rdot
    omega = np.stack([dR[:, 2, 1] - dR[:, 1, 2], dR[:, 0, 2] - dR[:, 2, 0], dR[:, 1, 0] - dR[:, 0, 1]], -1) * 0.5
    acc_S = np.einsum("nji,nj->ni", R_s, a_W + np.array([0, 0, g]))
    samples = np.zeros(len(ts), abi.imu_sample_dtype)
    samples["t_ns"] = ts_ns
    samples["gyro"] = omega + bg_true + rng.normal(0, imu.sigma_g_c * np.sqrt(cfg.imu_rate), omega.shape)
    samples["acc"] = acc_S + ba_true + rng.normal(0, imu.sigma_a_c * np.sqrt(cfg.imu_rate), acc_S.shape)
    terms = np.zeros(K - 1, abi.imu_term_dtype)
    for k in range(K - 1):
        lo = int(np.searchsorted(ts_ns, t_frames_ns[k], side="right")) - 1       # last sample <= t_k
        hi = int(np.searchsorted(ts_ns, t_frames_ns[k + 1], side="left"))        # first sample >= t_{k+1}
        lo = max(lo - (1 if ts_ns[lo] == t_frames_ns[k] else 0), 0)
        hi = min(hi + (1 if ts_ns[hi] == t_frames_ns[k + 1] else 0), len(ts_ns) - 1)
        terms[k] = (k, k, k + 1, k + 1, t_frames_ns[k], t_frames_ns[k + 1], lo, hi - lo + 1)

    # ---- landmarks with visibility runs
    T_CS = []
    for c in range(Cn):
        Rsc = R_from_quat(ext[c, 3:])
        T_CS.append((Rsc.T, -Rsc.T @ ext[c, :3]))
    centre = r_true.mean(0)
    lms_true = np.zeros((0, 3))
    vis_all = np.zeros((0, K, Cn), bool)
    px_all = np.zeros((0, K, Cn, 2))
    while len(lms_true) < L:
        n_try = max(4 * (L - len(lms_true)), 256)
        p = centre + rng.uniform([-10, -10, -3], [10, 10, 3], (n_try, 3))
        s = rng.integers(0, max(K - 2, 1), n_try)
        ln = np.minimum(rng.integers(3, K + 1, n_try), K)
        s = np.minimum(s, K - ln)                 # runs clipped at the window end are shifted back
        in_run = (np.arange(K)[None, :] >= s[:, None]) & (np.arange(K)[None, :] < (s + ln)[:, None])
        vis = np.zeros((n_try, K, Cn), bool)
        px = np.zeros((n_try, K, Cn, 2))
        for k in range(K):
            p_S = (p - r_true[k]) @ R_true[k]          # R^T (p - r)
            for c in range(Cn):
                p_C = p_S @ T_CS[c][0].T + T_CS[c][1]
                pxc, ok = project_points(cams[c], p_C)
                ok &= p_C[:, 2] >= 0.5
                vis[:, k, c] = ok & in_run[:, k]
                px[:, k, c] = pxc
        keep = vis.reshape(n_try, -1).sum(1) >= 2
        lms_true = np.concatenate([lms_true, p[keep]])[:L]
        vis_all = np.concatenate([vis_all, vis[keep]])[:L]
        px_all = np.concatenate([px_all, px[keep]])[:L]
    li, ki, ci = np.nonzero(vis_all)
    n_obs = len(li)
    obs = np.zeros(n_obs, abi.observation_dtype)
    obs["pose_idx"], obs["lm_idx"], obs["ext_idx"], obs["cam_idx"] = ki, li, ci, ci
    z = px_all[li, ki, ci] + rng.normal(0, cfg.pixel_sigma, (n_obs, 2))
    outl = rng.random(n_obs) < cfg.outlier_fraction
    z[outl] = rng.uniform([0, 0], [752, 480], (int(outl.sum()), 2))
    obs["z"] = z
    obs["sqrt_info"] = 1.0   # keypoint size 8 -> 64/size^2 = 1

    # ---- initial guess
    poses = poses_true.copy()
    sb = sb_true.copy()
    sb[:, 3:] = 0.0
    lm_init = lms_true.copy()
    if cfg.perturb:
        for k in range(K):
            poses[k] = pose_oplus(poses_true[k], np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.01, 3)]))
        sb[:, :3] += rng.normal(0, 0.05, (K, 3))
        lm_init = lms_true + rng.normal(0, 0.1, lms_true.shape)
    hp = np.concatenate([lm_init, np.ones((L, 1))], 1)
    hp /= np.linalg.norm(hp, axis=1, keepdims=True)

    # ---- priors of the first frame (Estimator.cpp:238-285)
    pp = np.zeros(1, abi.pose_prior_dtype)
    pp["pose_idx"] = 0
    pp["meas"] = poses[0]
    sq = np.diag([1e4, 1e4, 1e4, 0, 0, 1e8 if cfg.pose_prior_quirk else 1e4])
    pp["sqrt_info"] = sq.reshape(-1)
    sp = np.zeros(1, abi.sb_prior_dtype)
    sp["sb_idx"] = 0
    sp["meas"] = sb[0]
    sp["sqrt_info"] = np.diag([1, 1, 1] + [1 / imu.sigma_bg] * 3 + [1 / imu.sigma_ba] * 3).reshape(-1)

    w = Window(poses=np.ascontiguousarray(poses), speed_bias=np.ascontiguousarray(sb),
               extrinsics=np.ascontiguousarray(ext), extrinsics_fixed=ext_fixed,
               landmarks=np.ascontiguousarray(hp), cameras=cams, obs=obs, imu_terms=terms, imu_samples=samples,
               imu_params=imu, pose_priors=pp, sb_priors=sp, relpose_terms=np.zeros(0, abi.relpose_dtype),
               truth=dict(poses=poses_true, speed_bias=sb_true, landmarks=lms_true),
               name="cfg%d/w%d" % (config_id, window_idx))
    if cfg.with_marg_prior:
        w.marg = make_random_marg_prior(w, rng)
    return w


def make_random_marg_prior(w, rng, n_pose_blocks=3):
    """A random SPD linearised prior over the first `n_pose_blocks` poses and speed/bias 0, in the
    (J, e0, x0) form MarginalizationError::updateErrorComputation produces (dimension 6*n+9)."""
    n = 6 * n_pose_blocks + 9
    A = rng.normal(0, 1.0, (n + 8, n))
    scale = np.concatenate([[30.0] * 3 + [100.0] * 3] * n_pose_blocks + [[5.0] * 3 + [50.0] * 3 + [10.0] * 3])
    H = (A.T @ A) / (n + 8) * np.outer(scale, scale)
    lam, U = np.linalg.eigh(0.5 * (H + H.T))
    J = (U * np.sqrt(np.maximum(lam, 0))).T
    e0 = rng.normal(0, 0.3, n)
    kinds = np.array([abi.BLOCK_POSE] * n_pose_blocks + [abi.BLOCK_SPEED_BIAS], np.int32)
    idx = np.array(list(range(n_pose_blocks)) + [0], np.uint32)
    x0 = np.concatenate([w.poses[:n_pose_blocks].reshape(-1), w.speed_bias[0]])
    x0 = x0 + 0.0  # linearisation point = initial guess
    return dict(block_kind=kinds, block_idx=idx, x0=np.ascontiguousarray(x0), J=np.ascontiguousarray(J),
                e0=np.ascontiguousarray(e0))
