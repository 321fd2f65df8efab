#!/usr/bin/env python
"""Generator of tests/golden/numpy_costs.npz -- an INDEPENDENT pin of the cost functions of a keyframe window.

Everything here is plain NumPy / SciPy written from the reference's source (no liboracle.so, no CUDA library):
  * ReprojectionError residual            okvis_ceres/include/okvis/ceres/implementation/ReprojectionError.hpp:87-140
    (PinholeCamera<RadialTangentialDistortion>::projectHomogeneous; CauchyLoss(1) as a residual re-scaling whose squared
     norm is rho(s) = log(1 + s), so that 0.5 * sum ||r'||^2 is Ceres' robustified cost)
  * ImuError residual with a fresh preintegration at the CURRENT bias
                                          okvis_ceres/src/ImuError.cpp:76-284 (redoPreintegration), :514-560 (error)
  * PoseError / SpeedAndBiasError          okvis_ceres/src/PoseError.cpp:86-136, SpeedAndBiasError.cpp:89-116
The fixture holds a few states of two cfg-1 windows and the cost 0.5 * sum rho(||r||^2) of every residual family at
those states.  tests/test_golden.py checks the C++ oracle against it, tests/test_gpu_golden.py the device kernels
(okb_optimize with max_iterations = 0 reports the cost at the uploaded state).  Measured agreement: 1e-15 relative.
Independently of the fixture, at the state where the oracle's dogleg stops (FUNCTION_TOLERANCE after 36 iterations on
cfg-1 window 0) this objective evaluates to 446.196556366 against the oracle's 446.196556916 (1.2e-9: the reference
keeps the preintegration of an earlier bias and corrects to first order, ImuError.cpp:545-556, this script always
re-preintegrates) and its Jacobi-scaled gradient norm is 440 times smaller than at the initial guess.

Usage:  python tools/make_golden_numpy.py            (seconds)
        python tools/make_golden_numpy.py --optimize (scipy least_squares run, ~10 min, diagnostic only)"""
import os
import sys

import numpy as np
from scipy.optimize import least_squares
from scipy.sparse import lil_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from okvis_b200 import abi, synthetic  # noqa: E402  (window generator only: plain numpy)


# ---------------------------------------------------------------- kinematics (Hamilton quaternions, [x,y,z,w])
def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def qinv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]]) / (q @ q)


def q2R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def sinc(x):
    return np.sin(x) / x if abs(x) > 1e-6 else 1 - x * x / 6 + x ** 4 / 120


def delta_q(a):
    h = 0.5 * np.linalg.norm(a)
    return np.concatenate([sinc(h) * 0.5 * a, [np.cos(h)]])


def cross_mx(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def right_jacobian(phi):
    n = np.linalg.norm(phi)
    X = cross_mx(phi)
    if n < 1e-4:
        return np.eye(3) - 0.5 * X + X @ X / 6.0
    return np.eye(3) - (1 - np.cos(n)) / n ** 2 * X + (n - np.sin(n)) / n ** 3 * X @ X


def pose_plus(pose, d):
    """Transformation::oplus: t += d[0:3]; q <- normalize(deltaQ(d[3:6]) * q)."""
    q = qmul(delta_q(d[3:6]), pose[3:])
    return np.concatenate([pose[:3] + d[:3], q / np.linalg.norm(q)])


# ---------------------------------------------------------------- ImuError
def ns_to_sec(ns):
    return float(ns // 1000000000) + 1e-9 * float(ns % 1000000000)


def preintegrate(samples, prm, t0, t1, sb):
    t = np.array([int(x) for x in samples["t_ns"]], dtype=object)
    Dq = np.array([0.0, 0, 0, 1])
    C_int, C_dint = np.zeros((3, 3)), np.zeros((3, 3))
    a_int, a_dint = np.zeros(3), np.zeros(3)
    cross = np.zeros((3, 3))
    dalpha, dv, dp = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
    P = np.zeros((15, 15))
    time, Delta_t, started = t0, 0.0, False
    n = len(samples)
    for i in range(n - 1):
        w0, a0 = samples["gyro"][i].copy(), samples["acc"][i].copy()
        w1, a1 = samples["gyro"][i + 1].copy(), samples["acc"][i + 1].copy()
        nexttime = t[i + 1]
        dt = ns_to_sec(nexttime - time)
        if t1 < nexttime:
            interval = ns_to_sec(nexttime - t[i])
            nexttime = t1
            dt = ns_to_sec(nexttime - time)
            r = dt / interval
            w1 = (1 - r) * w0 + r * w1
            a1 = (1 - r) * a0 + r * a1
        if dt <= 0.0:
            continue
        Delta_t += dt
        if not started:
            started = True
            r = dt / ns_to_sec(nexttime - t[i])
            w0 = r * w0 + (1 - r) * w1
            a0 = r * a0 + (1 - r) * a1
        sg, sa = prm.sigma_g_c, prm.sigma_a_c
        if max(np.abs(w0).max(), np.abs(w1).max()) > prm.g_max:
            sg *= 100
        if max(np.abs(a0).max(), np.abs(a1).max()) > prm.a_max:
            sa *= 100
        w_true = 0.5 * (w0 + w1) - sb[3:6]
        th = np.linalg.norm(w_true) * 0.5 * dt
        dq = np.concatenate([sinc(th) * w_true * 0.5 * dt, [np.cos(th)]])
        Dq1 = qmul(Dq, dq)
        C, C1 = q2R(Dq), q2R(Dq1)
        a_true = 0.5 * (a0 + a1) - sb[6:9]
        C_int1 = C_int + 0.5 * (C + C1) * dt
        a_int1 = a_int + 0.5 * (C + C1) @ a_true * dt
        C_dint = C_dint + C_int * dt + 0.25 * (C + C1) * dt * dt
        a_dint = a_dint + a_int * dt + 0.25 * (C + C1) @ a_true * dt * dt
        Jr = right_jacobian(w_true * dt)
        dalpha = dalpha + C1 @ Jr * dt
        cross1 = q2R(qinv(dq)) @ cross + Jr * dt
        ax = cross_mx(a_true)
        dv1 = dv + 0.5 * dt * (C @ ax @ cross + C1 @ ax @ cross1)
        dp = dp + dt * dv + 0.25 * dt * dt * (C @ ax @ cross + C1 @ ax @ cross1)
        F = np.eye(15)
        F[0:3, 3:6] = -cross_mx(a_int * dt + 0.25 * (C + C1) @ a_true * dt * dt)
        F[0:3, 6:9] = np.eye(3) * dt
        F[0:3, 9:12] = dt * dv + 0.25 * dt * dt * (C @ ax @ cross + C1 @ ax @ cross1)
        F[0:3, 12:15] = -C_int * dt + 0.25 * (C + C1) * dt * dt
        F[3:6, 9:12] = -dt * C1
        F[6:9, 3:6] = -cross_mx(0.5 * (C + C1) @ a_true * dt)
        F[6:9, 9:12] = 0.5 * dt * (C @ ax @ cross + C1 @ ax @ cross1)
        F[6:9, 12:15] = -0.5 * (C + C1) * dt
        P = F @ P @ F.T
        s2v = dt * sa * sa
        P[3:6, 3:6] += np.eye(3) * dt * sg * sg
        P[6:9, 6:9] += np.eye(3) * s2v
        P[0:3, 0:3] += np.eye(3) * 0.5 * dt * dt * s2v
        P[9:12, 9:12] += np.eye(3) * dt * prm.sigma_gw_c ** 2
        P[12:15, 12:15] += np.eye(3) * dt * prm.sigma_aw_c ** 2
        Dq, C_int, a_int, cross, dv, time = Dq1, C_int1, a_int1, cross1, dv1, nexttime
        if nexttime == t1:
            break
    P = 0.5 * (P + P.T)
    info = np.linalg.inv(P)
    info = 0.5 * (info + info.T)
    sqrt_info = np.linalg.cholesky(info).T          # LLT(information).matrixL().transpose()
    return Dq, a_int, a_dint, sqrt_info, Delta_t


def imu_residual(samples, prm, t0, t1, pose0, sb0, pose1, sb1):
    Dq, a_int, a_dint, S, _ = preintegrate(samples, prm, t0, t1, sb0)
    Dt = ns_to_sec(t1 - t0)
    C_S0_W = q2R(pose0[3:] / np.linalg.norm(pose0[3:])).T
    g = np.array([0.0, 0.0, prm.g])
    e = np.zeros(15)
    e[0:3] = C_S0_W @ (pose0[:3] - pose1[:3] + sb0[:3] * Dt - 0.5 * g * Dt * Dt) + a_dint
    q0, q1 = pose0[3:] / np.linalg.norm(pose0[3:]), pose1[3:] / np.linalg.norm(pose1[3:])
    e[3:6] = 2.0 * qmul(Dq, qmul(qinv(q1), q0))[:3]
    e[6:9] = C_S0_W @ (sb0[:3] - sb1[:3] - g * Dt) + a_int
    e[9:15] = sb0[3:9] - sb1[3:9]
    return S @ e


# ---------------------------------------------------------------- the window's residual vector
class NumpyWindow:
    def __init__(self, w):
        self.w = w
        self.K, self.L = len(w.poses), len(w.landmarks)
        self.pose0 = w.poses.copy()
        self.n_par = 6 * self.K + 9 * self.K + 3 * self.L
        o = w.obs
        self.o_pose, self.o_lm, self.o_cam = o["pose_idx"].astype(int), o["lm_idx"].astype(int), o["cam_idx"].astype(int)
        self.R_SC = [q2R(e[3:]) for e in w.extrinsics]
        self.t_SC = [e[:3] for e in w.extrinsics]

    def unpack(self, x):
        K, L = self.K, self.L
        poses = np.stack([pose_plus(self.pose0[k], x[6 * k:6 * k + 6]) for k in range(K)])
        sb = x[6 * K:15 * K].reshape(K, 9)
        lm = self.w.landmarks.copy()
        lm[:, :3] = x[15 * K:].reshape(L, 3)
        return poses, sb, lm

    def x0(self):
        return np.concatenate([np.zeros(6 * self.K), self.w.speed_bias.reshape(-1), self.w.landmarks[:, :3].reshape(-1)])

    def residuals(self, x, parts=False):
        w = self.w
        poses, sb, lm = self.unpack(x)
        R_WS = np.stack([q2R(p[3:]) for p in poses])
        out = []
        # reprojection, per camera
        r_obs = np.zeros((len(w.obs), 2))
        for c in range(len(w.cameras)):
            m = self.o_cam == c
            if not m.any():
                continue
            X = lm[self.o_lm[m]]
            Rw = R_WS[self.o_pose[m]]
            t = poses[self.o_pose[m], :3]
            p_S = np.einsum("nji,nj->ni", Rw, X[:, :3] - t * X[:, 3:4])          # C_SW (X - t w)
            p_C = (p_S - self.t_SC[c] * X[:, 3:4]) @ self.R_SC[c]               # C_CS (p_S - t_SC w)
            sgn = np.where(X[:, 3] < 0, -1.0, 1.0)[:, None]
            px, _ = synthetic.project_points(w.cameras[c], p_C * sgn)
            r_obs[m] = (w.obs["z"][m] - px) * w.obs["sqrt_info"][m][:, None]
        s = np.sum(r_obs * r_obs, axis=1)
        scale = np.where(s > 1e-300, np.sqrt(np.log1p(s) / np.maximum(s, 1e-300)), 1.0)      # CauchyLoss(1): ||r'||^2 = log(1+s)
        out.append((r_obs * scale[:, None]).reshape(-1))
        for T in w.imu_terms:
            smp = w.imu_samples[T["sample_offset"]:T["sample_offset"] + T["sample_count"]]
            out.append(imu_residual(smp, w.imu_params, int(T["t0_ns"]), int(T["t1_ns"]), poses[T["pose0"]], sb[T["sb0"]], poses[T["pose1"]], sb[T["sb1"]]))
        for pr in w.pose_priors:
            T, m = poses[pr["pose_idx"]], pr["meas"]
            dq = qmul(m[3:], qinv(T[3:]))
            e = np.concatenate([m[:3] - T[:3], 2.0 * dq[:3]])
            out.append(pr["sqrt_info"].reshape(6, 6) @ e)
        for pr in w.sb_priors:
            out.append(pr["sqrt_info"].reshape(9, 9) @ (pr["meas"] - sb[pr["sb_idx"]]))
        return out if parts else np.concatenate(out)

    def sparsity(self):
        w, K = self.w, self.K
        n_res = 2 * len(w.obs) + 15 * len(w.imu_terms) + 6 * len(w.pose_priors) + 9 * len(w.sb_priors)
        S = lil_matrix((n_res, self.n_par), dtype=np.int8)
        for i in range(len(w.obs)):
            p, l = self.o_pose[i], self.o_lm[i]
            S[2 * i:2 * i + 2, 6 * p:6 * p + 6] = 1
            S[2 * i:2 * i + 2, 15 * K + 3 * l:15 * K + 3 * l + 3] = 1
        row = 2 * len(w.obs)
        for T in w.imu_terms:
            for p in (T["pose0"], T["pose1"]):
                S[row:row + 15, 6 * p:6 * p + 6] = 1
            for b in (T["sb0"], T["sb1"]):
                S[row:row + 15, 6 * K + 9 * b:6 * K + 9 * b + 9] = 1
            row += 15
        for pr in w.pose_priors:
            S[row:row + 6, 6 * pr["pose_idx"]:6 * pr["pose_idx"] + 6] = 1
            row += 6
        for pr in w.sb_priors:
            S[row:row + 9, 6 * K + 9 * pr["sb_idx"]:6 * K + 9 * pr["sb_idx"] + 9] = 1
            row += 9
        return S


def state_to_x(nw, poses, sb, lm):
    """Minimal coordinates of a state relative to the window's initial poses (inverse of NumpyWindow.unpack)."""
    x = nw.x0().copy()
    for k in range(nw.K):
        dq = qmul(poses[k, 3:], qinv(nw.pose0[k, 3:]))
        nv = np.linalg.norm(dq[:3])
        ang = 2 * np.arctan2(nv, dq[3])
        x[6 * k:6 * k + 3] = poses[k, :3] - nw.pose0[k, :3]
        x[6 * k + 3:6 * k + 6] = dq[:3] / max(nv, 1e-300) * ang
    x[6 * nw.K:15 * nw.K] = sb.reshape(-1)
    x[15 * nw.K:] = lm[:, :3].reshape(-1)
    return x


def family_costs(nw, x):
    parts = nw.residuals(x, parts=True)
    n_imu, n_pp = len(nw.w.imu_terms), len(nw.w.pose_priors)
    c = [0.5 * float(p @ p) for p in parts]
    return dict(reprojection=c[0], imu=sum(c[1:1 + n_imu]), pose_prior=sum(c[1 + n_imu:1 + n_imu + n_pp]), speed_bias_prior=sum(c[1 + n_imu + n_pp:]),
                total=sum(c))


def main():
    """Fixture: cost values of the NumPy objective at a handful of states of two windows.  The states are INPUTS
    (initial guess, random perturbations of it, one state far from it); the costs are the answers the oracle
    (tests/test_golden.py) and the device (tests/test_gpu_golden.py) must reproduce."""
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    fix = {}
    for tag, cfg_id, idx in (("cfg1_w0", 1, 0), ("cfg1_w3", 1, 3)):
        w = synthetic.make_window(cfg_id, idx)
        nw = NumpyWindow(w)
        rng = np.random.Generator(np.random.PCG64(0x0B200 + 77 + idx))
        states = [nw.x0()]
        for scale in (0.2, 1.0, 5.0):
            x = nw.x0().copy()
            x[:6 * nw.K] += scale * np.tile([0.01, 0.01, 0.01, 0.002, 0.002, 0.002], nw.K) * rng.normal(size=6 * nw.K)
            x[6 * nw.K:15 * nw.K] += scale * np.tile([0.02] * 3 + [0.001] * 3 + [0.01] * 3, nw.K) * rng.normal(size=9 * nw.K)
            x[15 * nw.K:] += scale * 0.03 * rng.normal(size=3 * nw.L)
            states.append(x)
        for i, x in enumerate(states):
            poses, sb, lm = nw.unpack(x)
            c = family_costs(nw, x)
            fix["%s_s%d_poses" % (tag, i)] = poses
            fix["%s_s%d_speed_bias" % (tag, i)] = sb
            fix["%s_s%d_landmarks" % (tag, i)] = lm
            fix["%s_s%d_cost" % (tag, i)] = np.array([c["total"], c["reprojection"], c["imu"], c["pose_prior"], c["speed_bias_prior"]])
            print(tag, i, c)
    fix["layout"] = np.array("<window>_s<i>_{poses,speed_bias,landmarks}: state; _cost: [total, reprojection (Cauchy), imu, pose prior, speed/bias prior]; "
                             "windows: synthetic.make_window(1, 0) and (1, 3); generator: tools/make_golden_numpy.py")
    np.savez_compressed(os.path.join(out_dir, "numpy_costs.npz"), **fix)


def golden_window():
    import dataclasses
    return synthetic.make_window(1, 0, cfg=dataclasses.replace(synthetic.CONFIGS[1], outlier_fraction=0.0))


def optimize():
    """Optional (`--optimize`, ~10 minutes): scipy.optimize.least_squares from the initial guess.  With finite-difference
    Jacobians it creeps along the flat valley of far landmarks and does not reach a certified optimum within the budget
    (it passes BELOW the cost at which Ceres' function tolerance stops the dogleg: 421.1 vs 446.2 on cfg-1 window 0), so
    no optimum is committed as a fixture; the committed fixture pins cost VALUES instead (main())."""
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    # cfg-1 without gross outliers: with CauchyLoss the cost of a window that contains outliers is non-convex and different
    # algorithms may settle in different local minima (observed: scipy 421.1 vs the dogleg's 446.2 on the standard
    # cfg-1 window); without outliers the minimum near the initial guess is unique and algorithm independent.
    w = golden_window()
    nw = NumpyWindow(w)
    x0 = nw.x0()
    r0 = nw.residuals(x0)
    print("initial cost %.9g, %d parameters, %d residuals" % (0.5 * r0 @ r0, len(x0), len(r0)))
    sol = least_squares(nw.residuals, x0, jac_sparsity=nw.sparsity(), method="trf", x_scale="jac", ftol=1e-15, xtol=1e-15, gtol=1e-12,
                        max_nfev=300, verbose=1)
    # polish: restart from the solution (a fresh scaling / trust region) until the cost stops moving
    for _ in range(3):
        sol2 = least_squares(nw.residuals, sol.x, jac_sparsity=nw.sparsity(), method="trf", x_scale="jac", ftol=1e-15, xtol=1e-15, gtol=1e-13,
                             max_nfev=100, verbose=0)
        if sol2.cost >= sol.cost * (1 - 1e-13):
            sol = sol2 if sol2.cost < sol.cost else sol
            break
        sol = sol2
    poses, sb, lm = nw.unpack(sol.x)
    print("final cost %.12g (status %d, %d evaluations)" % (sol.cost, sol.status, sol.nfev))
    print("optimality (inf-norm of the scaled gradient) %.3e" % sol.optimality)
    np.savez_compressed(os.path.join("/tmp", "cfg1_numpy_optimum.npz"), poses=poses, speed_bias=sb, landmarks=lm, cost=np.array(sol.cost),
                        initial_cost=np.array(0.5 * r0 @ r0), optimality=np.array(sol.optimality), status=np.array(sol.status),
                        note=np.array("scipy.optimize.least_squares optimum of cfg-1 (window 0) without outliers; see tools/make_golden_numpy.py"))


if __name__ == "__main__":
    optimize() if "--optimize" in sys.argv else main()
