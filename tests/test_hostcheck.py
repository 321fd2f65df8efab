"""GPU-less cross-check: the product's per-thread math (okvis_b200/csrc/okb_math.cuh, okb_imu.cuh),
compiled for the host, against the CPU oracle.  The same device functions are exercised on the GPU
through the C-ABI hooks in tests/test_gpu_functors.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from okvis_b200 import abi, synthetic
from test_oracle_functors import CAMS, make_test_cam, rand_pose, _imu_case

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "libhostcheck.so")
    deps = [src, os.path.join(HERE, "..", "okvis_b200", "csrc", "okb_math.cuh"),
            os.path.join(HERE, "..", "okvis_b200", "csrc", "okb_imu.cuh"),
            os.path.join(HERE, "..", "okvis_b200", "csrc", "okb_hostpack.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-x", "c++", "-shared", "-o", so, src])
    return C.CDLL(so)


def p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("name", list(CAMS))
def test_reprojection_matches_oracle(hc, oracle, name):
    cam = make_test_cam(name)
    cam_arr = np.array([cam], dtype=abi.camera_dtype)
    rng = np.random.default_rng(21)
    for it in range(200):
        T_WS, T_SC = rand_pose(rng, 1.0, 0.5), rand_pose(rng, 0.1, 0.2)
        p_C = np.array([rng.uniform(-1.2, 1.2), rng.uniform(-0.8, 0.8), rng.uniform(0.05 if it % 10 == 0 else 0.5, 8.0)])
        R_SC, R_WS = synthetic.R_from_quat(T_SC[3:]), synthetic.R_from_quat(T_WS[3:])
        p_W = R_WS @ (R_SC @ p_C + T_SC[:3]) + T_WS[:3]
        w = rng.uniform(0.2, 1.5) * (-1 if rng.random() < 0.2 else 1)
        hp = np.concatenate([p_W * w, [w]])
        z = rng.uniform([0, 0], [752, 480])
        sq = rng.uniform(0.5, 2.0)
        r0, a0, a1, a2 = oracle.eval_reprojection(cam, T_WS[None], hp[None], T_SC[None], z[None], np.array([sq]))
        r, J0, J1, J2 = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 3)), np.zeros((2, 6))
        hc.hc_reproj_full(p(cam_arr), p(T_WS), p(hp), p(T_SC), p(z), C.c_double(sq), p(r), p(J0), p(J1), p(J2))
        scale = max(1.0, np.abs(a0).max())
        assert np.abs(r - r0[0]).max() < 1e-9 * max(1, np.abs(r0).max())
        assert np.abs(J0 - a0[0]).max() < 1e-9 * scale
        assert np.abs(J1 - a1[0]).max() < 1e-9 * scale
        assert np.abs(J2 - a2[0]).max() < 1e-9 * scale


def test_pose_functors_match_oracle(hc, oracle):
    rng = np.random.default_rng(22)
    for _ in range(50):
        meas = rand_pose(rng)
        pose = oracle.pose_plus(meas, np.concatenate([rng.normal(0, 0.1, 3), rng.normal(0, 0.05, 3)]))
        A = rng.normal(0, 1, (6, 6))
        S, _ = oracle.sqrt_information(A @ A.T + 6 * np.eye(6))
        r0, J0 = oracle.eval_pose_error(meas, S, pose)
        r, J = np.zeros(6), np.zeros((6, 6))
        hc.hc_pose_error(p(meas), p(S), p(pose), p(r), p(J))
        assert np.abs(r - r0).max() < 1e-11 and np.abs(J - J0).max() < 1e-11
        r0, Ja, Jb = oracle.eval_relative_pose(S, meas, pose)
        r, J0_, J1_ = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
        hc.hc_relative_pose(p(S), p(meas), p(pose), p(r), p(J0_), p(J1_))
        assert np.abs(r - r0).max() < 1e-11 and np.abs(J0_ - Ja).max() < 1e-11 and np.abs(J1_ - Jb).max() < 1e-11
        d = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.3, 3)])
        o = np.zeros(7)
        hc.hc_pose_plus(p(meas), p(d), p(o))
        assert np.abs(o - oracle.pose_plus(meas, d)).max() < 1e-14
        dm = np.zeros(6)
        hc.hc_pose_minus(p(meas), p(pose), p(dm))
        assert np.abs(dm - oracle.pose_minus(meas, pose)).max() < 1e-14
        # marginalisation rotation block = (lift(x0) * plus(x))[3:6,3:6]
        B = np.zeros((3, 3))
        hc.hc_marg_rot_block(p(meas), p(pose), p(B))
        LP = oracle.pose_lift_jacobian(meas) @ oracle.pose_plus_jacobian(pose)
        assert np.abs(B - LP[3:, 3:]).max() < 1e-13


def test_eig3(hc):
    rng = np.random.default_rng(23)
    for _ in range(100):
        A = rng.normal(0, 1, (3, 3)) * rng.uniform(0.01, 100)
        S = A @ A.T
        s6 = np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])
        ev = np.zeros(3)
        hc.hc_eig3(p(s6), p(ev))
        assert np.allclose(ev, np.linalg.eigvalsh(S), rtol=1e-10, atol=1e-12 * ev[2])


def _hc_imu_eval(hc, prm, s, t0, t1, pose0, sb0, pose1, sb1, sb_ref=None):
    r = np.zeros(15)
    J = [np.zeros((15, 6)), np.zeros((15, 9)), np.zeros((15, 6)), np.zeros((15, 9))]
    sq = np.zeros((15, 15))
    hc.hc_imu_eval.restype = C.c_int
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pose0, sb0, pose1, sb1)]
    ref = np.ascontiguousarray(sb_ref, dtype=np.float64) if sb_ref is not None else None
    redo = hc.hc_imu_eval(C.byref(prm), p(s), C.c_int(len(s)), C.c_int64(int(t0)), C.c_int64(int(t1)),
                          *[p(x) for x in a], p(ref) if ref is not None else None, p(r), *[p(j) for j in J], p(sq))
    return r, J, sq, redo


def test_imu_matches_oracle(hc, oracle):
    rng = np.random.default_rng(24)
    for case in range(5):
        prm, s, t0, t1 = _imu_case(rng)
        if case == 4:
            s["gyro"][10:20, 1] = 9.0   # saturation branch
        pose0 = rand_pose(rng, 1.0, 0.5)
        sb0 = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3)])
        n, pose1, sb1, P, F = oracle.imu_propagate(prm, s, t0, t1, pose0, sb0)
        # propagation parity
        pp, ss = pose0.copy(), sb0.copy()
        cov, jac = np.zeros((15, 15)), np.zeros((15, 15))
        hc.hc_imu_propagate.restype = C.c_int
        n2 = hc.hc_imu_propagate(C.byref(prm), p(s), C.c_int(len(s)), C.c_int64(t0), C.c_int64(t1), p(pp), p(ss),
                                 p(cov), p(jac))
        assert n2 == n
        assert np.abs(pp - pose1).max() < 1e-12 and np.abs(ss - sb1).max() < 1e-12
        assert np.abs(cov - P).max() < 1e-12 * np.abs(P).max() + 1e-20
        assert np.abs(jac - F).max() < 1e-12 * max(1, np.abs(F).max())
        # residual + Jacobians, fresh functor and with a previous linearisation point
        pose1 = oracle.pose_plus(pose1, np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.005, 3)]))
        sb1 = sb1 + rng.normal(0, 0.01, 9)
        for ref in (None, sb0 + np.concatenate([np.zeros(3), rng.normal(0, 1e-5, 3), rng.normal(0, 1e-3, 3)])):
            r0, J0, sq0, redo0 = oracle.eval_imu(prm, s, t0, t1, pose0, sb0, pose1, sb1, sb_ref=ref)
            r, J, sq, redo = _hc_imu_eval(hc, prm, s, t0, t1, pose0, sb0, pose1, sb1, sb_ref=ref)
            assert redo == redo0
            # information = S^T S must agree (the sqrt factors are both upper Cholesky -> equal)
            info0, info = sq0.T @ sq0, sq.T @ sq
            assert np.abs(info - info0).max() < 1e-7 * np.abs(info0).max()
            assert np.abs(r - r0).max() < 1e-6 * max(1.0, np.abs(r0).max())
            for a, b in zip(J, J0):
                assert np.abs(a - b).max() < 1e-6 * np.abs(b).max()
            # cost parity (what the solver sees)
            assert abs(r @ r - r0 @ r0) < 1e-7 * (r0 @ r0)


def test_landmark_sort_by_frame_range(hc):
    """okb_window_upload's internal landmark order (okb_hostpack.hpp): a stable permutation sorted by
    (first, last) observing frame, unobserved landmarks last, consistent inverse map and per-tile frame ranges."""
    rng = np.random.default_rng(5)
    for L in (1, 31, 32, 33, 300, 2000):
        K = int(rng.integers(2, 21))
        vis = np.zeros(L, np.uint32)
        for l in range(L):
            kind = rng.integers(0, 10)
            if kind == 0:
                continue                                            # unobserved
            a = int(rng.integers(0, K)); b = int(rng.integers(a, K))
            m = ((1 << (b + 1)) - 1) ^ ((1 << a) - 1)                # run [a, b]
            if kind == 1:
                m &= int(rng.integers(1, 1 << 20)) | (1 << a) | (1 << b)   # gaps inside the run
            vis[l] = m
        perm, inv = np.zeros(L, np.uint32), np.zeros(L, np.uint32)
        n_tiles = (L + 31) // 32
        tr = np.zeros(n_tiles, np.uint32)
        hc.hc_sort_landmarks(p(vis), L, p(perm), p(inv), p(tr))
        assert sorted(perm.tolist()) == list(range(L))
        assert np.array_equal(inv[perm], np.arange(L))

        def key(m):
            m = int(m)
            if m == 0:
                return 32 * 32
            return ((m & -m).bit_length() - 1) * 32 + (m.bit_length() - 1)
        keys = [key(vis[l]) for l in perm]
        assert keys == sorted(keys)
        for k in set(keys):                                         # stable inside a key
            idx = [int(perm[j]) for j in range(L) if keys[j] == k]
            assert idx == sorted(idx)
        for t in range(n_tiles):
            ms = [int(vis[perm[j]]) for j in range(32 * t, min(L, 32 * t + 32)) if vis[perm[j]]]
            a, b = int(tr[t]) & 0xff, int(tr[t]) >> 8
            if not ms:
                assert a > b
            else:
                assert a == min((m & -m).bit_length() - 1 for m in ms) and b == max(m.bit_length() - 1 for m in ms)
