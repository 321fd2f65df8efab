"""GPU parity of the single-residual-block hooks (the device functions the solver uses) against the
CPU oracle, through the C-ABI.  Mirrors ErrorInterface::EvaluateWithMinimalJacobians call sites."""
import numpy as np
import pytest

from okvis_b200 import abi, synthetic
from test_oracle_functors import CAMS, make_test_cam, rand_pose, _imu_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 1)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(CAMS))
def test_reprojection_hook(ctx, oracle, name):
    cam = make_test_cam(name)
    rng = np.random.default_rng(31)
    n = 2000
    pose = np.stack([rand_pose(rng, 1.0, 0.5) for _ in range(n)])
    ext = np.stack([rand_pose(rng, 0.1, 0.2) for _ in range(n)])
    hp = np.zeros((n, 4))
    for i in range(n):
        near = i % 17 == 0          # exercises the "invalid point" branch (z/w < 0.2)
        p_C = np.array([rng.uniform(-1, 1), rng.uniform(-0.7, 0.7), rng.uniform(0.05 if near else 0.5, 8.0)])
        R_SC, R_WS = synthetic.R_from_quat(ext[i, 3:]), synthetic.R_from_quat(pose[i, 3:])
        p_W = R_WS @ (R_SC @ p_C + ext[i, :3]) + pose[i, :3]
        w = rng.uniform(0.2, 1.5) * (-1 if i % 11 == 0 else 1)   # w < 0 quirk
        hp[i] = np.concatenate([p_W * w, [w]])
    z = rng.uniform([0, 0], [752, 480], (n, 2))
    sq = rng.uniform(0.5, 2.0, n)
    got = ctx.eval_reprojection(cam, pose, hp, ext, z, sq)
    ref = oracle.eval_reprojection(cam, pose, hp, ext, z, sq)
    for g, r in zip(got, ref):
        scale = np.maximum(1.0, np.abs(r).reshape(n, -1).max(1)).reshape((n,) + (1,) * (r.ndim - 1))
        assert np.abs((g - r) / scale).max() < 1e-10


def test_imu_hooks(ctx, oracle):
    rng = np.random.default_rng(32)
    for case in range(4):
        prm, s, t0, t1 = _imu_case(rng)
        if case == 3:
            s["acc"][5:9, 2] = 200.0     # accelerometer saturation
        pose0 = rand_pose(rng, 1.0, 0.5)
        sb0 = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3)])
        n0, pose1, sb1, P0, F0 = oracle.imu_propagate(prm, s, t0, t1, pose0, sb0)
        n1, p1, s1, P1, F1 = ctx.imu_propagate(prm, s, t0, t1, pose0, sb0)
        assert n0 == n1
        assert np.abs(p1 - pose1).max() < 1e-12 and np.abs(s1 - sb1).max() < 1e-12
        assert np.abs(P1 - P0).max() < 1e-11 * np.abs(P0).max()
        assert np.abs(F1 - F0).max() < 1e-11 * max(1.0, np.abs(F0).max())
        pose1 = oracle.pose_plus(pose1, np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.005, 3)]))
        sb1 = sb1 + rng.normal(0, 0.01, 9)
        for ref_pt in (None, sb0 + np.concatenate([np.zeros(3), rng.normal(0, 1e-5, 3), rng.normal(0, 1e-3, 3)]),
                       sb0 + np.concatenate([np.zeros(3), [0.01, 0, 0], np.zeros(3)])):
            r0, J0, sq0, redo0 = oracle.eval_imu(prm, s, t0, t1, pose0, sb0, pose1, sb1, sb_ref=ref_pt)
            r1, J1, sq1, redo1 = ctx.eval_imu(prm, s, t0, t1, pose0, sb0, pose1, sb1, sb_ref=ref_pt)
            assert redo0 == redo1
            i0, i1 = sq0.T @ sq0, sq1.T @ sq1
            assert np.abs(i1 - i0).max() < 1e-7 * np.abs(i0).max()
            assert np.abs(r1 - r0).max() < 1e-6 * max(1.0, np.abs(r0).max())
            for a, b in zip(J1, J0):
                assert np.abs(a - b).max() < 1e-6 * np.abs(b).max()
            assert abs(r1 @ r1 - r0 @ r0) < 1e-7 * (r0 @ r0)


def test_prior_hooks(ctx, oracle):
    rng = np.random.default_rng(33)
    for _ in range(10):
        meas = rand_pose(rng)
        pose = oracle.pose_plus(meas, np.concatenate([rng.normal(0, 0.1, 3), rng.normal(0, 0.05, 3)]))
        A = rng.normal(0, 1, (6, 6))
        S, _ = oracle.sqrt_information(A @ A.T + 6 * np.eye(6))
        r0, J0 = oracle.eval_pose_error(meas, S, pose)
        r1, J1 = ctx.eval_pose_error(meas, S, pose)
        assert np.abs(r1 - r0).max() < 1e-11 and np.abs(J1 - J0).max() < 1e-11
        r0, Ja, Jb = oracle.eval_relative_pose(S, meas, pose)
        r1, Jc, Jd = ctx.eval_relative_pose(S, meas, pose)
        assert np.abs(r1 - r0).max() < 1e-11 and np.abs(Jc - Ja).max() < 1e-11 and np.abs(Jd - Jb).max() < 1e-11
        m9, x9 = rng.normal(0, 1, 9), rng.normal(0, 1, 9)
        S9 = np.diag([1, 1, 1] + [1 / 0.03] * 3 + [1 / 0.1] * 3).astype(np.float64)
        r0, J0 = oracle.eval_speed_bias_error(m9, S9, x9)
        r1, J1 = ctx.eval_speed_bias_error(m9, S9, x9)
        assert np.abs(r1 - r0).max() < 1e-12 and np.abs(J1 - J0).max() == 0


def test_marginalization_hook(ctx, oracle):
    rng = np.random.default_rng(34)
    w = synthetic.make_window(1, 0)
    marg = synthetic.make_random_marg_prior(w, rng)
    blocks, off = [], 0
    for k in marg["block_kind"]:
        dim = 9 if k == abi.BLOCK_SPEED_BIAS else 7
        b = marg["x0"][off:off + dim].copy()
        b = oracle.pose_plus(b, np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.02, 3)])) if dim == 7 \
            else b + rng.normal(0, 0.05, 9)
        blocks.append(b)
        off += dim
    x = np.concatenate(blocks)
    r0, J0 = oracle.eval_marginalization(marg, x)
    r1, J1 = ctx.eval_marginalization(marg, x)
    assert np.abs(r1 - r0).max() < 1e-10 * max(1.0, np.abs(r0).max())
    assert np.abs(J1 - J0).max() < 1e-10 * np.abs(J0).max()
