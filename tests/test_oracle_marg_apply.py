"""CPU: the oracle's marginalisation step (oracle/oracle_marg_apply.hpp) against an independent numpy assembly of the
same linear system from the single-functor hooks + a plain pseudo-inverse Schur complement.  Pins the bookkeeping of
the restatement (ordering, signs of b0, Cauchy correction, first-estimate points); unique quantities only."""
import numpy as np

from okvis_b200 import abi, synthetic
from oracle import oracle_py as op

P, SB = abi.BLOCK_POSE, abi.BLOCK_SPEED_BIAS


def numpy_marginalize(w, job_blocks, marg_flags, imu_terms, sb_priors, lms):
    dims = [6 if k == P else 9 for k, _ in job_blocks]
    col = np.concatenate([[0], np.cumsum(dims)])
    Nd = int(col[-1])
    N = Nd + 3 * len(lms)
    H, b = np.zeros((N, N)), np.zeros(N)
    blk = {kb: i for i, kb in enumerate(job_blocks)}

    def add(offs, Js, r):
        for oi, Ji in zip(offs, Js):
            b[oi:oi + Ji.shape[1]] -= Ji.T @ r
            for oj, Jj in zip(offs, Js):
                H[oi:oi + Ji.shape[1], oj:oj + Jj.shape[1]] += Ji.T @ Jj

    for i in sb_priors:
        pr = w.sb_priors[i]
        S = pr["sqrt_info"].reshape(9, 9)
        r = S @ (pr["meas"] - w.speed_bias[pr["sb_idx"]])
        add([col[blk[(SB, int(pr["sb_idx"]))]]], [-S], r)
    for t in imu_terms:
        T = w.imu_terms[t]
        smp = w.imu_samples[T["sample_offset"]:T["sample_offset"] + T["sample_count"]]
        r, J, _, _ = op.eval_imu(w.imu_params, smp, T["t0_ns"], T["t1_ns"], w.poses[T["pose0"]], w.speed_bias[T["sb0"]], w.poses[T["pose1"]],
                                 w.speed_bias[T["sb1"]])
        offs = [col[blk[(P, int(T["pose0"]))]], col[blk[(SB, int(T["sb0"]))]], col[blk[(P, int(T["pose1"]))]], col[blk[(SB, int(T["sb1"]))]]]
        add(offs, J, r)
    for j, l in enumerate(lms):
        for ob in w.obs[w.obs["lm_idx"] == l]:
            r, J0, J1, _ = op.eval_reprojection(w.cameras[ob["cam_idx"]], w.poses[ob["pose_idx"]][None], w.landmarks[l][None],
                                                w.extrinsics[ob["ext_idx"]][None], ob["z"][None], np.array([ob["sqrt_info"]]))
            s = np.sqrt(1.0 / (1.0 + r[0] @ r[0]))
            add([col[blk[(P, int(ob["pose_idx"]))]], Nd + 3 * j], [s * J0[0], s * J1[0]], s * r[0])
    # stage 1: landmark blocks (block-diagonal V, one 3x3 pseudo-inverse each); stage 2: the dense blocks
    Hd, bd = H[:Nd, :Nd].copy(), b[:Nd].copy()
    for j in range(len(lms)):
        sl = slice(Nd + 3 * j, Nd + 3 * j + 3)
        Vi = np.linalg.pinv(H[sl, sl], hermitian=True)
        Hd -= H[:Nd, sl] @ Vi @ H[sl, :Nd]
        bd -= H[:Nd, sl] @ Vi @ b[sl]
    keep = [i for i in range(Nd) if not marg_flags[np.searchsorted(col, i, side="right") - 1]]
    gone = [i for i in range(Nd) if i not in keep]
    V = Hd[np.ix_(gone, gone)]
    pv = np.sqrt(np.diag(V))
    Vi = np.linalg.pinv(0.5 * (V + V.T) / np.outer(pv, pv), hermitian=True) / np.outer(pv, pv)
    Wm = Hd[np.ix_(keep, gone)]
    return Hd[np.ix_(keep, keep)] - Wm @ Vi @ Wm.T, bd[keep] - Wm @ Vi @ bd[gone]


def test_oracle_marginalisation_step_equals_numpy_schur():
    w = synthetic.make_window(1, 0, cfg=synthetic.WindowConfig(n_frames=5, n_cams=1, n_landmarks=300, frame_dt=0.125, perturb=False))
    L = len(w.landmarks)
    last, first = np.zeros(L, int), np.full(L, 99)
    np.maximum.at(last, w.obs["lm_idx"], w.obs["pose_idx"])
    np.minimum.at(first, w.obs["lm_idx"], w.obs["pose_idx"])
    cnt = np.bincount(w.obs["lm_idx"], minlength=L)
    lms = np.nonzero((first == 0) & (last <= 2) & (cnt >= 3))[0].astype(np.uint32)[:40]
    blocks = [(P, 0), (SB, 0), (P, 1), (SB, 1), (P, 2)]
    flags = [1, 1, 0, 0, 0]
    job = abi.make_marg_job([k for k, _ in blocks], [i for _, i in blocks], [-1] * 5, flags, imu_terms=[0], sb_priors=[0], landmarks=lms)
    pb = op.OracleProblem(w)
    o = pb.marginalize(job)
    Hn, bn = numpy_marginalize(w, blocks, flags, [0], [0], lms)
    assert o["n"] == 21
    assert np.abs(o["H"] - Hn).max() <= 1e-7 * np.abs(Hn).max()
    assert np.abs(o["b0"] - bn).max() <= 1e-7 * max(1.0, np.abs(bn).max())
    JtJ = o["J"].T @ o["J"]
    assert np.abs(JtJ - 0.5 * (o["H"] + o["H"].T)).max() <= 1e-8 * np.abs(o["H"]).max()
    # e0 = -pinv(J^T) b0  <=>  J^T e0 = -b0 on the range of the preconditioned H (H is rank deficient: the gauge
    # freedom of the window is not fixed by the terms linearised here)
    pd = np.sqrt(np.diag(o["H"]))
    lam, U = np.linalg.eigh(0.5 * (o["H"] + o["H"].T) / np.outer(pd, pd))
    Ur = U[:, lam > np.finfo(float).eps * len(lam) * lam.max()]
    assert Ur.shape[1] == o["rank"]
    assert np.abs(Ur.T @ ((o["J"].T @ o["e0"] + o["b0"]) / pd)).max() <= 1e-6 * max(1.0, np.abs(o["b0"] / pd).max())
    assert np.array_equal(o["x0"][:7], w.poses[1]) and np.array_equal(o["x0"][7:16], w.speed_bias[1])
