"""Device-side marginalisation (okb_window_marginalize, okb_marg.cuh) against the CPU oracle's restatement of
MarginalizationError::addResidualBlock / marginalizeOut / updateErrorComputation
(okvis_ceres/src/MarginalizationError.cpp:127-435, 507-846) on the same window at the same estimates, and row M
(MarginalizationError::EvaluateWithMinimalJacobians, :893-946) on a prior the marginalisation PRODUCED.

Eigenvector bases are not unique, so J and e0 are compared through the invariants J^T J, J^T e0, rank, and through
the reduced system (H, b0) itself (see oracle/oracle_marg.hpp)."""
import dataclasses

import numpy as np
import pytest

from okvis_b200 import abi, synthetic

pytestmark = pytest.mark.gpu
P, SB = abi.BLOCK_POSE, abi.BLOCK_SPEED_BIAS


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 4)
    yield c
    c.close()


def oracle_like_device(oracle, ctx, win, w, est):
    """Oracle problem at the device's estimates whose ImuError caches are in the device's state (the functor cache is
    mutable state of the reference, ImuError.hpp:251-276: preintegrated at some earlier bias, corrected to first order)."""
    ref = oracle.OracleProblem(at_estimates(w, est))
    for t in range(len(w.imu_terms)):
        sb_ref, valid, _ = ctx.debug_imu_cache(win, t)
        ref.set_imu_cache(t, sb_ref, valid)
    return ref


def at_estimates(w, est):
    return dataclasses.replace(w, poses=est["poses"].copy(), speed_bias=est["speed_bias"].copy(), landmarks=est["landmarks"].copy())


def drop_oldest(w, lms_gone, marg, new_first_pose):
    """Window after the bookkeeping half of applyMarginalizationStrategy for one removed frame: frame 0 and the
    marginalised landmarks are gone, later frames move down, the first pose is re-fixed (Estimator.cpp:761-770)."""
    K = len(w.poses)
    keep_obs = (w.obs["pose_idx"] != 0) & ~np.isin(w.obs["lm_idx"], lms_gone)
    obs = w.obs[keep_obs].copy()
    obs["pose_idx"] -= 1
    terms = w.imu_terms[1:].copy()
    for k in ("pose0", "sb0", "pose1", "sb1"):
        terms[k] -= 1
    lm = w.landmarks.copy()
    lm[lms_gone] = 0.0
    pp = np.zeros(1, abi.pose_prior_dtype)
    pp["pose_idx"], pp["meas"] = 0, new_first_pose
    pp["sqrt_info"] = np.diag([1e7, 1e7, 1e7, 0, 0, 1e7]).reshape(-1)
    return dataclasses.replace(w, poses=np.ascontiguousarray(w.poses[1:]), speed_bias=np.ascontiguousarray(w.speed_bias[1:]),
                               landmarks=lm, obs=np.ascontiguousarray(obs), imu_terms=terms, pose_priors=pp,
                               sb_priors=np.zeros(0, abi.sb_prior_dtype), marg=marg), pp


def well_conditioned(oracle, w, lms, limit=1e7):
    """Landmarks whose robustified, preconditioned 3x3 block V has a condition number below `limit` at the window's
    current values.  An outlier observation (Cauchy weight ~1e-5) can leave V numerically rank deficient; its
    pseudo-inverse then amplifies rounding noise by 1/lambda_min in ANY implementation (the reference included), so such
    blocks cannot be compared between two implementations and are left out of the parity jobs."""
    keep = []
    for l in lms:
        V = np.zeros((3, 3))
        for ob in w.obs[w.obs["lm_idx"] == l]:
            r, _, J1, _ = oracle.eval_reprojection(w.cameras[ob["cam_idx"]], w.poses[ob["pose_idx"]][None], w.landmarks[l][None],
                                                   w.extrinsics[ob["ext_idx"]][None], ob["z"][None], np.array([ob["sqrt_info"]]))
            V += J1[0].T @ J1[0] / (1.0 + r[0] @ r[0])
        d = np.sqrt(np.diag(V))
        if d.min() <= 0:
            continue
        ev = np.linalg.eigvalsh(V / np.outer(d, d))
        if ev[0] > 0 and ev[-1] / ev[0] < limit:
            keep.append(l)
    return np.array(keep, np.uint32)


def first_job(oracle, w):
    """Marginalise frame 0 (pose + speed/bias): its ImuError and SpeedAndBiasError terms, and the landmarks seen in
    frame 0 whose track ended before frame 3 (so that poses 0..2 are connected).  `w` holds the current estimates."""
    L = len(w.landmarks)
    last = np.zeros(L, int)
    first = np.full(L, 99)
    np.maximum.at(last, w.obs["lm_idx"], w.obs["pose_idx"])
    np.minimum.at(first, w.obs["lm_idx"], w.obs["pose_idx"])
    cand = np.nonzero((first == 0) & (last <= 2))[0].astype(np.uint32)
    lms = well_conditioned(oracle, w, cand)
    job = abi.make_marg_job([P, SB, P, SB, P], [0, 0, 1, 1, 2], [-1] * 5, [1, 1, 0, 0, 0], imu_terms=[0], sb_priors=[0], landmarks=lms)
    return job, lms, cand      # cand \\ lms: dropped without being linearised (the "justDelete" path of Estimator.cpp:709-714)


def compare_prior(g, o, tol=2e-9):
    assert g["status"][0] == 0
    assert g["n"] == o["n"]
    assert np.array_equal(g["block_kind"], o["block_kind"]) and np.array_equal(g["block_idx"], o["block_idx"])
    assert np.array_equal(g["x0"], o["x0"])
    sH = np.abs(o["H"]).max()
    assert np.abs(g["H"] - o["H"]).max() <= tol * sH
    assert np.abs(g["b0"] - o["b0"]).max() <= tol * max(1.0, np.abs(o["b0"]).max())
    # eigenvalues of the gauge directions are rounding noise around the threshold eps * n * lambda_max: the rank may
    # differ by the number of such directions; J^T J and J^T e0 are insensitive to them
    assert abs(int(g["status"][1]) - o["rank"]) <= 3
    JtJ_g, JtJ_o = g["J"].T @ g["J"], o["J"].T @ o["J"]
    assert np.abs(JtJ_g - JtJ_o).max() <= 1e-8 * np.abs(JtJ_o).max()
    assert np.abs(g["J"].T @ g["e0"] - o["J"].T @ o["e0"]).max() <= 1e-7 * max(1.0, np.abs(o["J"].T @ o["e0"]).max())
    # J^T J reproduces H on its range: the prior is a faithful factorisation
    assert np.abs(JtJ_g - 0.5 * (g["H"] + g["H"].T)).max() <= 1e-8 * sH


def test_marginalize_oldest_frame_matches_oracle_and_drives_row_M(ctx, oracle):
    w = synthetic.make_window(1, 0)
    K, L = len(w.poses), len(w.landmarks)
    ctx.reserve(0, K, L, len(w.obs), len(w.imu_samples), 80)
    ctx.upload(0, w)
    ctx.optimize(0, 1, max_iterations=6)
    est = ctx.download(0)
    job, lms, cand = first_job(oracle, at_estimates(w, est))
    assert len(lms) > 5
    ref = oracle_like_device(oracle, ctx, 0, w, est)
    ctx.marginalize(0, job)
    g = ctx.download_marg(0)
    o = ref.marginalize(job)
    compare_prior(g, o)
    assert g["n"] == 21 and 15 <= g["status"][2] <= 15        # pose 1, sb 1, pose 2 kept; the 15 marginalised dense dims are full rank
    # ---- bookkeeping half, then optimize with the produced prior: GPU vs oracle (row M on a REAL prior)
    ctx.remove_landmarks(0, cand)
    ctx.remove_frame(0, 0, 0)
    marg_o = dict(block_kind=o["block_kind"], block_idx=(o["block_idx"] - 1).astype(np.uint32), x0=o["x0"], J=np.ascontiguousarray(o["J"]),
                  e0=o["e0"])
    w2, pp = drop_oldest(at_estimates(w, est), cand, marg_o, est["poses"][1])
    ctx.set_priors(0, pp, np.zeros(0, abi.sb_prior_dtype))
    s = ctx.optimize(0, 1, max_iterations=8)[0]
    got = ctx.download(0, dims=(K - 1, K - 1, L))
    ref2 = oracle.OracleProblem(w2)
    so = ref2.solve(8, 1)
    st = ref2.state()
    assert s["iterations"] == so["iterations"] and s["termination"] == so["termination"]
    assert abs(s["initial_cost"] - so["initial_cost"]) <= 1e-7 * so["initial_cost"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
    assert np.abs(got["poses"] - st["poses"]).max() < 1e-6
    assert np.abs(got["speed_bias"] - st["speed_bias"]).max() < 1e-6


def test_second_marginalisation_reuses_the_persisted_system(ctx, oracle):
    """Two marginalisations in a row: the second starts from the H / b0 the first left on the device
    (MarginalizationError keeps H_ / b0_ between calls) and keeps the first-estimate linearisation points."""
    w = synthetic.make_window(1, 3)
    K, L = len(w.poses), len(w.landmarks)
    ctx.reserve(1, K, L, len(w.obs), len(w.imu_samples), 80)
    ctx.upload(1, w)
    ctx.optimize(1, 1, max_iterations=5)
    est = ctx.download(1)
    job, lms, cand = first_job(oracle, at_estimates(w, est))
    ref1 = oracle_like_device(oracle, ctx, 1, w, est)
    ctx.marginalize(1, job)
    g1 = ctx.download_marg(1)
    o1 = ref1.marginalize(job)
    compare_prior(g1, o1)
    ctx.remove_landmarks(1, cand)
    ctx.remove_frame(1, 0, 0)
    marg_o = dict(block_kind=o1["block_kind"], block_idx=(o1["block_idx"] - 1).astype(np.uint32), x0=o1["x0"], J=np.ascontiguousarray(o1["J"]), e0=o1["e0"])
    w2, pp = drop_oldest(at_estimates(w, est), cand, marg_o, est["poses"][1])
    ctx.set_priors(1, pp, np.zeros(0, abi.sb_prior_dtype))
    ctx.optimize(1, 1, max_iterations=4)
    est2 = ctx.download(1, dims=(K - 1, K - 1, L))
    # second step on the 4-frame window: frame 0 again (prior blocks: pose 0, sb 0, pose 1)
    obs2 = w2.obs
    last = np.zeros(L, int)
    first = np.full(L, 99)
    np.maximum.at(last, obs2["lm_idx"], obs2["pose_idx"])
    np.minimum.at(first, obs2["lm_idx"], obs2["pose_idx"])
    lms2 = well_conditioned(oracle, at_estimates(w2, est2), np.nonzero((first == 0) & (last <= 2))[0])
    job2 = abi.make_marg_job([P, SB, P, SB, P], [0, 0, 1, 1, 2], [0, 1, 2, -1, -1], [1, 1, 0, 0, 0], imu_terms=[0], sb_priors=[], landmarks=lms2)
    ref2 = oracle_like_device(oracle, ctx, 1, w2, est2)
    ctx.marginalize(1, job2)
    g2 = ctx.download_marg(1)
    o2 = ref2.marginalize(job2, H_prev=o1["H"], b0_prev=o1["b0"])
    compare_prior(g2, o2, tol=5e-9)
    # linearisation point of the carried-over pose (old pose 1 -> new pose 0 ... now block 0 of the new prior is old block 2)
    assert np.array_equal(g2["x0"][:7], g1["x0"][16:23])


def test_speed_bias_only_marginalisation_and_remove_speed_bias(ctx, oracle):
    """The removeAllButPose frames (Estimator.cpp:483-554): speed/bias 0 is marginalised, pose 0 stays a keyframe pose."""
    w = synthetic.make_window(1, 1)
    K, L = len(w.poses), len(w.landmarks)
    ctx.reserve(2, K, L, len(w.obs), len(w.imu_samples), 80)
    ctx.upload(2, w)
    ctx.optimize(2, 1, max_iterations=5)
    est = ctx.download(2)
    job = abi.make_marg_job([P, SB, P, SB], [0, 0, 1, 1], [-1] * 4, [0, 1, 0, 0], imu_terms=[0], sb_priors=[0])
    ref0 = oracle_like_device(oracle, ctx, 2, w, est)
    ctx.marginalize(2, job)
    g = ctx.download_marg(2)
    o = ref0.marginalize(job)
    compare_prior(g, o)
    assert g["n"] == 21
    ctx.remove_speed_bias(2, 0)
    s = ctx.optimize(2, 1, max_iterations=6)[0]
    # oracle on the same graph: speed/bias blocks 1..K-1, IMU terms 1.., prior over (pose 0, pose 1, sb 0 [old 1])
    terms = w.imu_terms[1:].copy()
    terms["sb0"] -= 1
    terms["sb1"] -= 1
    idx = o["block_idx"].copy()
    idx[o["block_kind"] == SB] -= 1
    marg_o = dict(block_kind=o["block_kind"], block_idx=idx.astype(np.uint32), x0=o["x0"], J=np.ascontiguousarray(o["J"]), e0=o["e0"])
    we = at_estimates(w, est)
    w2 = dataclasses.replace(we, speed_bias=np.ascontiguousarray(we.speed_bias[1:]), imu_terms=terms, sb_priors=np.zeros(0, abi.sb_prior_dtype), marg=marg_o)
    ref = oracle.OracleProblem(w2)
    so = ref.solve(6, 1)
    assert s["iterations"] == so["iterations"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
    got = ctx.download(2, dims=(K, K - 1, L))
    assert np.abs(got["poses"] - ref.state()["poses"]).max() < 1e-6


def test_marginalize_rejects_a_job_that_misses_a_connected_block(ctx, okb):
    w = synthetic.make_window(1, 2)
    ctx.reserve(3, len(w.poses), len(w.landmarks), len(w.obs), len(w.imu_samples), 80)
    ctx.upload(3, w)
    ctx.optimize(3, 1, max_iterations=2)
    job = abi.make_marg_job([P, SB, P], [0, 0, 1], [-1] * 3, [1, 1, 0], imu_terms=[0])      # sb 1 of the IMU term is missing
    ctx.marginalize(3, job)
    assert ctx.download_marg(3)["status"][0] != 0
    with pytest.raises(okb.OkbError):
        ctx.optimize(3, 1, max_iterations=2)
    ctx.upload(3, w)                        # a full upload clears the error
    assert ctx.optimize(3, 1, max_iterations=2)[0]["termination"] != 6
