"""Resident windows: incremental graph updates through the C-ABI (okb_window_add_frame / remove_frame /
add_observations / remove_observations / set_landmarks / remove_landmarks / set_states / set_priors) must give
BIT-IDENTICAL results to a full okb_window_upload of the same graph -- the device-side command interpreter and
compile (okb_graph.cuh) replace the reference's addStates / addObservation / removeObservation bookkeeping
(okvis_ceres/src/Estimator.cpp:110-413)."""
import dataclasses

import numpy as np
import pytest

from okvis_b200 import abi, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 4)
    yield c
    c.close()


def sub_window(w, frames, keep_obs=None):
    """The window restricted to `frames` (ascending positions), re-indexed; landmarks keep their slots."""
    frames = list(frames)
    pos = {f: i for i, f in enumerate(frames)}
    obs = w.obs if keep_obs is None else w.obs[keep_obs]
    obs = obs[np.isin(obs["pose_idx"], frames)].copy()
    obs["pose_idx"] = np.array([pos[f] for f in obs["pose_idx"]], np.uint32)
    terms = []
    for t in w.imu_terms:
        if int(t["pose0"]) in pos and int(t["pose1"]) in pos:
            t = t.copy()
            t["pose0"], t["sb0"], t["pose1"], t["sb1"] = pos[int(t["pose0"])], pos[int(t["sb0"])], pos[int(t["pose1"])], pos[int(t["sb1"])]
            terms.append(t)
    terms = np.array(terms, abi.imu_term_dtype) if terms else np.zeros(0, abi.imu_term_dtype)
    pp = w.pose_priors[np.isin(w.pose_priors["pose_idx"], frames)].copy()
    pp["pose_idx"] = np.array([pos[int(f)] for f in pp["pose_idx"]], np.uint32)
    sp = w.sb_priors[np.isin(w.sb_priors["sb_idx"], frames)].copy()
    sp["sb_idx"] = np.array([pos[int(f)] for f in sp["sb_idx"]], np.uint32)
    return dataclasses.replace(w, poses=np.ascontiguousarray(w.poses[frames]), speed_bias=np.ascontiguousarray(w.speed_bias[frames]),
                               obs=np.ascontiguousarray(obs), imu_terms=terms, pose_priors=pp, sb_priors=sp)


def solve(ctx, win, n_iter, dims):
    s = ctx.optimize(win, 1, max_iterations=n_iter)[0]
    return s, ctx.download(win, dims=dims)


def assert_same(a, b):
    (sa, da), (sb, db) = a, b
    assert sa["final_cost"] == sb["final_cost"] and sa["initial_cost"] == sb["initial_cost"]
    assert sa["iterations"] == sb["iterations"] and sa["termination"] == sb["termination"]
    for k in ("poses", "speed_bias", "landmarks", "quality"):
        assert np.array_equal(da[k], db[k]), k


def term_samples(w, t):
    lo, n = int(t["sample_offset"]), int(t["sample_count"])
    tt = t.copy()
    tt["sample_offset"] = 0
    return tt, w.imu_samples[lo:lo + n]


def test_frames_added_one_by_one_equal_one_full_upload(ctx):
    w = synthetic.make_window(1, 0)
    K, L = len(w.poses), len(w.landmarks)
    dims = (K, K, L)
    ctx.upload(0, w)
    ref = solve(ctx, 0, 8, dims)
    # start from the first two frames; landmarks first seen later arrive with their frame (slots beyond the first two
    # frames' landmarks are created by okb_window_set_landmarks)
    first_seen = np.full(L, K)
    np.minimum.at(first_seen, w.obs["lm_idx"], w.obs["pose_idx"])
    w2 = sub_window(w, [0, 1])
    lm0 = w.landmarks.copy()
    lm0[first_seen >= 2] = 0.0
    w2 = dataclasses.replace(w2, landmarks=lm0)
    ctx.reserve(1, K, L, len(w.obs), len(w.imu_samples) + 64 * K)
    ctx.upload(1, w2)
    for k in range(2, K):
        tt, smp = term_samples(w, w.imu_terms[k - 1])
        ctx.add_frame(1, w.poses[k], w.speed_bias[k], tt, smp)
        new = np.nonzero(first_seen == k)[0]
        if len(new):
            ctx.set_landmarks(1, new, w.landmarks[new])
        ctx.add_observations(1, w.obs[w.obs["pose_idx"] == k])
        if k == 3:
            ctx.commit(1, 1)            # a commit in the middle must not change anything
    got = solve(ctx, 1, 8, dims)
    assert_same(ref, got)
    assert ctx.h2d_bytes(1) < 0.5 * ctx.h2d_bytes(0)      # the last commit carried frames, not the window


def test_remove_and_re_add_the_newest_frame(ctx):
    """The streaming pattern of bench.py's e2e leg: reset, drop the newest frame, add it again with its observations."""
    w = synthetic.make_window(2, 0, cfg=dataclasses.replace(synthetic.CONFIGS[2], n_landmarks=400))
    K, L = len(w.poses), len(w.landmarks)
    dims = (K, K, L)
    ctx.reserve(0, K, L, len(w.obs) + 4096, len(w.imu_samples) + 256)
    ctx.upload(0, w)
    ref = solve(ctx, 0, 6, dims)
    full_bytes = ctx.h2d_bytes(0)
    for _ in range(3):
        ctx.reset(0, 1)
        ctx.remove_frame(0, K - 1, K - 1)
        tt, smp = term_samples(w, w.imu_terms[K - 2])
        ctx.add_frame(0, w.poses[K - 1], w.speed_bias[K - 1], tt, smp)
        ctx.add_observations(0, w.obs[w.obs["pose_idx"] == K - 1])
        got = solve(ctx, 0, 6, dims)
        assert_same(ref, got)
        assert ctx.h2d_bytes(0) * 5 < full_bytes


def test_remove_oldest_frame_equals_upload_of_the_remaining_window(ctx):
    w = synthetic.make_window(1, 1)
    K, L = len(w.poses), len(w.landmarks)
    rest = sub_window(w, range(1, K))
    pp = np.zeros(1, abi.pose_prior_dtype)
    pp["pose_idx"], pp["meas"], pp["sqrt_info"] = 0, rest.poses[0], np.diag([1e4] * 3 + [0, 0, 1e4]).reshape(-1)
    sp = w.sb_priors.copy()
    sp["sb_idx"], sp["meas"] = 0, rest.speed_bias[0]
    rest = dataclasses.replace(rest, pose_priors=pp, sb_priors=sp)
    ctx.upload(0, rest)
    ref = solve(ctx, 0, 7, (K - 1, K - 1, L))
    ctx.upload(1, w)
    ctx.optimize(1, 1, max_iterations=2)         # a solve in between: estimates change ...
    ctx.reset(1, 1)                               # ... and are restored
    ctx.remove_frame(1, 0, 0)
    ctx.set_priors(1, pp, sp)
    got = solve(ctx, 1, 7, (K - 1, K - 1, L))
    assert_same(ref, got)


def test_remove_middle_frame_observations_and_landmarks(ctx):
    w = synthetic.make_window(1, 2)
    K, L = len(w.poses), len(w.landmarks)
    rng = np.random.Generator(np.random.PCG64(3))
    drop_lm = np.sort(rng.choice(L, 25, replace=False))
    cand = np.nonzero(~np.isin(w.obs["lm_idx"], drop_lm) & (w.obs["pose_idx"] != 2))[0]
    drop_obs = np.sort(rng.choice(cand, 300, replace=False))
    keep = np.ones(len(w.obs), bool)
    keep[drop_obs] = False
    keep &= ~np.isin(w.obs["lm_idx"], drop_lm)
    frames = [0, 1, 3, 4]
    rest = sub_window(w, frames, keep_obs=keep)
    lm = w.landmarks.copy()
    lm[drop_lm] = 0.0
    rest = dataclasses.replace(rest, landmarks=lm)
    ctx.upload(0, rest)
    ref = solve(ctx, 0, 6, (K - 1, K - 1, L))
    ctx.upload(1, w)
    keys = np.stack([w.obs["pose_idx"][drop_obs], w.obs["lm_idx"][drop_obs], w.obs["cam_idx"][drop_obs]], 1)
    ctx.remove_observations(1, keys)
    ctx.remove_landmarks(1, drop_lm)
    ctx.remove_frame(1, 2, 2)
    got = solve(ctx, 1, 6, (K - 1, K - 1, L))
    # the IMU terms around the removed frame are gone in both variants; everything else must agree bit by bit
    assert_same(ref, got)


def test_set_states_equals_upload_of_changed_states(ctx):
    w = synthetic.make_window(1, 3)
    K, L = len(w.poses), len(w.landmarks)
    poses, sb = w.poses.copy(), w.speed_bias.copy()
    poses[2, :3] += 0.01
    sb[3, :3] -= 0.02
    ctx.upload(0, dataclasses.replace(w, poses=poses, speed_bias=sb))
    ref = solve(ctx, 0, 5, (K, K, L))
    ctx.upload(1, w)
    ctx.set_states(1, [2], poses[2:3], [3], sb[3:4])
    got = solve(ctx, 1, 5, (K, K, L))
    assert_same(ref, got)


def test_marginalisation_prior_set_incrementally(ctx, oracle):
    cfg = dataclasses.replace(synthetic.CONFIGS[1], with_marg_prior=True)
    w = synthetic.make_window(1, 2, cfg=cfg)
    K, L = len(w.poses), len(w.landmarks)
    ctx.upload(0, w)
    ref = solve(ctx, 0, 6, (K, K, L))
    ctx.reserve(1, K, L, len(w.obs), len(w.imu_samples), w.marg["J"].shape[0])
    ctx.upload(1, dataclasses.replace(w, marg=None))
    ctx.set_priors(1, w.pose_priors, w.sb_priors, w.marg)
    got = solve(ctx, 1, 6, (K, K, L))
    assert_same(ref, got)


def test_device_side_validation_reports_at_optimize(ctx, okb):
    w = synthetic.make_window(1, 0)
    ctx.reserve(2, len(w.poses), len(w.landmarks), len(w.obs) + 16, len(w.imu_samples))
    ctx.upload(2, w)
    ctx.add_observations(2, w.obs[:1])                    # the same (frame, camera, landmark) twice
    with pytest.raises(okb.OkbError) as e:
        ctx.optimize(2, 1, max_iterations=2)
    assert e.value.code == abi.OKB_ERR_UNSUPPORTED
    ctx.upload(2, w)                                       # a full upload clears the error
    assert ctx.optimize(2, 1, max_iterations=2)[0]["termination"] != 6
    with pytest.raises(okb.OkbError) as e:                 # capacity is checked on the host, synchronously
        ctx.add_observations(2, np.concatenate([w.obs[:10]] * 3))
    assert e.value.code == abi.OKB_ERR_CAPACITY
    bad = w.obs[:1].copy()
    bad["pose_idx"] = 77
    with pytest.raises(okb.OkbError) as e:
        ctx.add_observations(2, bad)
    assert e.value.code == abi.OKB_ERR_INVALID_ARG
