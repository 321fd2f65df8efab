"""CPU: the scene and the acceptance criteria of the reference's own estimator test (okvis_ceres/test/TestEstimator.cpp:52-238,
case c = 0: fixed extrinsics) run through the oracle -- the restated Estimator::optimize, ImuError::propagation (state
prediction in addStates) and the marginalisation step driven with applyMarginalizationStrategy(2, 3)'s bookkeeping.

Scene (TestEstimator.cpp:60-205): 10 s of constant velocity (0, 1, 0) m/s, 100 Hz IMU with uniform noise, a stereo pair of the
test camera (PinholeCamera<EquidistantDistortion>::createTestObject(), 0.1 m baseline along y), a landmark grid on the
plane x = 3 (0.5 m pitch), K + 1 = 7 multi-frames 10/6 s apart (every third one a keyframe), keypoints = projection +
uniform(-1, 1) px with size 8, optimize(10) after every frame, then applyMarginalizationStrategy(2, 3) and a last
optimize(10).  Accepted (:226-237): |speed/bias error| < 0.04, rotation error < 1e-2, position error < 0.1 m at the
newest frame.  ImuParameters::sigma_bg / sigma_ba are left unset by the reference test; the shipped values are used."""
import dataclasses

import numpy as np

from okvis_b200 import abi, synthetic
from oracle import oracle_py as op

P, SB = abi.BLOCK_POSE, abi.BLOCK_SPEED_BIAS
DURATION, RATE, K = 10.0, 100.0, 6


def build_scene(seed=7):
    rng = np.random.Generator(np.random.PCG64(seed))
    imu = abi.make_imu_params(a_max=1000.0, g_max=1000.0, sigma_g_c=6.0e-4, sigma_a_c=2.0e-3, sigma_gw_c=3.0e-6, sigma_aw_c=2.0e-5,
                              tau=3600.0, g=9.81, rate=1000)
    dt = 1.0 / RATE
    n = int(DURATION * RATE) + 1
    smp = np.zeros(n, abi.imu_sample_dtype)
    t0 = 1_000_000_000
    smp["t_ns"] = t0 + np.round(np.arange(n) * dt * 1e9).astype(np.int64)
    smp["gyro"] = rng.uniform(-1, 1, (n, 3)) * imu.sigma_g_c * np.sqrt(dt)
    smp["acc"] = np.array([0, 0, imu.g]) + rng.uniform(-1, 1, (n, 3)) * imu.sigma_a_c * np.sqrt(dt)
    cam = abi.make_camera(abi.DIST_EQUIDISTANT, 752, 480, 350, 360, 378, 238, (-0.21, 0.14, 0.0006, 0.0003))
    cams = np.array([cam, cam], abi.camera_dtype)
    ext = np.array([[0, 0, 0, 0, 0, 0, 1.0], [0, 0.1, 0, 0, 0, 0, 1.0]])
    ys = np.arange(-10.0, DURATION * 0.1 + 10.0 + 1e-9, 0.5)
    zs = np.arange(-10.0, 10.0 + 1e-9, 0.5)
    lms = np.array([[3.0, y, z] for y in ys for z in zs])
    t_frames = t0 + np.round(np.arange(K + 1) * DURATION / K * 1e9).astype(np.int64)
    r_true = np.stack([np.array([0.0, k * DURATION / K, 0.0]) for k in range(K + 1)])
    obs = []
    for k in range(K + 1):
        for c in range(2):
            p_C = lms - r_true[k] - ext[c, :3]                      # T_WS and T_SC are pure translations
            px, ok = synthetic.project_points(cams[c], p_C)
            for j in np.nonzero(ok)[0]:
                obs.append((k, j, c, px[j] + rng.uniform(-1, 1, 2)))
    return dict(imu=imu, samples=smp, cams=cams, ext=ext, lms=lms, t_frames=t_frames, r_true=r_true, obs=obs)


def window_of(sc, poses, sbs, lm_est, frames, sb_frames, marg=None, drop_obs_of=()):
    """Window over `frames` (scene indices, ascending; speed/bias blocks exist for `sb_frames`), landmarks = those observed."""
    fidx = {f: i for i, f in enumerate(frames)}
    sidx = {f: i for i, f in enumerate(sb_frames)}
    ob = [o for o in sc["obs"] if o[0] in fidx and o[0] not in drop_obs_of]
    used = sorted({o[1] for o in ob})
    lidx = {j: i for i, j in enumerate(used)}
    obs = np.zeros(len(ob), abi.observation_dtype)
    for i, (k, j, c, z) in enumerate(ob):
        obs[i]["pose_idx"], obs[i]["lm_idx"], obs[i]["ext_idx"], obs[i]["cam_idx"] = fidx[k], lidx[j], c, c
        obs[i]["z"] = z
        obs[i]["sqrt_info"] = 1.0                              # keypoint size 8 -> 64 / size^2 = 1 (implementation/Estimator.hpp:62-65)
    ts = sc["samples"]["t_ns"]
    terms = []
    for a, b in zip(sb_frames[:-1], sb_frames[1:]):
        ta, tb = sc["t_frames"][a], sc["t_frames"][b]
        lo = max(int(np.searchsorted(ts, ta, side="right")) - 1, 0)
        hi = min(int(np.searchsorted(ts, tb, side="left")), len(ts) - 1)
        terms.append((fidx[a], sidx[a], fidx[b], sidx[b], ta, tb, lo, hi - lo + 1))
    terms = np.array(terms, abi.imu_term_dtype) if terms else np.zeros(0, abi.imu_term_dtype)
    hp = np.concatenate([lm_est[used], np.ones((len(used), 1))], 1)
    pp = np.zeros(1, abi.pose_prior_dtype)                     # first pose: Estimator.cpp:238-262 (uncertain yaw only through 1e4)
    pp["pose_idx"], pp["meas"] = 0, np.array([0, 0, 0, 0, 0, 0, 1.0])
    pp["sqrt_info"] = np.diag([1e4, 1e4, 1e4, 0, 0, 1e4]).reshape(-1)
    sp = np.zeros(0 if marg is not None else 1, abi.sb_prior_dtype)
    if marg is None:                                           # first speed / bias prior: Estimator.cpp:270-285
        sp["sb_idx"], sp["meas"] = 0, np.zeros(9)
        sp["sqrt_info"] = np.diag([1, 1, 1] + [1 / sc["imu"].sigma_bg] * 3 + [1 / sc["imu"].sigma_ba] * 3).reshape(-1)
    w = synthetic.Window(poses=np.ascontiguousarray(np.stack([poses[f] for f in frames])),
                         speed_bias=np.ascontiguousarray(np.stack([sbs[f] for f in sb_frames])), extrinsics=sc["ext"].copy(),
                         extrinsics_fixed=np.ones(2, np.uint8), landmarks=np.ascontiguousarray(hp), cameras=sc["cams"], obs=obs,
                         imu_terms=terms, imu_samples=sc["samples"], imu_params=sc["imu"], pose_priors=pp, sb_priors=sp,
                         relpose_terms=np.zeros(0, abi.relpose_dtype), marg=marg, truth=None, name="TestEstimator")
    return w, used


def test_reference_estimator_scene_meets_the_reference_acceptance_criteria():
    sc = build_scene()
    poses, sbs = {}, {}
    lm_est = sc["lms"].copy()                                  # estimator.addLandmark with the true points (:141-146)
    poses[0] = np.array([0, 0, 0, 0, 0, 0, 1.0])
    sbs[0] = np.zeros(9)                                       # addStates on an empty estimator: zero speed and biases
    costs = []
    for k in range(K + 1):
        if k > 0:                                              # addStates: ImuError::propagation from the previous estimate (Estimator.cpp:145-147)
            ts = sc["samples"]["t_ns"]
            ta, tb = sc["t_frames"][k - 1], sc["t_frames"][k]
            lo = max(int(np.searchsorted(ts, ta, side="right")) - 1, 0)
            hi = min(int(np.searchsorted(ts, tb, side="left")), len(ts) - 1)
            n, p, s, _, _ = op.imu_propagate(sc["imu"], sc["samples"][lo:hi + 1], ta, tb, poses[k - 1], sbs[k - 1], want_cov=False)
            assert n > 0
            poses[k], sbs[k] = p, s
        frames = list(range(k + 1))
        w, used = window_of(sc, poses, sbs, lm_est, frames, frames)
        pb = op.OracleProblem(w)
        so = pb.solve(10, 4)                                   # estimator.optimize(10, 4, false)
        st = pb.state(with_quality=False)
        for i, f in enumerate(frames):
            poses[f], sbs[f] = st["poses"][i], st["speed_bias"][i]
        lm_est[used] = st["landmarks"][:, :3] / st["landmarks"][:, 3:4]
        costs.append((so["initial_cost"], so["final_cost"]))
        assert so["final_cost"] <= so["initial_cost"]
    # every optimize after the first frames starts near the optimum: the scene is consistent
    assert costs[-1][1] < 1.5 * len([o for o in sc["obs"]])    # about one unit of (uniform +-1 px)^2 / 2 per residual pair at most

    # ---- applyMarginalizationStrategy(2, 3) (Estimator.cpp:434-773): the three newest frames (4, 5, 6) keep their speed/bias
    # blocks; of the older ones the keyframes 0 and 3 keep their poses (2 keyframes allowed), the non-keyframes 1 and 2 go
    # entirely -- their observations are dropped, not linearised (:600-640) -- and speed/bias 0..3 are marginalised with the
    # IMU terms 0-1 .. 3-4 and the first speed/bias prior
    frames = list(range(K + 1))
    w, used = window_of(sc, poses, sbs, lm_est, frames, frames)
    pb = op.OracleProblem(w)
    kinds = [P, SB, P, SB, P, SB, P, SB, P, SB]
    idx = [0, 0, 1, 1, 2, 2, 3, 3, 4, 4]
    flags = [0, 1, 1, 1, 1, 1, 0, 1, 0, 0]
    job = abi.make_marg_job(kinds, idx, [-1] * 10, flags, imu_terms=[0, 1, 2, 3], sb_priors=[0])
    o = pb.marginalize(job)
    assert o["n"] == 6 * 3 + 9
    assert [int(x) for x in o["block_kind"]] == [P, P, P, SB] and [int(x) for x in o["block_idx"]] == [0, 3, 4, 4]
    lam = np.linalg.eigvalsh(0.5 * (o["H"] + o["H"].T))
    assert lam.min() > -1e-6 * lam.max()                       # the Schur complement of a PSD system stays PSD
    keep_frames, keep_sb = [0, 3, 4, 5, 6], [4, 5, 6]
    marg = dict(block_kind=o["block_kind"], block_idx=np.array([0, 1, 2, 0], np.uint32), x0=o["x0"], J=np.ascontiguousarray(o["J"]),
                e0=o["e0"])
    w2, used2 = window_of(sc, poses, sbs, lm_est, keep_frames, keep_sb, marg=marg)
    pb2 = op.OracleProblem(w2)
    so = pb2.solve(10, 4)                                      # the last optimize (:215-217)
    st = pb2.state(with_quality=False)
    assert so["final_cost"] <= so["initial_cost"] * (1 + 1e-12)

    # ---- the reference's assertions (:226-237) on the newest frame
    T = st["poses"][-1]
    sb_err = np.linalg.norm(st["speed_bias"][-1] - np.array([0, 1, 0, 0, 0, 0, 0, 0, 0.0]))
    rot_err = 2 * np.linalg.norm(T[3:6])                       # truth is the identity rotation
    pos_err = np.linalg.norm(T[:3] - np.array([0, DURATION, 0.0]))
    assert sb_err < 0.04, sb_err
    assert rot_err < 1e-2, rot_err
    assert pos_err < 1e-1, pos_err
