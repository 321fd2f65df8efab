"""Diagnosis: device marginalisation vs oracle, one residual family at a time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from okvis_b200 import abi, capi, synthetic
from oracle import oracle_py as op
import test_gpu_marginalization as T
P, SB = abi.BLOCK_POSE, abi.BLOCK_SPEED_BIAS
ctx = capi.Context(0, 2)
w = synthetic.make_window(1, 0)
K, L = len(w.poses), len(w.landmarks)
job_full, lms, _ = T.first_job(op, w)
cases = {
    "sbprior": dict(kinds=[P, SB, P, SB, P], idx=[0, 0, 1, 1, 2], marg=[1, 1, 0, 0, 0], imu=[], sbp=[0], lms=[]),
    "imu": dict(kinds=[P, SB, P, SB, P], idx=[0, 0, 1, 1, 2], marg=[1, 1, 0, 0, 0], imu=[0], sbp=[], lms=[]),
    "imu_nomarg": dict(kinds=[P, SB, P, SB], idx=[0, 0, 1, 1], marg=[0, 0, 0, 0], imu=[0], sbp=[], lms=[]),
    "lms_nomarg": dict(kinds=[P, P, P], idx=[0, 1, 2], marg=[0, 0, 0], imu=[], sbp=[], lms=lms),
    "lms": dict(kinds=[P, SB, P, SB, P], idx=[0, 0, 1, 1, 2], marg=[1, 0, 0, 0, 0], imu=[], sbp=[], lms=lms),
    "all": dict(kinds=[P, SB, P, SB, P], idx=[0, 0, 1, 1, 2], marg=[1, 1, 0, 0, 0], imu=[0], sbp=[0], lms=lms),
}
for name, cs in cases.items():
    ctx.reserve(0, K, L, len(w.obs), len(w.imu_samples), 80)
    ctx.upload(0, w)
    ctx.optimize(0, 1, max_iterations=6)
    est = ctx.download(0)
    job = abi.make_marg_job(cs["kinds"], cs["idx"], [-1] * len(cs["kinds"]), cs["marg"], imu_terms=cs["imu"], sb_priors=cs["sbp"], landmarks=cs["lms"])
    ref = T.oracle_like_device(op, ctx, 0, w, est)
    ctx.marginalize(0, job)
    g = ctx.download_marg(0)
    o = ref.marginalize(job)
    dH = np.abs(g["H"] - o["H"])
    i, j = np.unravel_index(dH.argmax(), dH.shape)
    print("%-12s n %d/%d status %s rank %d  max|dH| %.3e (|H| %.3e) at (%d,%d)  max|db| %.3e (|b| %.3e)" % (
        name, g["n"], o["n"], g["status"].tolist(), o["rank"], dH.max(), np.abs(o["H"]).max(), i, j, np.abs(g["b0"] - o["b0"]).max(), np.abs(o["b0"]).max()))
    rel = dH / (np.abs(o["H"]) + 1e-300)
    print("   rows with rel err > 1e-8:", sorted(set(np.nonzero((dH > 1e-9 * np.abs(o["H"]).max()))[0].tolist())))

print("---- single landmarks")
ctx.reserve(0, K, L, len(w.obs), len(w.imu_samples), 80)
ctx.upload(0, w)
ctx.optimize(0, 1, max_iterations=6)
est = ctx.download(0)
we = T.at_estimates(w, est)
rows = []
for l in lms:
    job = abi.make_marg_job([P, P, P], [0, 1, 2], [-1] * 3, [0, 0, 0], landmarks=[l])
    ref = op.OracleProblem(we)
    ctx.upload(0, we)
    ctx.marginalize(0, job)
    g = ctx.download_marg(0)
    o = ref.marginalize(job)
    obs = we.obs[we.obs["lm_idx"] == l]
    rs = []
    for ob in obs:
        r, J0, J1, _ = op.eval_reprojection(we.cameras[ob["cam_idx"]], we.poses[ob["pose_idx"]][None], we.landmarks[l][None], we.extrinsics[ob["ext_idx"]][None], ob["z"][None], np.array([ob["sqrt_info"]]))
        rs.append(float(np.linalg.norm(r)))
    rows.append((np.abs(g["H"] - o["H"]).max() / max(np.abs(o["H"]).max(), 1e-300), int(l), len(obs), rs, np.abs(o["H"]).max()))
rows.sort(reverse=True)
for r in rows[:8]:
    print(r)
print("median rel err", np.median([r[0] for r in rows]))
