"""Development probe run on the GPU box: hook parity, solver parity vs the oracle, first timings.
Writes a text report to gpurun_out/probe.txt (tests/ are the real gate; this is for diagnosis)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from okvis_b200 import abi, capi, synthetic  # noqa: E402
from oracle import oracle_py as op  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, "probe.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    log.write(s + "\n")
    log.flush()


def main():
    ctx = capi.Context(0, 4)
    from test_oracle_functors import CAMS, make_test_cam, rand_pose, _imu_case
    rng = np.random.default_rng(0)
    # ---- hooks
    for name in CAMS:
        cam = make_test_cam(name)
        n = 256
        pose = np.stack([rand_pose(rng, 1.0, 0.5) for _ in range(n)])
        ext = np.stack([rand_pose(rng, 0.1, 0.2) for _ in range(n)])
        hp = np.zeros((n, 4))
        for i in range(n):
            p_C = np.array([rng.uniform(-1, 1), rng.uniform(-0.7, 0.7), rng.uniform(0.5, 8.0)])
            R_SC, R_WS = synthetic.R_from_quat(ext[i, 3:]), synthetic.R_from_quat(pose[i, 3:])
            p_W = R_WS @ (R_SC @ p_C + ext[i, :3]) + pose[i, :3]
            w = rng.uniform(0.2, 1.5)
            hp[i] = np.concatenate([p_W * w, [w]])
        z = rng.uniform([0, 0], [752, 480], (n, 2))
        sq = rng.uniform(0.5, 2.0, n)
        a = ctx.eval_reprojection(cam, pose, hp, ext, z, sq)
        b = op.eval_reprojection(cam, pose, hp, ext, z, sq)
        P("reproj", name, [float(np.abs(x - y).max() / max(1, np.abs(y).max())) for x, y in zip(a, b)])
    prm, s, t0, t1 = _imu_case(rng)
    pose0 = rand_pose(rng, 1.0, 0.5)
    sb0 = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3)])
    n0, pose1, sb1, P0, F0 = op.imu_propagate(prm, s, t0, t1, pose0, sb0)
    n1, p1, s1, P1, F1 = ctx.imu_propagate(prm, s, t0, t1, pose0, sb0)
    P("propagate", n0, n1, np.abs(p1 - pose1).max(), np.abs(s1 - sb1).max(), np.abs(P1 - P0).max() / np.abs(P0).max(),
      np.abs(F1 - F0).max())
    r0, J0, sq0, redo0 = op.eval_imu(prm, s, t0, t1, pose0, sb0, pose1, sb1)
    r1, J1, sq1, redo1 = ctx.eval_imu(prm, s, t0, t1, pose0, sb0, pose1, sb1)
    P("imu eval", redo0, redo1, np.abs(r1 - r0).max(), [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(J1, J0)],
      np.abs(sq1.T @ sq1 - sq0.T @ sq0).max() / np.abs(sq0.T @ sq0).max())

    # ---- solver parity
    for cfg_id in (1, 2):
        w = synthetic.make_window(cfg_id, 0)
        ctx.upload(0, w)
        t = time.time()
        sg = ctx.optimize(0, 1, max_iterations=10)[0]
        tg = time.time() - t
        got = ctx.download(0)
        ref = op.OracleProblem(w)
        t = time.time()
        so = ref.solve(10, 1)
        to = time.time() - t
        rst = ref.state()
        P("cfg", cfg_id, "obs", len(w.obs))
        P("  gpu   ", json.dumps(sg))
        P("  oracle", json.dumps({k: v for k, v in so.items() if k not in ("trace", "phase_times")}))
        P("  oracle trace cost", so["trace"][:, 0].tolist())
        P("  rel cost diff", abs(sg["final_cost"] - so["final_cost"]) / so["final_cost"])
        P("  max |dpose t|", np.abs(got["poses"][:, :3] - rst["poses"][:, :3]).max(), "|dq|",
          np.abs(got["poses"][:, 3:] - rst["poses"][:, 3:]).max(), "|dsb|",
          np.abs(got["speed_bias"] - rst["speed_bias"]).max(), "|dlm|",
          np.abs(got["landmarks"] - rst["landmarks"]).max(), "|dq|", np.abs(got["quality"] - rst["quality"]).max())
        P("  wall gpu %.4f s oracle %.4f s" % (tg, to))
        # repeat on resident data for timing
        for rep in range(3):
            ctx.reset(0, 1)
            t = time.time()
            sg = ctx.optimize(0, 1, max_iterations=10)[0]
            P("  resident optimize wall %.5f s, device solve_time %.6f s, it %d" % (time.time() - t, sg["solve_time_s"], sg["iterations"]))
    # batch of cfg-2 windows
    B = 64
    ctx2 = capi.Context(0, B)
    ws = [synthetic.make_window(2, i) for i in range(8)]
    for i in range(B):
        ctx2.upload(i, ws[i % 8])
    for rep in range(3):
        ctx2.reset(0, B)
        t = time.time()
        ss = ctx2.optimize(0, B, max_iterations=10)
        dt = time.time() - t
        its = sum(x["iterations"] for x in ss)
        P("batch %d: wall %.4f s, iterations %d, %.1f iter/s, solve_time %.5f" % (B, dt, its, its / dt, ss[0]["solve_time_s"]))
    P("launches", ctx.kernel_launches, ctx2.kernel_launches)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # noqa
        import traceback
        P("EXCEPTION", traceback.format_exc())
        raise
