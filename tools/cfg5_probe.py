import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from okvis_b200 import capi, synthetic
w = synthetic.make_window(5, 0)
ctx = capi.Context(0, 1)
ctx.upload(0, w)
ctx.optimize(0, 1, max_iterations=10)
for rep in range(2):
    ctx.reset(0, 1)
    ctx.profile_enable(True)
    t = time.time(); s = ctx.optimize(0, 1, max_iterations=10); dt = time.time() - t
    pr = ctx.profile_read()
    print("cfg5 wall %.4f s device %.5f s | A %.3f ms/launch x%d, S %.3f ms/launch x%d, Q %.3f ms" % (dt, s[0]["solve_time_s"], pr["landmarks_ms"] / max(pr["landmarks_launches"], 1), pr["landmarks_launches"], pr["solve_ms"] / max(pr["solve_launches"], 1), pr["solve_launches"], pr["quality_ms"]))
    print("   phases(us, summed over rounds):", json.dumps({k: round(v, 1) for k, v in list(ctx.debug_phase_us(0).items())[:7]}))
