import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from okvis_b200 import capi, synthetic
from oracle import oracle_py as op
import test_gpu_shard as T
w = synthetic.make_window(1, 0)
s, parts, stats = T.solve_sharded_local(capi, w, 2, 60)
ref = op.OracleProblem(w); so = ref.solve(60, 2)
print("graph" if not os.environ.get("OKB_NO_GRAPH") else "nograph", [(x["iterations"], x["num_successful_steps"], x["termination"], x["final_cost"]) for x in s], (so["iterations"], so["num_successful_steps"], so["termination"], so["final_cost"]))
c = capi.Context(0, 1); c.upload(0, w); print("single", c.optimize(0, 1, max_iterations=60)[0])
