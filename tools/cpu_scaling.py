"""How does the CPU oracle scale with host threads on this box?  (picks the reference arm's thread count)"""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from okvis_b200 import synthetic
from oracle import oracle_py as op
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
ws = [synthetic.make_window(2, i) for i in range(8)]
def one(i):
    p = op.OracleProblem(ws[i % 8]); s = p.solve(10, 1); p.close(); return s["iterations"]
one(0)
for T in (1, 8, 16, 32, 64, 128):
    n = max(T, 16)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(T) as ex:
        it = sum(ex.map(one, range(n)))
    dt = time.perf_counter() - t0
    print("threads %3d: %d windows in %.2f s -> %.0f iterations/s" % (T, n, dt, it / dt), flush=True)
