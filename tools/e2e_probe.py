"""Where does the end-to-end (host buffers) time go?  Times upload / optimize / download of Be windows separately."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from okvis_b200 import capi, synthetic
Be = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ctx = capi.Context(0, 2 * Be)
ws = [synthetic.make_window(2, i) for i in range(8)]
for i in range(2 * Be):
    ctx.upload(i, ws[i % 8])
ctx.optimize(0, Be, max_iterations=10)
torch.cuda.synchronize()
for T in (1, 4, 8, 16):
    pool = ThreadPoolExecutor(T)
    t0 = time.perf_counter(); list(pool.map(lambda i: ctx.upload(i, ws[i % 8]), range(Be))); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ctx.optimize(0, Be, max_iterations=10); t3 = time.perf_counter()
    list(pool.map(lambda i: ctx.download(i), range(Be))); t4 = time.perf_counter()
    print("threads %2d: upload host %.2f ms (+drain %.2f), optimize %.2f ms, download %.2f ms" % (
        T, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
    pool.shutdown()
# python-only part of upload: desc()
t0 = time.perf_counter()
for i in range(Be): ws[i % 8].desc()
print("desc() x%d: %.2f ms" % (Be, (time.perf_counter() - t0) * 1e3))
