"""Tiny driver for ncu: B resident cfg-2 windows, one optimize() of a few iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from okvis_b200 import capi, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0, B)
ws = [synthetic.make_window(2, i) for i in range(min(B, 4))]
for i in range(B):
    ctx.upload(i, ws[i % len(ws)])
ctx.optimize(0, B, max_iterations=iters)
ctx.reset(0, B)
s = ctx.optimize(0, B, max_iterations=iters)
print(B, iters, s[0])
