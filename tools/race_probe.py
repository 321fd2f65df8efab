import os, sys, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from okvis_b200 import abi, capi, synthetic
w = synthetic.make_window(1, 0)
K, L = len(w.poses), len(w.landmarks)
c2 = capi.Context(0, 1)
c2.reserve(0, K, L, len(w.obs) + 512, len(w.imu_samples) + 64, 80)
c2.upload(0, w)
c2.optimize(0, 1, max_iterations=3)
last, first = np.zeros(L, int), np.full(L, 99)
np.maximum.at(last, w.obs["lm_idx"], w.obs["pose_idx"])
np.minimum.at(first, w.obs["lm_idx"], w.obs["pose_idx"])
lms = np.nonzero((first == 0) & (last <= 2))[0].astype(np.uint32)
P, SB = abi.BLOCK_POSE, abi.BLOCK_SPEED_BIAS
c2.marginalize(0, abi.make_marg_job([P, SB, P, SB, P], [0, 0, 1, 1, 2], [-1] * 5, [1, 1, 0, 0, 0], imu_terms=[0], sb_priors=[0], landmarks=lms))
c2.remove_landmarks(0, lms)
c2.remove_frame(0, 0, 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
if mode == "full":
    c2.marginalize(0, abi.make_marg_job([P, SB, P, SB], [0, 0, 1, 1], [0, 1, 2, -1], [0, 1, 0, 0], imu_terms=[0]))
    c2.remove_speed_bias(0, 0)
print(c2.optimize(0, 1, max_iterations=3)[0])
m = c2.download_marg(0)
print(m["n"], m["block_kind"], m["block_idx"], m["status"])
