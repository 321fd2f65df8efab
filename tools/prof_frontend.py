"""Tiny driver for ncu: frontend kernels at the cfg-3 shape and OKVIS' production parameters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from okvis_b200 import capi, images, synthetic
ctx = capi.Context(0, 1)
cam = synthetic.make_window(1, 0).cameras[0]
dense = images.textured_image(0x0B200 + 3000, n_shapes=2600)
left, right = images.stereo_pair()
for rep in range(2):
    kl, dl = ctx.detect_describe(dense, cam, np.eye(3), uniformity_radius=15.0, max_keypoints=1000, cam_slot=0)
    kr, dr = ctx.detect_describe(left, cam, np.eye(3), uniformity_radius=40.0, max_keypoints=400, cam_slot=1)
    rng = np.random.Generator(np.random.PCG64(5))
    A = rng.integers(0, 256, (1000, 48), dtype=np.uint8)
    B = A[rng.permutation(1000)].copy()
    B[:, :4] ^= rng.integers(0, 256, (1000, 4), dtype=np.uint8)
    ctx.hamming_match(A, B)
    ctx.hamming_candidates(dl, dl)
print(len(kl), len(kr))
