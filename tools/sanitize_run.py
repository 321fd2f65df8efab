#!/usr/bin/env python
"""Small end-to-end run of every product kernel for compute-sanitizer (memcheck / racecheck / synccheck):
cfg-1 and a reduced cfg-2 / cfg-5 optimize, a 2-rank landmark-sharded solve, hooks, frontend.  See tools/collect_profiles.sh."""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from okvis_b200 import capi, images, sharding, synthetic  # noqa: E402


def main():
    ctx = capi.Context(0, 4)
    ws = [synthetic.make_window(1, 0),
          synthetic.make_window(2, 0, cfg=dataclasses.replace(synthetic.CONFIGS[2], n_landmarks=300)),
          synthetic.make_window(5, 0, cfg=dataclasses.replace(synthetic.CONFIGS[5], n_landmarks=200)),
          synthetic.make_window(1, 2, cfg=dataclasses.replace(synthetic.CONFIGS[1], with_marg_prior=True))]
    ctx.upload_batch(0, ws, 2)
    s = ctx.optimize(0, 4, max_iterations=4)
    ctx.download_batch(0, 4)
    print("optimize:", [x["final_cost"] for x in s])
    cs = [capi.Context(0, 1) for _ in range(2)]
    capi.Context.shard_connect_local(cs, max_frames=5)
    for r, c in enumerate(cs):
        c.upload(0, sharding.shard_window(ws[0], r, 2)[0])
    for c in cs:
        c.optimize_async(0, 1, max_iterations=3)
    print("sharded:", [c.optimize_finish(0, 1)[0]["final_cost"] for c in cs])
    for c in cs:
        c.close()
    # resident window: incremental edits, device marginalisation, speed/bias removal (okb_graph.cuh, okb_marg.cuh)
    from okvis_b200 import abi
    w = ws[0]
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
    c2.marginalize(0, abi.make_marg_job([P, SB, P, SB], [0, 0, 1, 1], [0, 1, 2, -1], [0, 1, 0, 0], imu_terms=[0]))
    c2.remove_speed_bias(0, 0)
    s2 = c2.optimize(0, 1, max_iterations=3)
    m = c2.download_marg(0)
    print("marginalised:", m["n"], m["status"].tolist(), s2[0]["final_cost"])
    c2.close()
    left, right = images.stereo_pair()
    cam = ws[0].cameras[0]
    kd, dd = ctx.detect_describe(images.textured_image(n_shapes=2600), cam, np.eye(3), uniformity_radius=15.0, max_keypoints=1000, cam_slot=2)
    kl, dl = ctx.detect_describe(left, cam, np.eye(3), cam_slot=0)
    kr, dr = ctx.detect_describe(right, cam, np.eye(3), cam_slot=1)
    m = ctx.hamming_match(dl, dr)
    ctx.hamming_candidates(dl, dr)
    print("frontend:", len(kl), len(kr), int((m["pairs"]["index_a"] >= 0).sum()))
    ctx.close()


if __name__ == "__main__":
    main()
