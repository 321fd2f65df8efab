#!/usr/bin/env python
"""One rank of a landmark-sharded window solve (launched by torchrun; used by tests/test_gpu_shard.py and by hand):
exports its mailbox handle, all-gathers the handles with torch.distributed, connects, uploads its shard, optimizes,
and rank 0 writes every rank's estimates to --out."""
import argparse
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--landmarks", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=6)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from okvis_b200 import capi, sharding, synthetic
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = synthetic.CONFIGS[a.config]
    if a.landmarks:
        cfg = dataclasses.replace(cfg, n_landmarks=a.landmarks)
    w = synthetic.make_window(a.config, 0, cfg=cfg)
    ctx = capi.Context(local_rank, 1)
    sharding.connect_shards(ctx, dist, torch.device('cuda', local_rank), rank, world, len(w.poses))
    sw, idx = sharding.shard_window(w, rank, world)
    ctx.upload(0, sw)
    dist.barrier()
    s = ctx.optimize(0, 1, max_iterations=a.iterations)[0]
    d = ctx.download(0)
    res = dict(summary=s, idx=idx, stats=ctx.shard_stats(0), **d)
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0 and a.out:
        out = {"summaries": np.array([r["summary"] for r in allres], dtype=object)}
        for r, x in enumerate(allres):
            out["idx_%d" % r], out["poses_%d" % r], out["sb_%d" % r] = x["idx"], x["poses"], x["speed_bias"]
            out["lm_%d" % r], out["q_%d" % r] = x["landmarks"], x["quality"]
        np.savez(a.out, **out)
        print("shard stats:", [r["stats"] for r in allres], "summary:", allres[0]["summary"])
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
