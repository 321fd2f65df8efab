"""Multi-GPU plumbing for the hot path.  Keyframe windows are independent (no shared state between
okvis::Estimator instances), so the path shards by window with NO data-path collective:
window w runs on rank w mod world.  torch.distributed is used only to (a) agree on the partition and
(b) reduce the measurement (max time over ranks, sums of counters) -- works with nccl and gloo alike.
"""
import numpy as np


def shard_indices(n_items, world, rank):
    """Round-robin ownership: item i belongs to rank i % world (window w -> GPU w mod n)."""
    return list(range(rank, n_items, world))


def reduce_measurement(dist, device, elapsed_ms, counters):
    """Returns (max elapsed over ranks, counters summed over ranks).  `dist` may be None (single rank)."""
    import torch
    t = torch.tensor(list(elapsed_ms), dtype=torch.float64, device=device)
    c = torch.tensor([float(x) for x in counters], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in c]


def gather_summaries(dist, summaries):
    """All ranks' per-window summaries, ordered by global window index (rank-major round robin undone)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(summaries)
    world = dist.get_world_size()
    out = [None] * world
    dist.all_gather_object(out, list(summaries))
    n = sum(len(x) for x in out)
    merged = [None] * n
    for r, part in enumerate(out):
        for j, s in enumerate(part):
            merged[r + j * world] = s
    return merged
