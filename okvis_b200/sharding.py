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


def shard_window(window, rank, world):
    """Landmark-block shard of ONE window (BASELINE.json configs[4]): rank r keeps the
    landmarks with lm_idx % world == r and their observations; poses, speed/bias, extrinsics, IMU terms and priors
    are replicated.  Returns (shard window, global landmark indices of the shard's landmarks)."""
    import dataclasses
    L = len(window.landmarks)
    mine = np.arange(rank, L, world)
    local = np.full(L, -1, np.int64)
    local[mine] = np.arange(len(mine))
    keep = local[window.obs["lm_idx"]] >= 0
    obs = window.obs[keep].copy()
    obs["lm_idx"] = local[obs["lm_idx"]].astype(np.uint32)
    w = dataclasses.replace(window, landmarks=np.ascontiguousarray(window.landmarks[mine]), obs=np.ascontiguousarray(obs),
                            name="%s/shard%d of %d" % (window.name, rank, world))
    return w, mine


def merge_shard_results(n_landmarks, parts):
    """parts: list over ranks of (global landmark indices, download dict).  Dense blocks are identical on all ranks."""
    out = dict(poses=parts[0][1]["poses"], speed_bias=parts[0][1]["speed_bias"],
               landmarks=np.zeros((n_landmarks, 4)), quality=np.zeros(n_landmarks))
    for idx, d in parts:
        out["landmarks"][idx] = d["landmarks"]
        if d.get("quality") is not None:
            out["quality"][idx] = d["quality"]
    return out


def connect_shards(ctx, dist, device, rank, world, max_frames):
    """Multi-process mailbox set-up of a landmark-sharded context: export this rank's cudaIpc handle, all-gather the
    handles with torch.distributed (64 bytes per rank; plumbing only), map the peers."""
    import torch
    handle = torch.from_numpy(ctx.shard_export(rank, world, max_frames)).to(device)
    gathered = [torch.zeros_like(handle) for _ in range(world)]
    dist.all_gather(gathered, handle)
    ctx.shard_connect(np.stack([g.cpu().numpy() for g in gathered]))
