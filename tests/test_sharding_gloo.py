"""world_size-2 gloo test of the N>1 host path: window partition + measurement reduction + summary
gather (the only multi-rank logic of the hot path; windows need no data-path collective).  Each rank
solves ITS windows with the CPU oracle standing in for the device (this is a test of the plumbing)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from okvis_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_windows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from okvis_b200 import synthetic
        from oracle import oracle_py as op
        mine = sharding.shard_indices(n_windows, world, rank)
        summaries = []
        for w_idx in mine:
            p = op.OracleProblem(synthetic.make_window(1, w_idx))
            s = p.solve(3, 1)
            summaries.append({"window": w_idx, "final_cost": s["final_cost"], "iterations": s["iterations"]})
        t, c = sharding.reduce_measurement(dist, torch.device("cpu"), [10.0 + rank], [sum(x["iterations"] for x in summaries), len(mine)])
        merged = sharding.gather_summaries(dist, summaries)
        if rank == 0:
            q.put((t, c, merged))
    finally:
        dist.destroy_process_group()


def test_two_rank_window_sharding():
    world, n_windows = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_windows, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, c, merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == [11.0]                        # max over ranks
    assert c == [3.0 * n_windows, float(n_windows)]   # iterations and window counts summed over ranks
    assert [m["window"] for m in merged] == list(range(n_windows))
    # same answer as a single-rank run
    from okvis_b200 import synthetic
    from oracle import oracle_py as op
    for m in merged:
        s = op.OracleProblem(synthetic.make_window(1, m["window"])).solve(3, 1)
        assert abs(s["final_cost"] - m["final_cost"]) < 1e-12 * s["final_cost"]


def test_partition_is_exact_cover():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            all_idx = sorted(i for r in range(world) for i in sharding.shard_indices(n, world, r))
            assert all_idx == list(range(n))


def test_landmark_shards_partition_one_window():
    """shard_window: the shards' landmarks / observations are an exact cover of the window; dense blocks replicated."""
    from okvis_b200 import synthetic
    w = synthetic.make_window(1, 0)
    for world in (2, 3, 8):
        shards = [sharding.shard_window(w, r, world) for r in range(world)]
        assert sorted(np.concatenate([idx for _, idx in shards]).tolist()) == list(range(len(w.landmarks)))
        assert sum(len(s.obs) for s, _ in shards) == len(w.obs)
        for s, idx in shards:
            assert np.array_equal(s.landmarks, w.landmarks[idx]) and s.poses is w.poses and s.imu_terms is w.imu_terms
            assert s.obs["lm_idx"].max() < len(idx)
        back = np.concatenate([np.stack([idx[s.obs["lm_idx"]], s.obs["pose_idx"], s.obs["cam_idx"]], 1) for s, idx in shards])
        orig = np.stack([w.obs["lm_idx"], w.obs["pose_idx"], w.obs["cam_idx"]], 1)
        assert sorted(map(tuple, back.tolist())) == sorted(map(tuple, orig.tolist()))
