"""Host-side logic of the multi-GPU modes, on CPU with the gloo backend and world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_odometry_b200 import sharding


def test_shard_bounds_cover_in_order():
    for m in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(m, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_and_scale():
    assert sharding.shard_plan([10, 0, 5], 0) == (0, 15)
    assert sharding.shard_plan([10, 0, 5], 2) == (10, 15)
    r = np.abs(np.random.default_rng(0).normal(0, 0.3, 5000))
    s = sharding.scale_from_moments(len(r), float(r.sum()), float((r * r).sum()))
    assert abs(s - r.std() / 6.0) < 1e-12


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    res = np.abs(rng.normal(0, 0.2, 1001))            # the "global" accepted residuals, in query order
    lo, hi = sharding.shard_bounds(len(res), world, rank)
    mine = res[lo:hi]
    # collective 1: counts + raw moments
    stats = torch.tensor([len(mine), mine.sum(), (mine * mine).sum()], dtype=torch.float64)
    gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, stats)
    g = torch.stack(gathered).numpy()
    offset, total = sharding.shard_plan(g[:, 0], rank)
    scale = sharding.scale_from_moments(total, g[:, 1].sum(), g[:, 2].sum())
    # collective 2: the sample drawn by GLOBAL position: owner contributes, others add 0
    idx = np.random.default_rng(1).permutation(total)[:100]
    sample = torch.zeros(128, dtype=torch.float64)
    for j, ci in enumerate(idx):
        if offset <= ci < offset + len(mine):
            sample[j] = mine[ci - offset] / scale
    dist.all_reduce(sample)
    # collective 3: 28 partial sums
    acc = torch.tensor(np.r_[np.full(27, mine.sum()), len(mine)], dtype=torch.float64)
    dist.all_reduce(acc)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([0.01 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(out, offset=offset, total=total, scale=scale, sample=sample.numpy(), acc=acc.numpy(), t=t.numpy(), ref_sample=res[idx] / scale,
                 ref_sum=res.sum(), ref_scale=res.std() / 6.0)
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_exchange_world2(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    assert int(z["total"]) == 1001 and int(z["offset"]) == 0
    assert abs(float(z["scale"]) - float(z["ref_scale"])) < 1e-12
    assert np.array_equal(z["sample"][:100], z["ref_sample"])          # every sampled residual arrives exactly once
    assert np.all(z["sample"][100:] == 0)
    assert abs(z["acc"][0] - float(z["ref_sum"])) < 1e-9 and z["acc"][27] == 1001
    assert abs(float(z["t"][0]) - 0.02) < 1e-12
