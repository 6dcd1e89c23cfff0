"""World-size-2 gloo test of the multi-GPU exchange logic (fugue_b200/dist.py) on CPU:
count all-gather -> all-to-all per column -> segment rearrangement, checked against the oracle's
partition of the concatenated table."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hash_partition as hp

NUM = 16
ROWS = [5000, 3777]


def _shard(rank: int):
    rng = np.random.default_rng(100 + rank)
    n = ROWS[rank]
    return [rng.integers(0, 500, n).astype("int64"), rng.standard_normal(n),
            (np.arange(n) + rank * 1_000_000).astype("int64")]


def _worker(rank: int, world: int, port: int, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fugue_b200.dist import ExchangePlan, exchange_column, gather_counts, owner_range, rearrange_cpu

        cols = _shard(rank)
        part, off = hp.partition_table(cols, [0], NUM)          # what K1-K3 do on each GPU
        counts = gather_counts(torch.from_numpy(np.diff(off)))
        assert counts.shape == (world, NUM)
        plan = ExchangePlan(counts, rank)
        assert sum(plan.send_rows) == ROWS[rank]
        recv = [exchange_column(torch.from_numpy(c), plan) for c in part]
        outs = rearrange_cpu(recv, plan)
        lo, hi = owner_range(NUM, world, rank)
        ret[rank] = ([o.numpy() for o in outs], plan.out_offsets.numpy(), lo, hi)
    finally:
        dist.destroy_process_group()


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_rank_exchange_matches_global_oracle():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    shards = [_shard(r) for r in range(world)]
    glob = [np.concatenate([s[c] for s in shards]) for c in range(3)]
    exp_cols, exp_off = hp.partition_table(glob, [0], NUM)
    covered = 0
    for r in range(world):
        outs, off, lo, hi = ret[r]
        assert off[0] == 0 and len(off) == hi - lo + 1
        for j, p in enumerate(range(lo, hi)):
            a, b = exp_off[p], exp_off[p + 1]
            assert off[j + 1] - off[j] == b - a
            for c in range(3):
                got = outs[c][off[j]:off[j + 1]]
                assert np.array_equal(got.view("u1"), exp_cols[c][a:b].view("u1")), (r, p, c)
            covered += b - a
    assert covered == sum(ROWS)


def test_owner_ranges_cover_everything():
    from fugue_b200.dist import owner_range

    for num in (2, 3, 16, 255, 256, 1000):
        for world in (1, 2, 3, 8):
            if num < world:
                continue
            r = [owner_range(num, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == num
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert all(b > a for a, b in r)
