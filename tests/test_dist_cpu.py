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


def _shard(rank: int, rows=None, key_range: int = 500):
    rng = np.random.default_rng(100 + rank)
    n = (rows or ROWS)[rank]
    return [rng.integers(0, key_range, n).astype("int64"), rng.standard_normal(n),
            (np.arange(n) + rank * 1_000_000).astype("int64")]


def _worker(rank: int, world: int, port: int, ret, num: int = NUM, rows=None, key_range: int = 500):
    global NUM, ROWS
    NUM, ROWS = num, list(rows or ROWS)
    _shard_ = lambda r: _shard(r, ROWS, key_range)  # noqa: E731
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fugue_b200.dist import ExchangePlan, exchange_column, gather_counts, owner_range, rearrange_cpu

        cols = _shard_(rank)
        part, off = hp.partition_table(cols, [0], NUM)          # what K1-K3 do on each GPU
        counts = gather_counts(torch.from_numpy(np.diff(off)))
        assert counts.shape == (world, NUM)
        plan = ExchangePlan(counts, rank)
        assert sum(plan.send_rows) == ROWS[rank]
        recv = [exchange_column(torch.from_numpy(c), plan) for c in part]
        outs = rearrange_cpu(recv, plan)
        lo, hi = owner_range(NUM, world, rank)
        # the one-run-per-source pull and the (source, partition) segment index describe the same layout
        from fugue_b200.dist import compact_plan

        src, dst, ln, off = compact_plan(plan.segment_offsets)
        assert torch.equal(src, plan.seg_src_off) and torch.equal(dst, plan.seg_dst_off)
        assert torch.equal(ln, plan.seg_len) and torch.equal(off, plan.out_offsets)
        a, b = int(plan.pull_start[rank]), int(plan.recv_base[rank])
        for c, r in zip(part, recv):
            assert np.array_equal(r[b:b + plan.recv_rows[rank]].numpy(), c[a:a + plan.recv_rows[rank]])
        assert int(plan.segment_offsets[-1, -1]) == plan.total_recv
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
@pytest.mark.parametrize("world,num,rows,key_range", [
    (2, 16, [5000, 3777], 500),
    (3, 7, [0, 2500, 1], 500),          # an empty shard, a one-row shard, partitions not divisible by the ranks
    (3, 3, [40, 40, 40], 1),            # one key: every row lands in one partition of one rank
    (2, 256, [300, 200], 100000),       # more partitions than distinct keys per rank: many empty segments
], ids=["2x16", "3x7-empty-shard", "3x3-one-key", "2x256-sparse"])
def test_exchange_matches_global_oracle(world, num, rows, key_range):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, num, rows, key_range), nprocs=world, join=True)
    NUM, ROWS = num, rows
    shards = [_shard(r, rows, key_range) for r in range(world)]
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


def _dict_worker(rank: int, world: int, port: int, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types

        import pyarrow as pa

        from fugue_b200.dist import DistributedB200Engine
        from fugue_b200.schema import Schema
        from fugue_b200.table import B200Table

        words = [["b", "a", "c"], ["c", "d", "a", "e"]][rank]
        codes = torch.tensor([[0, 1, 2, 1, 0], [3, 0, 1, 2, 2, 0]][rank], dtype=torch.int32)
        valid = torch.tensor([[1, 1, 1, 0, 1], [1, 1, 1, 1, 1, 1]][rank], dtype=torch.uint8)
        t = B200Table(Schema("s:str,v:long"), [codes, torch.arange(len(codes))], [valid, None],
                      {"s": pa.array(words, type=pa.string())})
        fake = types.SimpleNamespace(_world=world, _group=None)
        # second table (the other side of a join): its own dictionary, other order, one new word
        words2 = [["e", "b"], ["z", "a"]][rank]
        t2 = B200Table(Schema("s:str"), [torch.tensor([1, 0, 1], dtype=torch.int32)], None,
                       {"s": pa.array(words2, type=pa.string())})
        g, g2 = DistributedB200Engine._globalize_tables(fake, [t, t2])
        decoded = [None if m == 0 else g.dictionaries["s"][int(c)].as_py()
                   for c, m in zip(g.columns[0].tolist(), valid.tolist())]
        decoded2 = [g2.dictionaries["s"][int(c)].as_py() for c in g2.columns[0].tolist()]
        ret[rank] = (g.dictionaries["s"].to_pylist(), decoded, g2.dictionaries["s"].to_pylist(), decoded2)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_string_dictionaries_are_unified_across_ranks():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dict_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    # ONE union per column name over ranks AND tables (both sides of a join share the code space),
    # first-appearance order: rank 0 (table 1, table 2), then rank 1
    assert ret[0][0] == ret[1][0] == ret[0][2] == ret[1][2] == ["b", "a", "c", "e", "d", "z"]
    assert ret[0][1] == ["b", "a", "c", None, "b"]                      # codes re-mapped, NULL kept
    assert ret[1][1] == ["e", "c", "d", "a", "a", "c"]
    assert ret[0][3] == ["b", "e", "b"] and ret[1][3] == ["a", "z", "a"]


def test_partial_final_decomposition_of_aggregates():
    """Host logic shared by the multi-GPU group-by and COUNT(DISTINCT): SUM/COUNT/MIN/MAX/AVG as
    (partial, final, post) and the AVG finishing step, on CPU tensors."""
    import pyarrow as pa

    from fugue_b200.column import all_cols, col, functions as ff
    from fugue_b200.dataframe import B200DataFrame
    from fugue_b200.execution_engine import decompose_aggs, finish_avgs
    from fugue_b200.schema import Schema
    from fugue_b200.table import B200Table

    aggs = [ff.sum(col("v")).alias("s"), ff.count(all_cols()).alias("c"), ff.avg(col("v")).alias("m"),
            ff.max(col("v")).alias("hi")]
    partial, final, post = decompose_aggs(aggs)
    assert [str(a) for a in partial] == ["SUM(v) AS __p0", "COUNT(*) AS __p1", "SUM(v) AS __p2s", "COUNT(v) AS __p2c",
                                         "MAX(v) AS __p3"]
    assert [str(a) for a in final] == ["SUM(__p0) AS s", "SUM(__p1) AS c", "SUM(__p2s) AS __p2s",
                                       "SUM(__p2c) AS __p2c", "MAX(__p3) AS hi"]
    assert post == [("m", "__p2s", "__p2c")]
    t = B200Table(Schema("k:long,s:double,c:long,__p2s:double,__p2c:long,hi:double"),
                  [torch.tensor([1, 2]), torch.tensor([3.0, 4.0], dtype=torch.float64), torch.tensor([2, 0]),
                   torch.tensor([3.0, 0.0], dtype=torch.float64), torch.tensor([2, 0]),
                   torch.tensor([2.0, 0.0], dtype=torch.float64)])
    res = finish_avgs(B200DataFrame(t), post, ["k", "s", "c", "m", "hi"]).native
    assert res.schema == Schema("k:long,s:double,c:long,m:double,hi:double")
    i = res.schema.index_of_key("m")
    assert res.columns[i].tolist()[0] == 1.5 and res.valid[i].tolist() == [1, 0]   # AVG of no rows is NULL
    with pytest.raises(NotImplementedError):
        decompose_aggs([ff.first(col("v")).alias("f")])
