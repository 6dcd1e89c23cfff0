"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_check.py
Every rank hash-repartitions its row shard through DistributedB200Engine (NCCL all-to-all);
rank 0 checks the union against the oracle's partition of the concatenated table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fugue_b200.dataframe import B200DataFrame  # noqa: E402
from fugue_b200.dist import DistributedB200Engine, owner_range  # noqa: E402
from fugue_b200.partition import PartitionSpec  # noqa: E402
from fugue_b200.schema import Schema  # noqa: E402
from fugue_b200.table import B200Table  # noqa: E402
from oracle import hash_partition as hp  # noqa: E402

NUM = 256
_FAILS = []  # rank-0 comparison failures; collected, not raised, so that no rank skips a collective


class _guard:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        return self

    def __exit__(self, tp, val, tb):
        if tp is not None:
            import traceback

            _FAILS.append(f"{self.name}: " + "".join(traceback.format_exception_only(tp, val)).strip()[:400] +
                          " @ " + traceback.format_tb(tb)[-1].strip().split("\n")[0][-120:])
        return True


def shard(rank: int, n: int):
    rng = np.random.default_rng(7 + rank)
    return [rng.integers(0, 1 << 16, n).astype("int64"), rng.standard_normal(n),
            (np.arange(n) + rank * 10_000_000).astype("int64"), rng.integers(0, 255, n).astype("uint8")]


def check_relational(eng, rank, world, dev):
    """Distributed GROUP BY and JOIN vs the pandas oracle on the concatenated tables."""
    import pandas as pd

    from fugue_b200.column import all_cols, col, functions as ff
    from oracle import native_engine as ora

    def frames(r):
        rng = np.random.default_rng(1000 + r)
        n = 40_000 + 1000 * r
        fact = pd.DataFrame({"key": rng.integers(0, 5000, n), "v0": rng.standard_normal(n)})
        dim = pd.DataFrame({"key": rng.integers(0, 8000, 3000), "rv": rng.standard_normal(3000)})
        return fact, dim

    fact, dim = frames(rank)
    agg = eng.aggregate(eng.to_df(fact), PartitionSpec(by=["key"]),
                        [ff.sum(col("v0")).alias("s"), ff.count(all_cols()).alias("c"),
                         ff.avg(col("v0")).alias("a"), ff.max(col("v0")).alias("m")])
    jn = eng.join(eng.to_df(fact), eng.to_df(dim), "inner", ["key"])
    gathered = [None] * world
    dist.all_gather_object(gathered, (agg.as_pandas(), jn.as_pandas()))
    if rank == 0:
        with _guard("block 1"):
            facts, dims = zip(*[frames(r) for r in range(world)])
            F, D = pd.concat(facts, ignore_index=True), pd.concat(dims, ignore_index=True)
            got = pd.concat([g[0] for g in gathered], ignore_index=True).sort_values("key").reset_index(drop=True)
            exp = ora.aggregate(F, ["key"], {"s": ("v0", "sum"), "c": ("*", "count"), "a": ("v0", "avg"),
                                             "m": ("v0", "max")}).sort_values("key").reset_index(drop=True)
            assert len(got) == len(exp) and got["key"].is_unique
            assert np.array_equal(got["key"], exp["key"]) and np.array_equal(got["c"], exp["c"])
            assert np.array_equal(got["m"], exp["m"])
            assert np.max(np.abs(got["s"] - exp["s"]) / np.maximum(np.abs(exp["s"]), 1e-300)) <= 1e-9
            assert np.allclose(got["a"], exp["a"], rtol=1e-9)
            gj = pd.concat([g[1] for g in gathered], ignore_index=True)
            ej = ora.join(F, D, "inner")
            cols = list(ej.columns)
            a = gj.sort_values(cols).reset_index(drop=True)
            b = ej.sort_values(cols).reset_index(drop=True)
            pd.testing.assert_frame_equal(a, b, check_exact=True, check_dtype=False)
            print(f"dist relational ok: {len(got)} groups, {len(gj)} joined rows")
    # aggregating select with expressions, WHERE and HAVING: row-wise parts local, partials shuffled
    from fugue_b200.column import SelectColumns
    from oracle import expressions as OX

    sel = SelectColumns(col("key"), (ff.sum(col("v0") * 2) / ff.count(all_cols())).alias("m2"),
                        ff.max(col("v0") + col("key")).alias("mx"))
    where, having = col("v0") > -1.0, ff.count(all_cols()) > 3
    part = eng.select(eng.to_df(fact), sel, where=where, having=having)
    gathered = [None] * world
    dist.all_gather_object(gathered, part.as_pandas())
    if rank == 0:
        with _guard("block 2"):
            got = pd.concat(gathered, ignore_index=True).sort_values("key").reset_index(drop=True)
            exp = OX.select(F, sel, where=where, having=having).sort_values("key").reset_index(drop=True)
            assert len(got) == len(exp) and got["key"].is_unique
            assert np.array_equal(got["key"].to_numpy(), exp["key"].to_numpy(dtype="int64"))
            assert np.allclose(got["m2"].to_numpy(), exp["m2"].to_numpy(dtype="float64"), rtol=1e-9)
            assert np.allclose(got["mx"].to_numpy(), exp["mx"].to_numpy(dtype="float64"), rtol=1e-12)
            print(f"dist select ok: {len(got)} groups")
    # string key column: per-rank dictionaries are unified before the shuffle, so equal strings of
    # different ranks land in the same group
    words = [f"w{i:03d}" for i in range(50)]
    rng = np.random.default_rng(77 + rank)
    sdf = pd.DataFrame({"k": rng.choice(words[rank * 5:rank * 5 + 30], 20_000), "v": rng.integers(0, 100, 20_000)})
    sagg = eng.aggregate(eng.to_df(sdf), PartitionSpec(by=["k"]), [ff.sum(col("v")).alias("s"),
                                                                   ff.count(all_cols()).alias("c")])
    gathered = [None] * world
    dist.all_gather_object(gathered, (sagg.as_pandas(), sdf))
    if rank == 0:
        with _guard("block 3"):
            got = pd.concat([g[0] for g in gathered], ignore_index=True).sort_values("k").reset_index(drop=True)
            allrows = pd.concat([g[1] for g in gathered], ignore_index=True)
            exp = allrows.groupby("k").agg(s=("v", "sum"), c=("v", "size")).reset_index().sort_values("k").reset_index(drop=True)
            assert got["k"].is_unique and list(got["k"]) == list(exp["k"])
            assert np.array_equal(got["s"].to_numpy(), exp["s"].to_numpy()) and np.array_equal(got["c"].to_numpy(), exp["c"].to_numpy())
            print(f"dist string keys ok: {len(got)} groups")


def check_string_join(eng, rank, world):
    """String join keys: both sides are re-coded against ONE union dictionary (all ranks, both
    tables) before their codes are hashed, so equal strings meet on one rank."""
    import pandas as pd

    from oracle import native_engine as ora

    def frames(r):
        rng = np.random.default_rng(500 + r)
        words = [f"k{i:04d}" for i in range(400)]
        # different ranks / sides see different subsets in different orders -> different local codes
        left = pd.DataFrame({"k": rng.choice(words[r * 20:r * 20 + 300], 5000), "lv": rng.integers(0, 1000, 5000)})
        right = pd.DataFrame({"k": rng.permutation(words[100:])[:250], "rv": rng.integers(0, 1000, 250)})
        return left, right

    left, right = frames(rank)
    out = {}
    for how in ("inner", "left_outer", "anti"):
        out[how] = eng.join(eng.to_df(left), eng.to_df(right), how, ["k"]).as_pandas()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        with _guard("block 4"):
            ls, rs = zip(*[frames(r) for r in range(world)])
            L, R = pd.concat(ls, ignore_index=True), pd.concat(rs, ignore_index=True)
            for how in out:
                got = pd.concat([g[how] for g in gathered], ignore_index=True)
                exp = ora.join(L, R, how)
                cols = list(exp.columns)
                a = got[cols].sort_values(cols).reset_index(drop=True)
                b = exp.sort_values(cols).reset_index(drop=True)
                pd.testing.assert_frame_equal(a, b, check_exact=True, check_dtype=False)
            print(f"dist string-key joins ok: {len(got)} rows in the last")


def check_repartition(eng, rank, world, dev):
    n = 300_000 + 12345 * rank
    cols = shard(rank, n)
    t = B200Table(Schema("key:long,v:double,rid:long,b:ubyte"), [torch.from_numpy(c).to(dev) for c in cols])
    seg_t = eng.repartition(B200DataFrame(t), PartitionSpec(by="key", algo="hash", num=NUM)).native
    torch.cuda.synchronize()
    lo, hi = owner_range(NUM, world, rank)
    assert seg_t.num_partitions == hi - lo and seg_t.global_partition_range == (lo, hi)
    assert seg_t.offsets is None and tuple(seg_t.segment_offsets.shape) == (world, hi - lo + 1)
    res = seg_t.compacted()
    torch.cuda.synchronize()
    got = [c.cpu().numpy() for c in res.columns]
    seg_cols = [c.cpu().numpy() for c in seg_t.columns]
    off = res.offsets.cpu().numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (got, off, lo, hi, n, seg_cols, seg_t.segment_offsets.numpy()))
    if rank == 0:
        with _guard("block 5"):
            shards = [shard(r, 300_000 + 12345 * r) for r in range(world)]
            glob = [np.concatenate([s[c] for s in shards]) for c in range(4)]
            exp_cols, exp_off = hp.partition_table(glob, [0], NUM)
            per_src = [hp.partition_table(s, [0], NUM) for s in shards]
            total = 0
            for g_cols, g_off, g_lo, g_hi, _, g_seg_cols, g_seg in gathered:
                for j, p in enumerate(range(g_lo, g_hi)):
                    a, b = exp_off[p], exp_off[p + 1]
                    assert g_off[j + 1] - g_off[j] == b - a, (p, g_off[j + 1] - g_off[j], b - a)
                    for c in range(4):
                        assert np.array_equal(g_cols[c][g_off[j]:g_off[j + 1]].view("u1"),
                                              exp_cols[c][a:b].view("u1")), (p, c)
                    total += b - a
                    # the un-compacted result: segment (s, j) = rows of shard s in partition p, source order
                    for s in range(world):
                        sc, so = per_src[s]
                        x, y = g_seg[s, j], g_seg[s, j + 1]
                        assert y - x == so[p + 1] - so[p], (s, p)
                        for c in range(4):
                            assert np.array_equal(g_seg_cols[c][x:y].view("u1"), sc[c][so[p]:so[p + 1]].view("u1")), (s, p, c)
            assert total == sum(x[4] for x in gathered)
            print(f"dist_gpu_check ok: world={world}, {total} rows, bit-exact vs oracle (stable order)")


def check_streaming(eng, rank, world):
    """Host table in, host table out through fa.transform(..., as_local=True) on the distributed engine: the
    pipelined H2D / shuffle / D2H path.  Union over ranks == the input rows; every key lives on one rank."""
    import pyarrow as pa

    from fugue_b200 import api as fa

    n = 200_000 + 777 * rank
    cols = shard(rank, n)[:3]
    tbl = pa.table({"key": cols[0], "v": cols[1], "rid": cols[2]})

    def ident(t: B200Table) -> B200Table:
        return t

    out = fa.transform(tbl, ident, schema="*", partition=PartitionSpec(by="key", algo="hash", num=NUM), engine=eng,
                       as_local=True, as_fugue=True).as_arrow()
    gathered = [None] * world
    dist.all_gather_object(gathered, (out.column("key").to_numpy(), out.column("v").to_numpy(), out.column("rid").to_numpy()))
    if rank == 0:
        with _guard("streaming"):
            shards = [shard(r, 200_000 + 777 * r)[:3] for r in range(world)]
            exp = np.stack([np.concatenate([s_[c] for s_ in shards]).view("i8") for c in range(3)], 1)
            got = np.stack([np.concatenate([g[c] for g in gathered]).view("i8") for c in range(3)], 1)
            assert got.shape == exp.shape
            order_e, order_g = np.lexsort(exp.T[::-1]), np.lexsort(got.T[::-1])
            assert np.array_equal(exp[order_e], got[order_g])
            owner = {}
            for r, g in enumerate(gathered):
                for k in np.unique(g[0]).tolist():
                    assert owner.setdefault(k, r) == r, f"key {k} on ranks {owner[k]} and {r}"
            print(f"dist streaming transform ok: {len(got)} rows")


def run_checks(eng, rank, world, dev):
    """All multi-GPU parity checks; collective (every rank calls it).  Returns (ok, message) -
    meaningful on rank 0, where the comparisons against the oracle run."""
    del _FAILS[:]
    ok, msg = True, ""
    try:
        check_repartition(eng, rank, world, dev)
        check_relational(eng, rank, world, dev)
        check_string_join(eng, rank, world)
        check_streaming(eng, rank, world)
    except Exception as e:  # noqa: BLE001 - reported, not swallowed
        ok, msg = False, repr(e)[:300]
    if _FAILS:
        ok, msg = False, " | ".join(_FAILS)[:600]
    flags = [None] * world
    dist.all_gather_object(flags, (ok, msg))
    bad = [f for f in flags if not f[0]]
    return (len(bad) == 0, bad[0][1] if bad else f"world={world}: repartition (segments + compacted) bit-exact vs "
            "oracle; GROUP BY / JOIN / select / string keys vs pandas; pipelined host-to-host transform")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    eng = DistributedB200Engine({"fugue.b200.device": local_rank})
    ok, msg = run_checks(eng, rank, world, dev)
    if rank == 0:
        with _guard("block 6"):
            print("PARITY", ok, msg)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
