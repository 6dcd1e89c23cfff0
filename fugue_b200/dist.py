"""Multi-GPU hash repartition: one process per GPU, rows range-sharded.

    local K1-K3 into `num` partitions          (libfugue_b200.so)
 -> all-gather of the per-partition counts     (num x world int64)
 -> ONE all-to-all per column over NCCL/NVLink (partition p is owned by rank p*world//num,
                                                so every send region is contiguous)
 -> segment copy: received (source, partition) runs -> (partition, source) order, so every
    owned partition is contiguous and rows keep (source rank, source row) order = stable.

The reference has no shuffle of its own (Dask: set_index + repartition(divisions),
fugue_dask/_utils.py:124-130, 166-169; Spark: df.repartition, fugue_spark/_utils/partition.py:23);
the contract honoured is SURVEY.md 3.3: equal keys land in the same physical partition.

The planning functions work on CPU tensors with any torch.distributed backend, so the
world-size-2 ``gloo`` tests (tests/test_dist_cpu.py) cover this logic without a GPU.
"""
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .dataframe import B200DataFrame
from .execution_engine import B200ExecutionEngine, assert_or_throw
from .partition import PartitionSpec
from .table import B200Table


def owner_range(num: int, world: int, rank: int) -> Tuple[int, int]:
    """Physical partitions [lo, hi) owned by ``rank`` (balanced contiguous ranges)."""
    return rank * num // world, (rank + 1) * num // world


class ExchangePlan:
    """Everything derived from the count matrix ``counts[src][p]``."""

    def __init__(self, counts: torch.Tensor, rank: int):
        world, num = counts.shape
        self.world, self.num, self.rank = world, num, rank
        self.counts = counts
        lo, hi = owner_range(num, world, rank)
        self.lo, self.hi = lo, hi
        bounds = [owner_range(num, world, r) for r in range(world)]
        # rows this rank sends to every destination / receives from every source
        self.send_rows = [int(counts[rank, a:b].sum()) for a, b in bounds]
        self.recv_rows = [int(counts[s, lo:hi].sum()) for s in range(world)]
        self.total_recv = sum(self.recv_rows)
        # received layout: [src 0: partitions lo..hi) | src 1: ... ]; wanted: partition-major
        mine = counts[:, lo:hi]                                   # [world, nown]
        recv_base = torch.zeros(world, dtype=torch.int64)
        recv_base[1:] = torch.cumsum(torch.tensor(self.recv_rows[:-1], dtype=torch.int64), 0)
        within = torch.cumsum(mine, 1) - mine                     # offset of p inside src's run
        src_off = recv_base[:, None] + within                     # [world, nown]
        flat_pm = mine.t().contiguous().reshape(-1)               # partition-major lengths
        dst_off = torch.cumsum(flat_pm, 0) - flat_pm
        self.seg_src_off = src_off.t().contiguous().reshape(-1)   # partition-major order
        # pull exchange: the same runs addressed inside every SOURCE rank's partitioned table
        src_part_off = torch.cumsum(counts, 1) - counts            # [world, num] offset of p in src table
        self.pull_src_off = src_part_off[:, lo:hi].t().contiguous().reshape(-1)
        self.pull_src_rank = torch.arange(world, dtype=torch.int32).repeat(hi - lo)
        self.rows_per_rank = counts.sum(1)                        # local row count of every rank
        self.max_seg = int(flat_pm.max()) if flat_pm.numel() > 0 else 0
        self.seg_dst_off = dst_off
        self.seg_len = flat_pm
        part_counts = mine.sum(0)
        self.out_offsets = torch.zeros(hi - lo + 1, dtype=torch.int64)
        self.out_offsets[1:] = torch.cumsum(part_counts, 0)


def gather_counts(local_counts: torch.Tensor, group: Any = None) -> torch.Tensor:
    """all-gather of the per-partition row counts -> CPU int64 matrix [world, num]."""
    world = dist.get_world_size(group)
    out = torch.empty(world * local_counts.numel(), dtype=torch.int64, device=local_counts.device)
    dist.all_gather_into_tensor(out, local_counts.contiguous(), group=group)
    return out.view(world, -1).cpu()


def exchange_column(col: torch.Tensor, plan: ExchangePlan, group: Any = None) -> torch.Tensor:
    out = torch.empty(plan.total_recv, dtype=col.dtype, device=col.device)
    dist.all_to_all_single(out, col, output_split_sizes=plan.recv_rows, input_split_sizes=plan.send_rows,
                           group=group)
    return out


def rearrange_cpu(recv: Sequence[torch.Tensor], plan: ExchangePlan) -> List[torch.Tensor]:
    """CPU restatement of fb_copy_segments, used by the gloo tests only."""
    outs = [torch.empty_like(c) for c in recv]
    for s, d, n in zip(plan.seg_src_off.tolist(), plan.seg_dst_off.tolist(), plan.seg_len.tolist()):
        for c, o in zip(recv, outs):
            o[d:d + n] = c[s:s + n]
    return outs


class DistributedB200Engine(B200ExecutionEngine):
    """``B200ExecutionEngine`` whose ``repartition`` shuffles across the GPUs of the process group."""

    def __init__(self, conf: Any = None, group: Any = None, **kwargs: Any):
        super().__init__(conf, **kwargs)
        assert_or_throw(dist.is_initialized(), RuntimeError(
            "DistributedB200Engine needs an initialised torch.distributed process group"))
        self._group = group
        self._world = dist.get_world_size(group)
        self._rank = dist.get_rank(group)

    @property
    def is_distributed(self) -> bool:
        return True

    def get_current_parallelism(self) -> int:
        return self._world

    # ---- symmetric arena: this rank's partitioned columns, readable by every peer over NVLink
    def _ensure_arena(self, nbytes: int) -> None:
        """All ranks call this with the same ``nbytes`` (derived from the gathered counts)."""
        import torch.distributed._symmetric_memory as symm_mem

        if getattr(self, "_arena", None) is not None and self._arena.numel() >= nbytes:
            return
        cap = ((int(nbytes * 1.25) + (1 << 21) - 1) >> 21) << 21
        self._arena = symm_mem.empty(cap, dtype=torch.uint8, device=self._device)
        self._arena_hdl = symm_mem.rendezvous(self._arena, group=self._group or dist.group.WORLD)

    @staticmethod
    def _col_offsets(nrows: int, widths: List[int]) -> List[int]:
        off, out = 0, []
        for w in widths:
            out.append(off)
            off += (nrows * w + 255) & ~255
        out.append(off)
        return out

    def repartition(self, df: Any, partition_spec: PartitionSpec) -> B200DataFrame:
        """Shuffle = local K1-K3 into a symmetric arena + a pull over NVLink peer memory:
        pass 1 -> count all-gather -> scatter into the arena -> barrier -> every rank copies the
        (source rank, partition) runs it owns straight from the peers' arenas into its final,
        partition-contiguous output (copy engines, or the fb_copy_segments kernel) -> barrier."""
        from . import kernels as K

        keys = partition_spec.partition_by
        edf = self.to_df(df)
        if len(keys) == 0 or self._world == 1:
            return super().repartition(edf, partition_spec)
        t: B200Table = edf.native
        for k in keys:
            assert_or_throw(k in t.schema, lambda: KeyError(f"{k} not in {t.schema}"))
        num = self._num_partitions(partition_spec, t.num_rows)
        assert_or_throw(num <= K.MAX_PARTITIONS, NotImplementedError(
            f"num_partitions={num}: one radix pass handles up to {K.MAX_PARTITIONS} partitions"))
        assert_or_throw(num >= self._world, ValueError(
            f"num_partitions={num} must be >= the number of GPUs ({self._world})"))
        t = self._globalize_dictionaries(t)  # string columns: one code space on all ranks
        dev = t.device
        kidx = [t.schema.index_of_key(k) for k in keys]
        kvalid = [t.valid[i] for i in kidx]
        cols = list(t.columns)
        vpos: Dict[int, int] = {}
        for i, v in enumerate(t.valid):
            if v is not None:
                vpos[i] = len(cols)
                cols.append(v)
        widths = [c.element_size() for c in cols]
        # ---- pass 1 on the local shard, counts to everybody
        scratch = self._pool.scratch(dev, K.partition_scratch_bytes(dev, t.num_rows, num))
        plan_local = K.partition_plan([t.columns[i] for i in kidx], num, kvalid, scratch=scratch)
        counts = gather_counts(plan_local.offsets[1:] - plan_local.offsets[:-1], self._group)
        plan = ExchangePlan(counts, self._rank)
        mode = os.environ.get("FB_DIST_EXCHANGE", "pull")
        if mode == "nccl":
            return self._repartition_nccl(t, keys, cols, vpos, plan_local, plan)
        rows = [int(x) for x in plan.rows_per_rank.tolist()]
        self._ensure_arena(max(self._col_offsets(r, widths)[-1] for r in rows))
        # ---- pass 2 + exchange in column groups, so that moving group A over NVLink overlaps the
        #      HBM-bound scatter of group B (peers' reads of the previous call are over: every call
        #      ends with a barrier).
        #      "dma" (default): the (source rank, partition) runs are copied by the COPY ENGINES
        #      (fb_copy_runs_dma): the scatter kernel fills every SM's registers and shared memory, so
        #      a pull *kernel* cannot co-run with it and the two serialise; DMA needs no SM.
        #      "pull": one fb_copy_segments kernel with peer pointers (16-byte loads over NVLink).
        my_off = self._col_offsets(t.num_rows, widths)
        parts = [self._arena[my_off[i]:my_off[i] + t.num_rows * w].view(c.dtype)
                 for i, (c, w) in enumerate(zip(cols, widths))]
        base = self._arena_hdl.buffer_ptrs
        peer_off = [self._col_offsets(rows[s], widths) for s in range(self._world)]
        outs = [torch.empty(plan.total_recv, dtype=c.dtype, device=dev) for c in cols]
        s_main = torch.cuda.current_stream(dev)
        if getattr(self, "_pull_stream", None) is None:
            self._pull_stream = torch.cuda.Stream(dev)
        s_pull = self._pull_stream
        per = max(1, int(os.environ.get("FB_DIST_GROUP_COLS", "4")))
        groups = [list(range(a, min(a + per, len(cols)))) for a in range(0, len(cols), per)]
        if mode == "dma":
            import numpy as np

            keep = plan.seg_len.numpy() > 0
            r_rank = plan.pull_src_rank.numpy()[keep].astype(np.int64)
            r_src = plan.pull_src_off.numpy()[keep].astype(np.uint64)
            r_dst = plan.seg_dst_off.numpy()[keep].astype(np.uint64)
            r_len = plan.seg_len.numpy()[keep].astype(np.uint64)
            col_base = np.array([[int(base[s]) + peer_off[s][i] for i in range(len(cols))]
                                 for s in range(self._world)], dtype=np.uint64)   # [rank][column]
        else:
            seg_src, seg_dst = plan.pull_src_off.to(dev), plan.seg_dst_off.to(dev)
            seg_len, seg_rank = plan.seg_len.to(dev), plan.pull_src_rank.to(dev)
        for gi, idx in enumerate(groups):
            K.partition_apply(plan_local, [cols[i] for i in idx], [parts[i] for i in idx])
            ev = torch.cuda.Event()
            ev.record(s_main)
            with torch.cuda.stream(s_pull):
                s_pull.wait_event(ev)
                dist.barrier(group=self._group)  # stream-ordered: this group is complete on all ranks
                if mode == "dma":
                    src = np.concatenate([col_base[r_rank, i] + r_src * np.uint64(widths[i]) for i in idx])
                    dst = np.concatenate([np.uint64(outs[i].data_ptr()) + r_dst * np.uint64(widths[i]) for i in idx])
                    nb = np.concatenate([r_len * np.uint64(widths[i]) for i in idx])
                    K.copy_runs_dma(dev, src, dst, nb)
                else:
                    src_ptrs = [int(base[s]) + peer_off[s][i] for s in range(self._world) for i in idx]
                    K.copy_segments(None, [outs[i] for i in idx], seg_src, seg_dst, seg_len, max_len=plan.max_seg,
                                    src_table=seg_rank, src_ptrs=src_ptrs)
        with torch.cuda.stream(s_pull):
            dist.barrier(group=self._group)  # nobody overwrites an arena that is still being read
        s_main.wait_stream(s_pull)
        for o in outs:
            o.record_stream(s_pull)
        ncol = len(t.columns)
        valid = [outs[vpos[i]] if i in vpos else None for i in range(ncol)]
        res = B200Table(t.schema, outs[:ncol], valid, t.dictionaries, plan.out_offsets.to(dev), list(keys))
        res.global_partition_range = (plan.lo, plan.hi)  # which physical partitions this GPU owns
        return B200DataFrame(res)

    def _globalize_dictionaries(self, t: B200Table) -> B200Table:
        """String columns are dictionary encoded per rank; before rows travel (and before codes are
        hashed as partition keys) every rank re-codes them against the union dictionary: the
        dictionaries are gathered on the host (small), the union is taken in first-appearance order
        over ranks (identical everywhere), and the codes are remapped on the device."""
        if len(t.dictionaries) == 0:
            return t
        import pyarrow as pa
        import pyarrow.compute as pc

        names = sorted(t.dictionaries)
        local = {k: t.dictionaries[k].to_pylist() for k in names}
        gathered: List[Any] = [None] * self._world
        dist.all_gather_object(gathered, local, group=self._group)
        cols = list(t.columns)
        dicts: Dict[str, Any] = {}
        for k in names:
            union = pc.unique(pa.array([x for g in gathered for x in g[k]], type=pa.string()))
            pos = pc.index_in(t.dictionaries[k], value_set=union).to_numpy(zero_copy_only=False).astype("int32")
            i = t.schema.index_of_key(k)
            if len(pos) > 0:
                m = torch.from_numpy(pos).to(t.device)
                cols[i] = m[cols[i].long().clamp_(min=0)].contiguous()
            dicts[k] = union
        return B200Table(t.schema, cols, t.valid, dicts, t.offsets, t.partition_keys)

    def _repartition_nccl(self, t: B200Table, keys: List[str], cols: List[torch.Tensor],
                          vpos: Dict[int, int], plan_local: Any, plan: ExchangePlan) -> B200DataFrame:
        """Baseline exchange (FB_DIST_EXCHANGE=nccl): one NCCL all-to-all per column + local
        segment copy.  Kept for comparison; the pull kernel above is the product path."""
        from . import kernels as K

        dev = t.device
        parts = K.partition_apply(plan_local, cols)
        recv = [exchange_column(c, plan, self._group) for c in parts]
        outs = [torch.empty_like(c) for c in recv]
        K.copy_segments(recv, outs, plan.seg_src_off.to(dev), plan.seg_dst_off.to(dev), plan.seg_len.to(dev),
                        max_len=plan.max_seg)
        ncol = len(t.columns)
        valid = [outs[vpos[i]] if i in vpos else None for i in range(ncol)]
        res = B200Table(t.schema, outs[:ncol], valid, t.dictionaries, plan.out_offsets.to(dev), list(keys))
        res.global_partition_range = (plan.lo, plan.hi)
        return B200DataFrame(res)

    # ---- distributed relational operators (BASELINE configs 4 and 5) -------------------------
    def _shuffle_partitions(self) -> int:
        from .execution_engine import FUGUE_B200_CONF_DEFAULT_PARTITIONS, FUGUE_B200_DEFAULT_PARTITIONS

        n = int(self._conf.get(FUGUE_B200_CONF_DEFAULT_PARTITIONS, FUGUE_B200_DEFAULT_PARTITIONS))
        return max(n, self._world)

    def aggregate(self, df: Any, partition_spec: Optional[PartitionSpec], agg_cols: List[Any]) -> B200DataFrame:
        """GROUP BY across GPUs: local partial aggregation (K6) -> shuffle of the partials by key
        (pull exchange) -> final aggregation of the partials.  Every group ends up on exactly one
        rank; the result stays sharded.  SUM/COUNT/MIN/MAX decompose directly, AVG as SUM + COUNT."""
        from .column import AggFuncExpr, col

        keys = [] if partition_spec is None else list(partition_spec.partition_by)
        if self._world == 1 or not self._plain_aggs(agg_cols):
            # aggregations of expressions / expressions of aggregations: the base class evaluates the
            # row-wise parts locally and comes back here with plain FUNC(column) aggregations
            return super().aggregate(df, partition_spec, agg_cols)
        from .execution_engine import decompose_aggs, finish_avgs

        partial, final, post = decompose_aggs(agg_cols)
        local = super().aggregate(df, partition_spec, partial)
        if len(keys) == 0:
            # global aggregate: every rank reduces its partial row; gather the W partial rows everywhere
            shuffled = self._allgather_rows(local)
        else:
            shuffled = self.repartition(local, PartitionSpec(by=keys, num=self._shuffle_partitions()))
        res = super().aggregate(shuffled, partition_spec, final)
        return finish_avgs(res, post, keys + [a.output_name for a in agg_cols])

    def _allgather_rows(self, df: B200DataFrame) -> B200DataFrame:
        t: B200Table = df.native
        cols, valid = [], []
        for c, v in zip(t.columns, t.valid):
            out = torch.empty(c.shape[0] * self._world, dtype=c.dtype, device=c.device)
            dist.all_gather_into_tensor(out, c.contiguous(), group=self._group)
            cols.append(out)
            if v is None:
                v = torch.ones(c.shape[0], dtype=torch.uint8, device=c.device)
            vo = torch.empty(c.shape[0] * self._world, dtype=torch.uint8, device=c.device)
            dist.all_gather_into_tensor(vo, v.contiguous(), group=self._group)
            valid.append(vo)
        return B200DataFrame(B200Table(t.schema, cols, valid, t.dictionaries))

    def join(self, df1: Any, df2: Any, how: str, on: Optional[List[str]] = None) -> B200DataFrame:
        """Equi-join across GPUs: co-partition both sides on the join keys with the same hash and the
        same partition -> rank ownership (two pull exchanges), then join locally (K7).  Rows with equal
        keys of both tables are on the same rank; NULL keys are co-located too (they never match, but
        outer joins must still emit them exactly once).  Cross joins are not distributed."""
        from .join import get_join_schemas

        if self._world == 1:
            return super().join(df1, df2, how, on)
        e1, e2 = self.to_df(df1), self.to_df(df2)
        key_schema, _ = get_join_schemas(e1, e2, how, on)
        assert_or_throw(how.lower() != "cross", NotImplementedError("distributed cross join"))
        spec = PartitionSpec(by=key_schema.names, num=self._shuffle_partitions())
        return super().join(self.repartition(e1, spec), self.repartition(e2, spec), how, on)
