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

    def repartition(self, df: Any, partition_spec: PartitionSpec) -> B200DataFrame:
        from . import kernels as K

        local = super().repartition(df, partition_spec)  # K1-K3 on this GPU's row shard
        keys = partition_spec.partition_by
        if len(keys) == 0 or self._world == 1:
            return local
        t: B200Table = local.native
        num = t.num_partitions
        assert_or_throw(num >= self._world, ValueError(
            f"num_partitions={num} must be >= the number of GPUs ({self._world})"))
        assert_or_throw(len(t.dictionaries) == 0, NotImplementedError(
            "string (dictionary-encoded) columns need a global dictionary before a multi-GPU shuffle"))
        counts = gather_counts(t.offsets[1:] - t.offsets[:-1], self._group)
        plan = ExchangePlan(counts, self._rank)
        cols = list(t.columns)
        vpos: Dict[int, int] = {}
        for i, v in enumerate(t.valid):
            if v is not None:
                vpos[i] = len(cols)
                cols.append(v)
        recv = [exchange_column(c, plan, self._group) for c in cols]  # NCCL all-to-all per column
        outs = [torch.empty_like(c) for c in recv]
        dev = t.device
        K.copy_segments(recv, outs, plan.seg_src_off.to(dev), plan.seg_dst_off.to(dev), plan.seg_len.to(dev))
        ncol = len(t.columns)
        valid = [outs[vpos[i]] if i in vpos else None for i in range(ncol)]
        res = B200Table(t.schema, outs[:ncol], valid, t.dictionaries, plan.out_offsets.to(dev), list(keys))
        res.global_partition_range = (plan.lo, plan.hi)  # which physical partitions this GPU owns
        return B200DataFrame(res)
