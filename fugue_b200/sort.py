"""Device sort (LSD radix passes of the partition kernels), logical partitions, ``take``.

Reference semantics:
  * presort inside a partition   fugue/execution/native_execution_engine.py:107-115, 157-160
                                 (``pdf.sort_values(presort_keys, ascending=...)``)
  * ``take``                     fugue/execution/native_execution_engine.py:350-384
                                 (sort with ``na_position`` then ``head(n)`` per group; NULL keys form
                                 a group: ``groupby(dropna=False)``)
  * logical partitions           one per distinct key tuple, NULLs grouped

A column becomes an order-preserving UNSIGNED 64-bit key (sign bit flipped for signed ints, the
usual total-order transform for floats, dictionary rank for strings, complement for DESC); (key,
row index) pairs are sorted with one stable 8-bit radix pass (``fb_radix_pass``) per varying byte, the
least significant sort column first; NULLS FIRST/LAST is one more pass on the validity flag.  The
payload is gathered once at the end (``fb_gather_rows``).  Key preparation uses torch integer ops on
the device; the passes and the gather are the library's kernels.
"""
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import torch

from . import _lib
from . import kernels as K
from .table import B200Table

_SIGN = -(1 << 63)


def _unsigned_order_key(t: B200Table, name: str, ascending: bool) -> torch.Tensor:
    """int64 tensor whose bit pattern, read as unsigned, orders like the column."""
    i = t.schema.index_of_key(name)
    c, tp = t.columns[i], t.schema.types[i]
    if name in t.dictionaries:  # strings: rank of every dictionary entry in sorted order
        d = t.dictionaries[name]
        order = pc.sort_indices(d).to_numpy()
        rank = np.empty(len(d), dtype=np.int64)
        rank[order] = np.arange(len(d), dtype=np.int64)
        r = torch.from_numpy(rank).to(c.device)
        key = r[c.long().clamp(min=0)] if len(d) > 0 else torch.zeros_like(c, dtype=torch.int64)
    elif pa.types.is_floating(tp):
        b = (c if c.dtype == torch.float64 else c.to(torch.float64)).view(torch.int64)
        key = b ^ ((b >> 63) | _SIGN)  # negative: flip all bits; non-negative: flip the sign bit
    elif tp in (pa.uint8(), pa.bool_()):
        key = c.to(torch.int64)
    elif tp == pa.uint16():
        key = c.to(torch.int64) & 0xFFFF
    elif tp == pa.uint32():
        key = c.to(torch.int64) & 0xFFFFFFFF
    elif tp == pa.uint64():
        key = c  # the int64 bit view already is the unsigned value
    else:  # signed integers, dates, timestamps
        key = c.to(torch.int64) ^ _SIGN
    if not ascending:
        key = ~key
    return key.contiguous()


def _radix_sort_pairs(key: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable sort of (key, idx) pairs by the unsigned value of key; only varying bytes get a pass."""
    lib = _lib.load()
    dev = key.device
    n = int(key.shape[0])
    if n <= 1:
        return key, idx
    lo, hi = int(key.min().item()), int(key.max().item())
    diff = (lo ^ hi) & 0xFFFFFFFFFFFFFFFF
    # signed min/max are not the unsigned extremes when signs differ: then every byte may vary
    if (lo < 0) != (hi < 0):
        diff = 0xFFFFFFFFFFFFFFFF
    nbytes = (diff.bit_length() + 7) // 8
    scratch = torch.empty(K.partition_scratch_bytes(dev, n, 256) + 256, dtype=torch.uint8, device=dev)
    offsets = torch.empty(257, dtype=torch.int64, device=dev)
    k2, i2 = torch.empty_like(key), torch.empty_like(idx)
    widths = _lib.i32_array([8, 8])
    for b in range(nbytes):
        _lib.check(lib.fb_radix_pass(dev.index, K._stream_ptr(dev), n, key.data_ptr(), 8 * b, 2,
                                     _lib.ptr_array([key.data_ptr(), idx.data_ptr()]), widths,
                                     _lib.ptr_array([k2.data_ptr(), i2.data_ptr()]), scratch.data_ptr(),
                                     scratch.numel(), offsets.data_ptr()))
        key, k2 = k2, key
        idx, i2 = i2, idx
    return key, idx


def argsort_rows(t: B200Table, sorts: "OrderedDict[str, bool]", na_position: str = "last") -> torch.Tensor:
    """Row permutation that sorts the table by ``sorts`` (name -> ascending), stable."""
    if na_position not in ("first", "last"):
        raise ValueError(f"invalid na_position {na_position}")
    dev = t.device
    n = t.num_rows
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    for name, asc in reversed(list(sorts.items())):
        key = _unsigned_order_key(t, name, asc)
        v = t.valid[t.schema.index_of_key(name)]
        if v is not None:
            # the value stored under a NULL is undefined (Arrow / parquet leave garbage there): give all
            # NULL rows one constant key, so that they keep the order set by the less significant columns
            key = torch.where(v.bool(), key, torch.zeros_like(key))
        _, idx = _radix_sort_pairs(key[idx].contiguous(), idx)
        if v is not None:
            flag = v.to(torch.int64) if na_position == "first" else (1 - v.to(torch.int64))
            _, idx = _radix_sort_pairs(flag[idx].contiguous(), idx)
    return idx


def take_rows(t: B200Table, idx: torch.Tensor) -> B200Table:
    cols, valid = K.gather_rows(t.columns, t.valid, idx.contiguous(), want_valid=False)
    return B200Table(t.schema, cols, valid, t.dictionaries)


def sort_table(t: B200Table, sorts: "OrderedDict[str, bool]", na_position: str = "last") -> B200Table:
    if len(sorts) == 0 or t.num_rows <= 1:
        return t
    return take_rows(t, argsort_rows(t, sorts, na_position))


def group_starts(t: B200Table, keys: List[str]) -> torch.Tensor:
    """For a table in which equal key tuples are adjacent: bool mask, True at the first row of
    every logical partition (NULL == NULL for grouping)."""
    n = t.num_rows
    first = torch.zeros(n, dtype=torch.bool, device=t.device)
    if n == 0:
        return first
    first[0] = True
    for k in keys:
        i = t.schema.index_of_key(k)
        c, v = t.columns[i], t.valid[i]
        if c.dtype in (torch.float32, torch.float64):
            c = c.view(torch.int32 if c.dtype == torch.float32 else torch.int64)
        if v is None:
            diff = c[1:] != c[:-1]
        else:
            diff = (v[1:] != v[:-1]) | ((v[1:] != 0) & (c[1:] != c[:-1]))
        first[1:] |= diff
    return first


def logical_offsets(t: B200Table, keys: List[str]) -> torch.Tensor:
    """int64 offsets (length groups + 1) of the logical partitions of a key-sorted table."""
    starts = K.compact_indices((group_starts(t, keys)).contiguous())
    end = torch.tensor([t.num_rows], dtype=torch.int64, device=t.device)
    return torch.cat([starts, end])


def take(t: B200Table, n: int, sorts: "OrderedDict[str, bool]", na_position: str,
         partition_by: List[str]) -> B200Table:
    """First ``n`` rows (per logical partition when ``partition_by`` is given) after sorting."""
    if not isinstance(n, int) or isinstance(n, bool):
        raise ValueError("n needs to be an integer")
    if len(partition_by) == 0:
        s = sort_table(t, sorts, na_position)
        return s.slice(0, min(n, s.num_rows))
    full: "OrderedDict[str, bool]" = OrderedDict((k, True) for k in partition_by)
    for k, v in sorts.items():
        if k not in full:
            full[k] = v
    # group columns first (any consistent order groups equal keys; NULL keys last), presort inside
    idx = argsort_rows(t, OrderedDict((k, v) for k, v in full.items()), na_position)
    if any(k in sorts for k in partition_by):  # a partition key re-listed in presort keeps its own direction
        full2 = OrderedDict((k, sorts.get(k, True)) for k in partition_by)
        for k, v in sorts.items():
            if k not in full2:
                full2[k] = v
        idx = argsort_rows(t, full2, na_position)
    s = take_rows(t, idx)
    first = group_starts(s, partition_by)
    pos = torch.arange(s.num_rows, dtype=torch.int64, device=s.device)
    start_of = torch.cummax(torch.where(first, pos, torch.zeros_like(pos)), 0).values
    keep = K.compact_indices(((pos - start_of) < n).contiguous())
    return take_rows(s, keep)
