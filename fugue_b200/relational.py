"""Set operations, NULL handling, sampling and IO of ``B200ExecutionEngine``.

Reference semantics: fugue/execution/native_execution_engine.py:243-412
(union / subtract / intersect / distinct / dropna / fillna / sample / load_df / save_df) and the
literal expectations of fugue_test/execution_suite.py:545-760.  DISTINCT-style operations treat
NULLs as equal (SQL set semantics), unlike joins.  They are built on the device group-by (K6): the
rows of both inputs are tagged, grouped on ALL columns, and a group survives depending on the
MIN/MAX of its tags.  Row masks and compaction indices are torch tensor ops on the device; the
gathers go through ``fb_gather_rows``.
"""
import os
from typing import Any, Dict, List, Optional

import pyarrow as pa
import pyarrow.compute as pc
import torch

from . import kernels as K
from .column import agg as _agg, col
from .dataframe import B200DataFrame
from .partition import PartitionSpec
from .schema import Schema
from .table import B200Table

_TAG = "__fb_tag"


def _check_same_schema(df1: B200DataFrame, df2: B200DataFrame) -> None:
    if df1.schema != df2.schema:
        raise ValueError(f"{df1.schema} != {df2.schema}")


def concat_tables(t1: B200Table, t2: B200Table) -> B200Table:
    """UNION ALL of two device tables with equal schemas (string dictionaries are unified)."""
    cols, valid, dicts = [], [], {}
    n1, n2 = t1.num_rows, t2.num_rows
    for i, name in enumerate(t1.schema.names):
        c1, c2 = t1.columns[i], t2.columns[i]
        if name in t1.dictionaries or name in t2.dictionaries:
            d1, d2 = t1.dictionaries[name], t2.dictionaries[name]
            merged = pa.concat_arrays([d1, d2.filter(pc.invert(pc.is_in(d2, value_set=d1)))])
            pos = pc.index_in(d2, value_set=merged)
            m = torch.from_numpy(pos.to_numpy(zero_copy_only=False).astype("int32")).to(c2.device)
            c2 = m[c2.long().clamp(min=0)] if n2 > 0 and len(d2) > 0 else c2
            dicts[name] = merged
        cols.append(torch.cat([c1, c2]))
        v1, v2 = t1.valid[i], t2.valid[i]
        if v1 is None and v2 is None:
            valid.append(None)
        else:
            one = lambda n: torch.ones(n, dtype=torch.uint8, device=c1.device)  # noqa: E731
            valid.append(torch.cat([v1 if v1 is not None else one(n1), v2 if v2 is not None else one(n2)]))
    return B200Table(t1.schema, cols, valid, dicts)


def _take_rows(t: B200Table, idx: torch.Tensor) -> B200Table:
    cols, valid = K.gather_rows(t.columns, t.valid, idx.contiguous(), want_valid=False)
    return B200Table(t.schema, cols, valid, t.dictionaries)


def distinct(engine: Any, df: B200DataFrame) -> B200DataFrame:
    t: B200Table = df.native
    if t.num_rows == 0:
        return df
    res = engine.aggregate(df, PartitionSpec(by=t.schema.names), [_agg("COUNT", col("*"), "__fb_n")])
    return res[t.schema.names]


def union(engine: Any, df1: B200DataFrame, df2: B200DataFrame, distinct_: bool = True) -> B200DataFrame:
    _check_same_schema(df1, df2)
    both = B200DataFrame(concat_tables(df1.native, df2.native))
    return distinct(engine, both) if distinct_ else both


def _tagged_groups(engine: Any, df1: B200DataFrame, df2: B200DataFrame) -> B200DataFrame:
    """Groups of identical rows over both inputs with MIN/MAX of the side tag (1 = df1, 2 = df2)."""
    t1, t2 = df1.native, df2.native
    dev = t1.device

    def tag(t: B200Table, v: int) -> B200Table:
        tagcol = torch.full((t.num_rows,), v, dtype=torch.int64, device=dev)
        return B200Table(Schema(t.schema, f"{_TAG}:long"), list(t.columns) + [tagcol], list(t.valid) + [None],
                         t.dictionaries)

    both = B200DataFrame(concat_tables(tag(t1, 1), tag(t2, 2)))
    return engine.aggregate(both, PartitionSpec(by=t1.schema.names),
                            [_agg("MIN", col(_TAG), "__fb_lo"), _agg("MAX", col(_TAG), "__fb_hi")])


def _keep_groups(groups: B200DataFrame, names: List[str], lo: int, hi: int) -> B200DataFrame:
    t: B200Table = groups.native
    keep = (t.column("__fb_lo") == lo) & (t.column("__fb_hi") == hi)
    idx = K.compact_indices((keep).contiguous())
    return B200DataFrame(_take_rows(t, idx))[names]


def subtract(engine: Any, df1: B200DataFrame, df2: B200DataFrame, distinct_: bool = True) -> B200DataFrame:
    if not distinct_:
        raise NotImplementedError("EXCEPT ALL")
    _check_same_schema(df1, df2)
    if df1.native.num_rows == 0:
        return df1
    return _keep_groups(_tagged_groups(engine, df1, df2), df1.schema.names, 1, 1)


def intersect(engine: Any, df1: B200DataFrame, df2: B200DataFrame, distinct_: bool = True) -> B200DataFrame:
    if not distinct_:
        raise NotImplementedError("INTERSECT ALL")
    _check_same_schema(df1, df2)
    if df1.native.num_rows == 0:
        return df1
    return _keep_groups(_tagged_groups(engine, df1, df2), df1.schema.names, 1, 2)


def dropna(df: B200DataFrame, how: str = "any", thresh: Optional[int] = None,
           subset: Optional[List[str]] = None) -> B200DataFrame:
    t: B200Table = df.native
    names = subset or t.schema.names
    if how not in ("any", "all"):
        raise ValueError(f"{how} is not one of any and all")
    nn = torch.zeros(t.num_rows, dtype=torch.int32, device=t.device)
    for n in names:
        v = t.valid[t.schema.index_of_key(n)]
        nn += 1 if v is None else v.to(torch.int32)
    need = thresh if thresh is not None else (len(names) if how == "any" else 1)
    idx = K.compact_indices((nn >= need).contiguous())
    if idx.numel() == t.num_rows:
        return df
    return B200DataFrame(_take_rows(t, idx))


def fillna(df: B200DataFrame, value: Any, subset: Optional[List[str]] = None) -> B200DataFrame:
    if isinstance(value, list) or value is None:
        raise ValueError("fillna value can not None or a list")
    t: B200Table = df.native
    if isinstance(value, dict):
        if None in value.values() or not any(v is not None for v in value.values()):
            raise ValueError("fillna dict can not contain None and needs at least one value")
        mapping = value
    else:
        mapping = {c: value for c in (subset or t.schema.names)}
    cols, valid, dicts = list(t.columns), list(t.valid), dict(t.dictionaries)
    for name, val in mapping.items():
        i = t.schema.index_of_key(name)
        if valid[i] is None:
            continue
        if name in dicts:
            d = dicts[name]
            pos = pc.index_in(pa.array([str(val)], type=d.type), value_set=d)[0].as_py()
            if pos is None:
                dicts[name] = pa.concat_arrays([d, pa.array([str(val)], type=d.type)])
                pos = len(d)
            fill = torch.tensor(pos, dtype=cols[i].dtype, device=t.device)
        else:
            fill = torch.tensor(val, device=t.device).to(cols[i].dtype)
        cols[i] = torch.where(valid[i] != 0, cols[i], fill)
        valid[i] = None
    return B200DataFrame(B200Table(t.schema, cols, valid, dicts))


def sample(df: B200DataFrame, n: Optional[int] = None, frac: Optional[float] = None, replace: bool = False,
           seed: Optional[int] = None) -> B200DataFrame:
    if (n is None) == (frac is None):
        raise ValueError("one and only one of n and frac should be set")
    t: B200Table = df.native
    g = torch.Generator(device=t.device)
    if seed is not None:
        g.manual_seed(int(seed))
    else:
        g.seed()
    k = int(n) if n is not None else int(round(t.num_rows * float(frac)))
    if replace:
        idx = torch.randint(0, max(t.num_rows, 1), (k,), dtype=torch.int64, device=t.device, generator=g)
    else:
        k = min(k, t.num_rows)
        idx = torch.randperm(t.num_rows, dtype=torch.int64, device=t.device, generator=g)[:k].sort().values
    return B200DataFrame(_take_rows(t, idx))


def load_df(engine: Any, path: Any, format_hint: Any = None, columns: Any = None, **kwargs: Any) -> B200DataFrame:
    """Host IO through pyarrow (fugue_b200/io.py: files, folders, patterns; parquet / csv / json), then one H2D
    per column (``ExecutionEngine.load_df`` execution_engine.py:1127-1146, fugue/_utils/io.py:107-147)."""
    from . import io as IO

    return engine.to_df(IO.load_df(path, format_hint, columns, **kwargs))


def save_df(engine: Any, df: Any, path: str, format_hint: Any = None, mode: str = "overwrite",
            **kwargs: Any) -> None:
    """One D2H (``as_local_bounded``), then the host writer (execution_engine.py:1148-1174)."""
    from . import io as IO

    IO.save_df(engine.to_df(df).as_local_bounded(), path, format_hint, mode, **kwargs)
