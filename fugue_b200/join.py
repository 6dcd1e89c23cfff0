"""Device join driver for ``B200ExecutionEngine.join`` (kernels: csrc/fb_join.cu).

Reference semantics:
  * schema rule ``get_join_schemas``   fugue/dataframe/utils.py:152-226
  * join types / NULL keys             fugue/execution/native_execution_engine.py:230-241,
                                       fugue_test/execution_suite.py:366-543
"""
from typing import Any, List, Optional, Tuple

import pyarrow as pa
import pyarrow.compute as pc
import torch

from . import kernels as K
from .dataframe import B200DataFrame
from .schema import Schema, SchemaError
from .table import B200Table

_JOIN_TYPES = ["semi", "left_semi", "anti", "left_anti", "inner", "left_outer", "right_outer",
               "full_outer", "cross"]


def get_join_schemas(df1: Any, df2: Any, how: str, on: Optional[List[str]]) -> Tuple[Schema, Schema]:
    """(key schema, output schema) - fugue/dataframe/utils.py:152-226."""
    assert how is not None, "how can't be None"
    how = how.lower()
    if how not in _JOIN_TYPES:
        if how == "outer":
            raise ValueError("'how' must use left_outer, right_outer, full_outer for outer joins")
        raise ValueError(f"{how} is not a valid join type")
    on = list(on) if on is not None else []
    if len(on) != len(set(on)):
        raise AssertionError(f"{on} has duplication")
    if how != "cross" and len(on) == 0:
        other = set(df2.columns)
        on = [c for c in df1.columns if c in other]
        if len(on) == 0:
            raise SchemaError(f"no common columns between {df1.columns} and {df2.columns}")
    schema2 = df2.schema
    if how in ["semi", "left_semi", "anti", "left_anti"]:
        schema2 = schema2.extract(on)
    if not (on in df1.schema and on in schema2):
        raise SchemaError(f"{on} is not the intersection of {df1.schema} & {df2.schema}")
    for k in on:
        if df1.schema[k].type != schema2[k].type:
            raise SchemaError(f"join key {k} has different types: {df1.schema[k].type} vs {schema2[k].type}")
    cm = df1.schema.intersect(on)
    if how == "cross":
        cs = df1.schema.intersect(schema2.names)
        if len(cs) > 0:
            raise SchemaError(f"invalid cross join, two dataframes have common columns {cs}")
    elif len(on) == 0:
        raise SchemaError("join on columns must be specified")
    return cm, df1.schema.union(schema2)


def _key64(t1: B200Table, t2: B200Table, keys: List[str]):
    """One 8-byte surrogate key per row on both sides + validity (NULL in any key -> never matches).
    exact == False means the surrogate is a hash and matches must be verified."""
    def norm(c: torch.Tensor) -> torch.Tensor:
        if c.dtype in (torch.float32, torch.float64):
            return torch.where(c == 0, torch.zeros_like(c), c)  # -0.0 == 0.0
        return c

    def cols_of(t: B200Table, remap=None):
        out, val = [], None
        for k in keys:
            i = t.schema.index_of_key(k)
            c = norm(t.columns[i])
            if remap is not None and k in remap:
                m = remap[k]
                c = m[c.long().clamp(min=0)]
                bad = (c < 0).to(torch.uint8) ^ 1
                val = bad if val is None else val & bad
            out.append(c.contiguous())
            if t.valid[i] is not None:
                val = t.valid[i] if val is None else val & t.valid[i]
        return out, val

    # string keys: translate the right side's dictionary codes into the left side's code space
    remap = {}
    for k in keys:
        if k in t1.dictionaries or k in t2.dictionaries:
            d1, d2 = t1.dictionaries[k], t2.dictionaries[k]
            pos = pc.index_in(d2, value_set=d1).fill_null(-1)
            remap[k] = torch.from_numpy(pos.to_numpy(zero_copy_only=False).astype("int32")).to(t2.device)
    c1, v1 = cols_of(t1)
    c2, v2 = cols_of(t2, remap)
    if len(keys) == 1 and c1[0].element_size() == 8:
        return c1[0].view(torch.int64), v1, c2[0].view(torch.int64), v2, True
    if len(keys) == 1:
        def widen(c):
            if c.dtype == torch.float32:
                c = c.view(torch.int32)
            return c.to(torch.int64)
        return widen(c1[0]), v1, widen(c2[0]), v2, True
    return K.row_hash64(c1), v1, K.row_hash64(c2), v2, False


def _verify(t1: B200Table, t2: B200Table, keys: List[str], li: torch.Tensor, ri: torch.Tensor):
    """Drop hash-collision candidates of a multi-column key (both indices >= 0)."""
    ok = torch.ones(li.shape[0], dtype=torch.bool, device=li.device)
    for k in keys:
        a = t1.column(k)[li]
        b = t2.column(k)[ri]
        if k in t1.dictionaries:  # compare decoded codes through the left dictionary
            d1, d2 = t1.dictionaries[k], t2.dictionaries[k]
            pos = pc.index_in(d2, value_set=d1).fill_null(-1)
            m = torch.from_numpy(pos.to_numpy(zero_copy_only=False).astype("int32")).to(li.device)
            b = m[b.long()]
        ok &= a == b
    return ok


RADIX_JOIN_MIN_ROWS = 2_000_000
RADIX_JOIN_PARTITIONS = 256


def _radix_partition(t: B200Table, key64: torch.Tensor, kvalid: Optional[torch.Tensor], num: int):
    """Hash-partition the surrogate key together with every column (and mask) of the table, so
    that build, probe and the output gathers all walk the data partition by partition
    (radix join: the hash table and the gathered rows stay L2-resident)."""
    cols: List[torch.Tensor] = [key64]
    index = {}
    for i, c in enumerate(t.columns):
        index[("c", i)] = len(cols)
        cols.append(c)
    for i, v in enumerate(t.valid):
        if v is not None:
            index[("v", i)] = len(cols)
            cols.append(v)
    if kvalid is not None:
        index["kv"] = len(cols)
        cols.append(kvalid)
    out, part_offsets = K.partition_columns(cols, [0], num, [kvalid])
    pt = B200Table(t.schema, [out[index[("c", i)]] for i in range(len(t.columns))],
                   [out[index[("v", i)]] if ("v", i) in index else None for i in range(len(t.columns))],
                   t.dictionaries)
    return pt, out[0], (out[index["kv"]] if kvalid is not None else None), part_offsets


def device_join(engine: Any, df1: B200DataFrame, df2: B200DataFrame, how: str,
                on: Optional[List[str]]) -> B200DataFrame:
    key_schema, out_schema = get_join_schemas(df1, df2, how, on)
    how = how.lower()
    keys = key_schema.names
    t1, t2 = df1.native, df2.native
    dev = t1.device
    n1, n2 = t1.num_rows, t2.num_rows
    if how == "cross":
        total = n1 * n2
        o = torch.arange(total, dtype=torch.int64, device=dev)
        li, ri = (o // max(n2, 1)), (o % max(n2, 1))
        return _assemble(t1, t2, keys, out_schema, li, ri, how)
    k1, v1, k2, v2, exact = _key64(t1, t2, keys)
    if getattr(t1, "global_num_partitions", None) or getattr(t2, "global_num_partitions", None):
        k1, k2 = K.scramble64(k1), K.scramble64(k2)  # shuffled input: see kernels.scramble64
    parts = 0
    po1 = po2 = None
    if min(n1, n2) >= RADIX_JOIN_MIN_ROWS:
        parts = RADIX_JOIN_PARTITIONS
        t1, k1, v1, po1 = _radix_partition(t1, k1, v1, parts)
        t2, k2, v2, po2 = _radix_partition(t2, k2, v2, parts)
    if how in ("semi", "left_semi", "anti", "left_anti"):
        if exact:
            counts = K.JoinTable(k2, v2, parts, po2).probe_counts(k1, v1, outer=False)
            hit = counts > 0
        else:
            li, ri = K.JoinTable(k2, v2, parts, po2).probe(k1, v1, outer=False)
            ok = _verify(t1, t2, keys, li, ri)
            hit = torch.zeros(n1, dtype=torch.bool, device=dev)
            hit[li[ok]] = True
        keep = hit if how in ("semi", "left_semi") else ~hit
        li = K.compact_indices((keep).contiguous())
        cols, valid = K.gather_rows(t1.columns, t1.valid, li, want_valid=False)
        return B200DataFrame(B200Table(out_schema, cols, valid, t1.dictionaries))
    if exact and how in ("inner", "left_outer") and n2 < (1 << 31) \
            and len(t1.columns) + sum(v is not None for v in t1.valid) <= K.JOIN2_MAX_COLS \
            and len(t2.columns) <= K.JOIN2_MAX_COLS:
        return _fused_join(t1, t2, keys, out_schema, k1, v1, k2, v2, parts, po2, how == "left_outer", po1)
    if how == "right_outer":
        # probe with the right side so that every right row appears
        tab = K.JoinTable(k1, v1, parts, po1)
        ri, li = tab.probe(k2, v2, outer=True)
        if not exact:
            li, ri = _drop_collisions(t1, t2, keys, li, ri, outer_side="right")
        return _assemble(t1, t2, keys, out_schema, li, ri, how)
    tab = K.JoinTable(k2, v2, parts, po2)
    li, ri = tab.probe(k1, v1, outer=how in ("left_outer", "full_outer"))
    if not exact:
        li, ri = _drop_collisions(t1, t2, keys, li, ri, outer_side="left" if how != "inner" else None)
    if how == "full_outer":
        matched = tab.matched_mask(ri) if exact else _matched_mask(n2, ri)
        extra = K.compact_indices((matched == 0).contiguous())
        li = torch.cat([li, torch.full_like(extra, -1)])
        ri = torch.cat([ri, extra])
    return _assemble(t1, t2, keys, out_schema, li, ri, how)


def _fused_join(t1: B200Table, t2: B200Table, keys: List[str], out_schema: Schema, k1: torch.Tensor,
                v1: Optional[torch.Tensor], k2: torch.Tensor, v2: Optional[torch.Tensor], parts: int,
                po2: Optional[torch.Tensor], outer: bool, po1: Optional[torch.Tensor] = None) -> B200DataFrame:
    """Inner / left outer join on one 8-byte key through the fused kernels (K.join_fused): the probe
    side's columns (and validity masks, as 1-byte columns) are copied, the build side's non-key columns
    are gathered - straight into the output table."""
    left: List[torch.Tensor] = list(t1.columns)
    lmask_at = {}
    for i, v in enumerate(t1.valid):
        if v is not None:
            lmask_at[i] = len(left)
            left.append(v)
    names2 = [n for n in t2.schema.names if n not in keys]
    idx2 = [t2.schema.index_of_key(n) for n in names2]
    louts, routs, rvout, _ = K.join_fused(k1.contiguous(), v1, k2.contiguous(), v2, left,
                                          [t2.columns[i] for i in idx2], [t2.valid[i] for i in idx2], outer, parts, po2, po1)
    ncol1 = len(t1.columns)
    lvalid = [louts[lmask_at[i]] if i in lmask_at else None for i in range(ncol1)]
    dicts = dict(t1.dictionaries)
    dicts.update({n: t2.dictionaries[n] for n in names2 if n in t2.dictionaries})
    return B200DataFrame(B200Table(out_schema, louts[:ncol1] + routs, lvalid + rvout, dicts))


def _matched_mask(n: int, ri: torch.Tensor) -> torch.Tensor:
    m = torch.zeros(n, dtype=torch.uint8, device=ri.device)
    m[ri[ri >= 0]] = 1
    return m


def _drop_collisions(t1, t2, keys, li, ri, outer_side):
    both = (li >= 0) & (ri >= 0)
    ok = torch.ones_like(both)
    if bool(both.any()):
        ok[both] = _verify(t1, t2, keys, li[both], ri[both])
    if outer_side is None:
        return li[ok], ri[ok]
    # a false candidate of an outer join degrades to "unmatched" unless the row has a true match
    probe = li if outer_side == "left" else ri
    n = (t1 if outer_side == "left" else t2).num_rows
    has = torch.zeros(n, dtype=torch.bool, device=li.device)
    has[probe[ok & both]] = True
    keep = ok.clone()
    bad = ~ok
    first_bad = torch.zeros_like(bad)
    if bool(bad.any()):
        # keep one NULL-extended row for probe rows that lost all their candidates
        idx = K.compact_indices((bad).contiguous())
        rows = probe[idx]
        lost = ~has[rows]
        uniq, inv = torch.unique(rows[lost], return_inverse=True)
        firsts = torch.full((uniq.numel(),), li.numel(), dtype=torch.int64, device=li.device)
        firsts.scatter_reduce_(0, inv, idx[lost], reduce="amin")
        first_bad[firsts] = True
    keep |= first_bad
    li, ri = li.clone(), ri.clone()
    if outer_side == "left":
        ri[first_bad] = -1
    else:
        li[first_bad] = -1
    return li[keep], ri[keep]


def _assemble(t1: B200Table, t2: B200Table, keys: List[str], out_schema: Schema, li: torch.Tensor,
              ri: torch.Tensor, how: str) -> B200DataFrame:
    """Gather the output columns: df1's columns, then df2's non-key columns; key columns of
    NULL-extended left rows (right/full outer) come from the right side."""
    l_null = how in ("right_outer", "full_outer")
    r_null = how in ("left_outer", "full_outer")
    lcols, lvalid = K.gather_rows(t1.columns, t1.valid, li, want_valid=l_null)
    names2 = [n for n in t2.schema.names if n not in keys]
    idx2 = [t2.schema.index_of_key(n) for n in names2]
    rcols, rvalid = K.gather_rows([t2.columns[i] for i in idx2], [t2.valid[i] for i in idx2], ri,
                                  want_valid=r_null)
    dicts = dict(t1.dictionaries)
    dicts.update({n: t2.dictionaries[n] for n in names2 if n in t2.dictionaries})
    if l_null and len(keys) > 0:
        kidx2 = [t2.schema.index_of_key(k) for k in keys]
        kc, kv = K.gather_rows([t2.columns[i] for i in kidx2], [t2.valid[i] for i in kidx2], ri, want_valid=True)
        from_right = li < 0
        for k, c2, v2 in zip(keys, kc, kv):
            i1 = t1.schema.index_of_key(k)
            if k in t1.dictionaries:  # translate right codes into the left dictionary, extending it
                d1, d2 = t1.dictionaries[k], t2.dictionaries[k]
                merged = pa.concat_arrays([d1, d2.filter(pc.invert(pc.is_in(d2, value_set=d1)))])
                pos = pc.index_in(d2, value_set=merged)
                m = torch.from_numpy(pos.to_numpy(zero_copy_only=False).astype("int32")).to(li.device)
                c2 = m[c2.long().clamp(min=0)]
                dicts[k] = merged
            lcols[i1] = torch.where(from_right, c2.to(lcols[i1].dtype), lcols[i1])
            lvalid[i1] = torch.where(from_right, v2, lvalid[i1])
    return B200DataFrame(B200Table(out_schema, lcols + rcols, lvalid + rvalid, dicts))
