"""``ColumnMap``: a per-row map of ``fa.transform`` written as column expressions (K4).

The reference applies the map function to every partition after grouping
(``PandasMapEngine.map_dataframe``, fugue/execution/native_execution_engine.py:156-164).  A map that is
a list of column expressions needs no second pass on the GPU: the B200 map engine evaluates affine
expressions (``x * a + y * b + c`` over at most two columns of one class) inside the scatter kernel
of the hash partition, between the gather from the staged tile and the store
(``fb_partition_apply_map``).  Everything else is still evaluated on the device, by the expression
evaluator (K8) over the partitioned table - calling the object does exactly that.

    fa.transform(df, ColumnMap("key", "v0", (col("v0") * 2 + col("v1")).alias("w")),
                 schema="key:long,v0:double,w:double", partition=PartitionSpec(by="key", algo="hash", num=256))
"""
import struct
from typing import Any, List, Optional, Tuple

import pyarrow as pa

from . import kernels as K
from .column import ColumnExpr, Kind, SelectColumns, col as _col, is_agg
from .table import B200Table


def _f64_bits(v: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(v)))[0]


class ColumnMap:
    def __init__(self, *columns: Any):
        assert len(columns) > 0, "ColumnMap needs at least one output column"
        self.columns: List[ColumnExpr] = [_col(c) for c in columns]
        for c in self.columns:
            if is_agg(c):
                raise ValueError(f"{c} is an aggregation: a map is row-wise")

    def select(self, t: B200Table) -> SelectColumns:
        return SelectColumns(*self.columns).replace_wildcard(t.schema).assert_all_with_names()

    def __call__(self, t: B200Table) -> B200Table:
        """Unfused evaluation on the device (one pass of the expression evaluator over ``t``)."""
        from . import expr as X

        out = X.project(t, self.select(t).all_cols)
        keep = t.partition_keys is not None and all(k in out.schema for k in t.partition_keys)
        return B200Table(out.schema, out.columns, out.valid, out.dictionaries, t.offsets if keep else None,
                         t.partition_keys if keep else None)

    # ---- fusion plan -----------------------------------------------------------------------------
    def fusion_units(self, t: B200Table) -> Optional[List[Tuple[Any, Any, int, int, int, int, pa.DataType]]]:
        """One ``(x, y, mode, a, b, c, type)`` per output column if EVERY column can be produced by the
        scatter kernel's epilogue (plain 8-byte NULL-free columns and affine expressions of them), else None."""
        units = []
        for e in self.select(t).all_cols:
            u = _unit_of(e, t)
            if u is None:
                return None
            units.append(u)
        return units


def _plain_column(e: ColumnExpr, t: B200Table) -> Optional[int]:
    if e.kind != Kind.NAMED or e.as_type is not None or e.name not in t.schema:
        return None
    i = t.schema.index_of_key(e.name)
    if t.columns[i].element_size() != 8 or t.valid[i] is not None or e.name in t.dictionaries:
        return None
    if t.columns[i].data_ptr() % 16 != 0:
        return None
    return i


def _term(e: ColumnExpr, t: B200Table) -> Optional[Tuple[Any, int]]:
    """``col`` | ``-col`` | ``col * lit`` | ``lit * col`` -> (coefficient, column index)."""
    if e.as_type is not None:
        return None
    if e.kind == Kind.NAMED:
        i = _plain_column(e, t)
        return None if i is None else (1, i)
    if e.kind == Kind.UNARY and e.op == "-":
        i = _plain_column(e.arg, t)
        return None if i is None else (-1, i)
    if e.kind == Kind.BINARY and e.op == "*":
        for c, l in ((e.left, e.right), (e.right, e.left)):
            if l.kind == Kind.LITERAL and l.as_type is None and type(l.value) in (int, float):
                i = _plain_column(c, t)
                if i is not None:
                    return (l.value, i)
    return None


def _unit_of(e: ColumnExpr, t: B200Table) -> Optional[Tuple[Any, Any, int, int, int, int, pa.DataType]]:
    i = _plain_column(e, t)
    if i is not None:
        return (t.columns[i], None, K.MAP_COPY, 0, 0, 0, t.schema.types[i])
    if e.as_type is not None:
        return None
    # peel: ((term [+-] term) [+-] lit) in exactly this association, so that the rounding order of the
    # fused form (a*x + b*y) + c is the evaluator's
    const: Any = None
    body = e
    if e.kind == Kind.BINARY and e.op in ("+", "-") and e.right.kind == Kind.LITERAL and e.right.as_type is None \
            and type(e.right.value) in (int, float) and e.left.as_type is None:
        const = e.right.value if e.op == "+" else -e.right.value
        body = e.left
    terms = []
    first = _term(body, t)
    if first is not None:
        terms = [first]
    elif body.kind == Kind.BINARY and body.op in ("+", "-") and body.as_type is None:
        a, b = _term(body.left, t), _term(body.right, t)
        if a is None or b is None:
            return None
        terms = [a, (b[0] if body.op == "+" else -b[0], b[1])]
    else:
        return None
    if const is None and len(terms) == 1 and terms[0][0] == 1:
        return None  # plain column with an alias only: handled as copy by the caller's NAMED case
    cols = [t.columns[i] for _, i in terms]
    tps = [t.schema.types[i] for _, i in terms]
    if all(tp == pa.float64() for tp in tps):
        coef = [_f64_bits(c) for c, _ in terms] + [0]
        cbits = _f64_bits(-0.0 if const is None else const)  # x + (-0.0) == x for every x, signed zeros included
        return (cols[0], cols[1] if len(cols) > 1 else None, K.MAP_AFFINE_F64, coef[0], coef[1], cbits, pa.float64())
    if all(tp == pa.int64() for tp in tps) and all(type(c) is int for c, _ in terms) and type(const or 0) is int:
        coef = [c & ((1 << 64) - 1) for c, _ in terms] + [0]
        return (cols[0], cols[1] if len(cols) > 1 else None, K.MAP_AFFINE_I64, coef[0], coef[1],
                (const or 0) & ((1 << 64) - 1), pa.int64())
    return None
