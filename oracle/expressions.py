"""Oracle: pandas restatement of ``ExecutionEngine.select / filter / assign`` on column expressions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) - never imported by the product path.

The reference turns an expression tree into SQL text (fugue/column/sql.py:275-347,
``SQLExpressionGenerator.select``) and runs it with qpd on pandas (third party, qpd>=0.4.4,
fugue/execution/native_execution_engine.py:59-66).  qpd is absent here, so this module restates what
that SQL means, operator by operator, on pandas *nullable* arrays (``Int64 / Float64 / boolean``),
which implement exactly the SQL rules the conformance tests pin
(fugue_test/execution_suite.py:85-206):

* arithmetic and comparisons propagate NULL; ``/`` is true division;
* ``AND`` / ``OR`` are Kleene three-valued (pandas ``boolean`` ``&`` / ``|``);
* ``WHERE`` / ``HAVING`` keep rows whose predicate is TRUE (NULL drops the row);
* ``GROUP BY`` keys are every non-aggregate, non-literal select column
  (``SelectColumns.group_keys``, fugue/column/sql.py:38-96), NULL keys form a group;
* aggregates skip NULLs; SUM/MIN/MAX/AVG of an all-NULL group is NULL; COUNT(*) counts rows;
* NaN in a float column of a pandas frame is NULL (Fugue's pandas convention).

The trees are the DSL objects of ``fugue_b200.column`` (API objects, not compute code).
Parity pinning: ``tests/test_oracle_expressions.py`` checks this module against the literal input/output
tables of the reference's test_filter / test_select / test_assign / test_aggregate.
"""
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
import pyarrow as pa

from fugue_b200.column import ColumnExpr, Kind, SelectColumns, col, is_agg


def _nullable(s: pd.Series) -> pd.Series:
    """Column of a (numpy-typed or object) pandas frame -> nullable dtype."""
    if pd.api.types.is_bool_dtype(s.dtype):
        return s.astype("boolean")
    if pd.api.types.is_float_dtype(s.dtype):
        return pd.Series(pd.array(s.to_numpy(dtype="float64", na_value=np.nan), dtype="Float64"), index=s.index)
    if pd.api.types.is_integer_dtype(s.dtype):
        return s.astype("Int64")
    if s.dtype == object or pd.api.types.is_string_dtype(s.dtype):
        return s.astype("string")
    return s


def _pandas_dtype(tp: pa.DataType) -> str:
    if pa.types.is_boolean(tp):
        return "boolean"
    if pa.types.is_floating(tp):
        return "Float64" if tp == pa.float64() else "Float32"
    if pa.types.is_integer(tp):
        return {8: "Int8", 16: "Int16", 32: "Int32", 64: "Int64"}[tp.bit_width]
    if pa.types.is_string(tp):
        return "string"
    raise NotImplementedError(str(tp))


def _cast(v: Any, tp: pa.DataType, n: int, index: Any) -> pd.Series:
    s = v if isinstance(v, pd.Series) else pd.Series([v] * n, index=index)
    target = _pandas_dtype(tp)
    if target == "string":
        if pd.api.types.is_bool_dtype(s.dtype):
            return s.map(lambda x: pd.NA if x is pd.NA or x is None else ("true" if x else "false")).astype("string")
        return s.map(lambda x: pd.NA if x is pd.NA or x is None or (isinstance(x, float) and np.isnan(x))
                     else str(x)).astype("string")
    if target.startswith("Int") and pd.api.types.is_float_dtype(s.dtype):
        s = pd.Series(pd.array(np.trunc(s.to_numpy(dtype="float64", na_value=np.nan)), dtype="Float64"),
                      index=s.index)  # SQL CAST truncates toward zero
    if target == "boolean" and not pd.api.types.is_bool_dtype(s.dtype):
        return (s != 0).astype("boolean")
    if s.dtype == object:
        s = s.astype(target)
    return s.astype(target)


def evaluate(e: Any, df: pd.DataFrame, aggs: Optional[Dict[str, pd.Series]] = None) -> Any:
    """Value of expression ``e`` on every row of ``df`` (Series) or a python scalar for literals."""
    if not isinstance(e, ColumnExpr):
        return e
    n = len(df)
    if e.kind == Kind.AGG:
        assert aggs is not None, f"aggregation {e} outside an aggregating select"
        res: Any = aggs[e.alias("").cast(None).fingerprint()]
    elif e.kind == Kind.NAMED:
        res = _nullable(df[e.name])
    elif e.kind == Kind.LITERAL:
        res = pd.NA if e.value is None else e.value
    elif e.kind == Kind.WILDCARD:
        raise ValueError("'*' has no value")
    elif e.kind == Kind.UNARY:
        v = evaluate(e.col, df, aggs)
        if e.op == "IS_NULL":
            res = v.isna().astype("boolean") if isinstance(v, pd.Series) else (v is pd.NA)
        elif e.op == "NOT_NULL":
            res = v.notna().astype("boolean") if isinstance(v, pd.Series) else (v is not pd.NA)
        elif e.op == "-":
            res = -v
        elif e.op == "~":
            res = ~v.astype("boolean") if isinstance(v, pd.Series) else (pd.NA if v is pd.NA else not bool(v))
        else:
            raise NotImplementedError(e.op)
    elif e.kind == Kind.BINARY:
        a, b = evaluate(e.left, df, aggs), evaluate(e.right, df, aggs)
        if not isinstance(a, pd.Series) and not isinstance(b, pd.Series):
            a = pd.Series(pd.array([a] * n), index=df.index)
            a = _nullable(a) if a.dtype != object else a.astype("Float64")
        op = e.op
        if op in ("&", "|"):
            a = a.astype("boolean") if isinstance(a, pd.Series) else a
            b = b.astype("boolean") if isinstance(b, pd.Series) else b
            res = (a & b) if op == "&" else (a | b)
        elif op == "+":
            res = a + b
        elif op == "-":
            res = a - b
        elif op == "*":
            res = a * b
        elif op == "/":
            res = (a.astype("Float64") if isinstance(a, pd.Series) else float(a)) / \
                (b.astype("Float64") if isinstance(b, pd.Series) else (pd.NA if b is pd.NA else float(b)))
        elif op == "<":
            res = a < b
        elif op == "<=":
            res = a <= b
        elif op == ">":
            res = a > b
        elif op == ">=":
            res = a >= b
        elif op == "==":
            res = a == b
        elif op == "!=":
            res = a != b
        else:
            raise NotImplementedError(op)
    elif e.kind == Kind.CALL and e.func.upper() == "COALESCE":
        vals = [evaluate(a, df, aggs) for a in e.args]
        series = [v for v in vals if isinstance(v, pd.Series)]
        is_f = any(pd.api.types.is_float_dtype(s.dtype) for s in series) or \
            any(isinstance(v, float) for v in vals if not isinstance(v, pd.Series))
        out = pd.Series(pd.array([pd.NA] * n, dtype="Float64" if is_f else "Int64"), index=df.index)
        if series and all(pd.api.types.is_bool_dtype(s.dtype) for s in series) and not is_f:
            out = out.astype("boolean")
        for v in vals:
            if isinstance(v, pd.Series):
                out = out.where(out.notna(), v.astype(out.dtype))
            elif v is not pd.NA:
                out = out.fillna(v)
        res = out
    else:
        raise NotImplementedError(str(e))
    if e.as_type is not None:
        res = _cast(res, e.as_type, n, df.index)
    return res


def _as_column(v: Any, n: int, index: Any, e: ColumnExpr) -> pd.Series:
    if isinstance(v, pd.Series):
        return v.reset_index(drop=True)
    if v is pd.NA:
        raise NotImplementedError(f"NULL literal column {e} needs a cast")
    dtype = "boolean" if isinstance(v, bool) else "Int64" if isinstance(v, int) else \
        "Float64" if isinstance(v, float) else "string"
    return pd.Series(pd.array([v] * n, dtype=dtype))


def _predicate(e: ColumnExpr, df: pd.DataFrame, aggs: Optional[Dict[str, pd.Series]] = None) -> np.ndarray:
    v = evaluate(e.alias("") if e.as_name else e, df, aggs)
    if not isinstance(v, pd.Series):
        return np.full(len(df), bool(v) if v is not pd.NA else False)
    return v.astype("boolean").fillna(False).to_numpy(dtype=bool)


def _find_aggs(e: Any, out: List[ColumnExpr]) -> None:
    if not isinstance(e, ColumnExpr):
        return
    if e.kind == Kind.AGG:
        out.append(e)
    elif e.has_args:
        for a in list(e.args) + list(e.kwargs.values()):
            _find_aggs(a, out)


def _agg_series(func: str, values: Optional[pd.Series], codes: np.ndarray, ngroups: int, n: int) -> pd.Series:
    g = pd.Series(codes)
    if func == "COUNT":
        if values is None:
            return pd.Series(np.bincount(codes, minlength=ngroups), dtype="Int64")
        return values.notna().astype("int64").groupby(codes).sum().reindex(range(ngroups), fill_value=0).astype("Int64")
    assert values is not None
    grouped = values.groupby(g.to_numpy())
    if func == "SUM":
        res = grouped.sum(min_count=1)
    elif func == "MIN":
        res = grouped.min()
    elif func == "MAX":
        res = grouped.max()
    elif func == "AVG":
        res = grouped.mean().astype("Float64")
    else:
        raise NotImplementedError(func)
    return res.reindex(range(ngroups))


def select(df: pd.DataFrame, cols: SelectColumns, where: Optional[ColumnExpr] = None,
           having: Optional[ColumnExpr] = None) -> pd.DataFrame:
    """``SELECT cols FROM df WHERE where GROUP BY <inferred> HAVING having`` -> nullable-typed frame."""
    names = list(df.columns)
    out_cols: List[ColumnExpr] = []
    for c in cols.all_cols:
        if c.kind == Kind.WILDCARD:
            out_cols.extend(col(n) for n in names)
        else:
            out_cols.append(c)
    sel = SelectColumns(*out_cols, arg_distinct=cols.is_distinct).assert_all_with_names()
    df = df.reset_index(drop=True)
    if where is not None:
        assert not is_agg(where)
        df = df[_predicate(where, df)].reset_index(drop=True)
    n = len(df)
    if not sel.has_agg:
        res = pd.DataFrame({c.output_name: _as_column(evaluate(c, df), n, df.index, c) for c in sel.all_cols})
        return res.drop_duplicates().reset_index(drop=True) if sel.is_distinct else res
    # ---- aggregation
    key_vals = [_as_column(evaluate(k, df), n, df.index, k) for k in sel.group_keys]
    if key_vals:
        kf = pd.DataFrame({f"k{i}": v for i, v in enumerate(key_vals)})
        codes = kf.groupby(list(kf.columns), dropna=False, sort=True).ngroup().to_numpy()
        ngroups = int(codes.max()) + 1 if n > 0 else 0
        first = pd.Series(np.arange(n)).groupby(codes).first().to_numpy() if n > 0 else np.array([], dtype=int)
        gkeys = kf.iloc[first].reset_index(drop=True)
    else:
        codes = np.zeros(n, dtype=np.int64)
        ngroups = 1  # a global aggregate always yields one row
        gkeys = pd.DataFrame(index=range(1))
    key_uuid = {k.fingerprint(): gkeys[f"k{i}"] for i, k in enumerate(sel.group_keys)}
    found: List[ColumnExpr] = []
    for c in sel.all_cols:
        _find_aggs(c, found)
    if having is not None:
        _find_aggs(having, found)
    aggs: Dict[str, pd.Series] = {}
    for a in found:
        uid = a.alias("").cast(None).fingerprint()
        if uid in aggs:
            continue
        if a.arg.kind == Kind.WILDCARD:
            vals = None
        else:
            vals = _as_column(evaluate(a.arg, df), n, df.index, a.arg)
        if a.is_distinct:  # COUNT(DISTINCT x): NULLs are not counted; COUNT(DISTINCT *): distinct rows
            assert a.func == "COUNT", str(a)
            if vals is None:
                sub = df.assign(__fb_code=codes).drop_duplicates()
                cnt = sub.groupby("__fb_code").size()
            else:
                cnt = vals.groupby(codes).nunique(dropna=True)
            aggs[uid] = cnt.reindex(range(ngroups), fill_value=0).astype("Int64").reset_index(drop=True)
            continue
        aggs[uid] = _agg_series(a.func, vals, codes, ngroups, n).reset_index(drop=True)
    gframe = pd.DataFrame(index=range(ngroups))

    def group_value(e: ColumnExpr) -> Any:
        """Evaluate a select column on the group table: group keys come from ``gkeys``."""
        uid = e.alias("").cast(None).fingerprint()
        if uid in key_uuid and not is_agg(e):
            v: Any = key_uuid[uid]
            return _cast(v, e.as_type, ngroups, gframe.index) if e.as_type is not None else v
        return _eval_on_groups(e)

    def _eval_on_groups(e: Any) -> Any:
        if not isinstance(e, ColumnExpr):
            return e
        uid = e.alias("").cast(None).fingerprint()
        if uid in key_uuid and not is_agg(e):
            v = key_uuid[uid]
        elif e.kind == Kind.AGG:
            v = aggs[uid]
        elif e.kind == Kind.LITERAL:
            v = pd.NA if e.value is None else e.value
        elif e.kind == Kind.UNARY:
            sub = _eval_on_groups(e.col)
            tmp = pd.DataFrame({"x": _as_column(sub, ngroups, gframe.index, e.col)})
            v = evaluate(ColumnExpr(Kind.UNARY, e.op, [col("x")]), tmp)
        elif e.kind == Kind.BINARY:
            a, b = _eval_on_groups(e.left), _eval_on_groups(e.right)
            tmp = pd.DataFrame({"a": _as_column(a, ngroups, gframe.index, e.left),
                                "b": _as_column(b, ngroups, gframe.index, e.right)})
            v = evaluate(ColumnExpr(Kind.BINARY, e.op, [col("a"), col("b")]), tmp)
        else:
            raise NotImplementedError(str(e))
        if e.as_type is not None:
            v = _cast(v, e.as_type, ngroups, gframe.index)
        return v

    if having is not None:
        hv = _eval_on_groups(having.alias("") if having.as_name else having)
        keep = hv.astype("boolean").fillna(False).to_numpy(dtype=bool) if isinstance(hv, pd.Series) \
            else np.full(ngroups, bool(hv))
    else:
        keep = np.ones(ngroups, dtype=bool)
    res = pd.DataFrame({c.output_name: _as_column(group_value(c), ngroups, gframe.index, c) for c in sel.all_cols})
    res = res[keep].reset_index(drop=True)
    return res.drop_duplicates().reset_index(drop=True) if sel.is_distinct else res


def filter_rows(df: pd.DataFrame, condition: ColumnExpr) -> pd.DataFrame:
    """``ExecutionEngine.filter`` (execution_engine.py:808-834) = SELECT * WHERE condition."""
    return df[_predicate(condition, df.reset_index(drop=True))].reset_index(drop=True)


def assign(df: pd.DataFrame, columns: List[ColumnExpr]) -> pd.DataFrame:
    """``ExecutionEngine.assign`` (execution_engine.py:836-887)."""
    SelectColumns(*columns).assert_no_wildcard().assert_all_with_names().assert_no_agg()
    pos = {n: i for i, n in enumerate(df.columns)}
    cols: List[ColumnExpr] = [col(n) for n in pos]
    for c in columns:
        c = c.infer_alias()
        if c.output_name in pos:
            cols[pos[c.output_name]] = c
        else:
            cols.append(c)
    return select(df, SelectColumns(*cols))
