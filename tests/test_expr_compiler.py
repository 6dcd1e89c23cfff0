"""Host-side expression compiler (fugue_b200/expr.py) checked WITHOUT a GPU: programs are generated for
a table of CPU tensors, executed by the numpy model of the accumulator machine (tests/_expr_sim.py)
and compared with oracle/expressions.py.  The -m gpu tests run the same programs on the device."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pytest
import torch

from fugue_b200 import expr as X
from fugue_b200 import kernels as K
from fugue_b200.column import SelectColumns, col, functions as ff, lit, null
from fugue_b200.schema import Schema
from fugue_b200.table import B200Table
from oracle import expressions as OX
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _expr_sim as sim  # noqa: E402


def _table(pdf: pd.DataFrame) -> B200Table:
    fields, cols, valids = [], [], []
    for name in pdf.columns:
        s = pdf[name]
        na = s.isna().to_numpy()
        if pd.api.types.is_bool_dtype(s.dtype):
            tp, arr = pa.bool_(), s.fillna(False).to_numpy(dtype=np.uint8)
        elif pd.api.types.is_float_dtype(s.dtype):
            tp, arr = pa.float64(), np.nan_to_num(s.to_numpy(dtype=np.float64, na_value=0.0))
        else:
            bits = 32 if str(s.dtype).lower() == "int32" else 64
            tp = pa.int32() if bits == 32 else pa.int64()
            arr = s.fillna(0).to_numpy(dtype=np.int32 if bits == 32 else np.int64)
        fields.append(pa.field(name, tp))
        cols.append(torch.from_numpy(np.ascontiguousarray(arr)))
        valids.append(torch.from_numpy((~na).astype(np.uint8)) if na.any() else None)
    return B200Table(Schema(fields), cols, valids)


def _run(t: B200Table, exprs):
    """Compile every expression into ONE program and simulate it; returns pandas nullable columns."""
    prog = X._Program(t)
    meta = []
    for e in exprs:
        cls, nullable = prog.compile(e, top=True)
        dtype = {"i": torch.int64, "f": torch.float64, "b": torch.uint8}[cls]
        prog.output(dtype, True)
        meta.append(cls)
    cols = [t.columns[i].numpy() for i in prog.cols]
    valid = [None if t.valid[i] is None else t.valid[i].numpy() for i in prog.cols]
    outs, outv = sim.run(t.num_rows, cols, valid, prog.ins, [K.expr_type_of(o[0]) for o in prog.outs])
    res = []
    for cls, o, v in zip(meta, outs, outv):
        dt = {"i": "Int64", "f": "Float64", "b": "boolean"}[cls]
        arr = pd.array(o.astype(bool) if cls == "b" else o, dtype=dt)
        arr[v == 0] = pd.NA
        res.append(pd.Series(arr))
    return res, prog


def _random(n=4000, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n)
    x[rng.random(n) < 0.2] = np.nan
    return pd.DataFrame({
        "a": rng.integers(-50, 50, n).astype(np.int64), "b": rng.integers(0, 7, n).astype(np.int32), "x": x,
        "y": rng.standard_normal(n) * 10,
        "g": pd.array(np.where(rng.random(n) < 0.1, None, rng.integers(0, 20, n)), dtype="Int64"),
        "p": pd.array(np.where(rng.random(n) < 0.15, None, rng.random(n) < 0.5), dtype="boolean")})


def _same(got: pd.Series, want: pd.Series, name: str):
    gn, wn = got.isna().to_numpy(), want.isna().to_numpy()
    assert (gn == wn).all(), name
    g = got[~gn].to_numpy(dtype=np.float64)
    w = want[~wn].to_numpy(dtype=np.float64)
    assert np.array_equal(g, w, equal_nan=True), name


EXPRS = [
    col("a") * col("b") - 3, col("x") / col("y"), col("a") / col("b"), -col("x"), -col("b"), 10 - col("a"),
    2.5 / col("y"), (col("a") + col("b")) * (col("x") - col("y")), (col("a") - 1) / (col("b") + 1),
    (col("x") > 0) & col("p"), (col("x") > 0) | col("p"), ~col("p"), col("x").is_null() | (col("a") >= col("b")),
    col("g") == col("b"), col("g") != 3, 3 < col("g"), (col("a") > 0) & ((col("b") < 3) | (col("x") > col("y"))),
    col("p") & null(), col("p") | null(), null() & col("p"), lit(True) & col("p"), col("a") + null(),
    ff.coalesce(col("x"), col("y")), ff.coalesce(col("g"), -1), ff.coalesce(col("g"), col("x"), 0.5),
    ff.coalesce(col("x") * 2, col("g") + 1, col("a")), ff.coalesce(null(), col("g")),
    (col("a") + 1.5).cast(int), col("x").cast("long"), col("b").cast(float), (col("a") > 0).cast(int),
    col("g").cast(bool), (col("x") * 3).cast(bool) & col("p"), (col("a") & col("b")), ~(col("a") - 1),
    (col("x") + col("y") > 0) | (col("g").is_null() & col("p")), (col("x").not_null() & (col("x") < 0.5)),
]


def test_compiled_programs_match_oracle():
    pdf = _random()
    t = _table(pdf)
    named = [e.alias(f"c{i}") for i, e in enumerate(EXPRS)]
    want = OX.select(pdf, SelectColumns(*named))
    for lo in range(0, len(named), 8):  # several expressions share one program
        got, prog = _run(t, named[lo:lo + 8])
        assert len(prog.ins) <= K.EXPR_MAX_INS
        for e, s in zip(named[lo:lo + 8], got):
            _same(s, want[e.output_name], str(e))


def test_leaf_operands_need_no_temporaries():
    t = _table(_random(64))
    prog = X._Program(t)
    prog.compile(col("x") * col("y") + col("a"))
    assert [i[0] for i in prog.ins] == [K.X_MOV, K.X_MUL_F, K.X_ADD_F]
    assert prog.ins[2][3] == K.XF_B_I2F and all(i[1] == K.XK_COL for i in prog.ins)
    prog = X._Program(t)
    prog.compile((col("x") > 0) & (col("y") < 0.5))
    assert [i[0] for i in prog.ins] == [K.X_MOV, K.X_GT_F, K.X_ST, K.X_MOV, K.X_LT_F, K.X_AND]
    assert len(prog.free) == K.EXPR_NREGS
    prog = X._Program(t)
    prog.compile(10 - col("a") * 2)
    assert [i[0] for i in prog.ins] == [K.X_MOV, K.X_MUL_I, K.X_RSUB_I]
    prog = X._Program(t)
    prog.compile(col("a") / col("b"))       # integer leaves are converted while they are loaded
    assert [(i[0], i[3]) for i in prog.ins] == [(K.X_MOV, K.XF_B_I2F), (K.X_DIV_F, K.XF_B_I2F)]
    prog = X._Program(t)
    prog.compile((col("a") + 1) / 2)        # ... but not after integer arithmetic
    assert [i[0] for i in prog.ins] == [K.X_MOV, K.X_ADD_I, K.X_I2F, K.X_DIV_F]


def test_resource_limits_and_errors():
    t = _table(_random(64))
    deep = col("a")
    for i in range(6):  # right-nested compound operands: one temporary per level
        deep = (col("a") + i) * (deep - col("b") * 2)
    with pytest.raises(X._OutOfResources):
        X._Program(t).compile(deep)
    with pytest.raises(KeyError):
        X._Program(t).compile(col("nope") + 1)
    with pytest.raises(ValueError):
        X._Program(t).compile(ff.max(col("a")) + 1)
    with pytest.raises(NotImplementedError):
        X._Program(t).compile(col("a") + "s")
