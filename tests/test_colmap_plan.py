"""K4 fusion planning (``ColumnMap.fusion_units``, fugue_b200/colmap.py) on CPU: for random affine maps the plan
(x, y, mode, a, b, c) - evaluated by a numpy restatement of the scatter kernel's epilogue ``ws_apply_map``
(csrc/fb_partition.cu: ``(a*x [+ b*y]) + c`` with separately rounded multiply / add, wrapping int64) - must give
BIT-identical results to the same expression compiled for the K8 evaluator and run on its machine model
(tests/_expr_sim.py).  That is the claim of DESIGN.md 6 ("fused map == partition, then evaluator"); the GPU tests
check the kernel itself against the same model."""
import struct

import numpy as np
import pyarrow as pa
import pytest
import torch

from fugue_b200 import kernels as K
from fugue_b200.colmap import ColumnMap
from fugue_b200.column import col, lit
from fugue_b200.schema import Schema
from fugue_b200.table import B200Table
from test_expr_compiler import _run


def _table(n: int, seed: int) -> B200Table:
    rng = np.random.default_rng(seed)
    special = np.array([0.0, -0.0, np.inf, -np.inf, 1e308, -1e308, 5e-324, 1.5, -2.25], dtype=np.float64)
    f = [np.where(rng.random(n) < 0.1, rng.choice(special, n), rng.standard_normal(n) * 10.0 ** rng.integers(-3, 6, n))
         for _ in range(3)]
    i = [rng.integers(-2**62, 2**62, n, dtype=np.int64), rng.integers(-1000, 1000, n, dtype=np.int64)]
    v = (rng.random(n) < 0.8).astype(np.uint8)
    cols = [torch.from_numpy(np.ascontiguousarray(a)) for a in f + i] + [torch.from_numpy(f[0].copy())]
    sch = Schema("x:double,y:double,z:double,k:long,m:long,nul:double")
    return B200Table(sch, cols, [None] * 5 + [torch.from_numpy(v)])


def _epilogue(unit, n: int) -> np.ndarray:
    """numpy model of ``ws_apply_map`` over whole columns; returns the output bit patterns (uint64)."""
    x, y, mode, a, b, c, _ = unit
    xs = x.numpy()
    if mode == K.MAP_COPY:
        return xs.view(np.uint64).copy()
    as_f = lambda bits: struct.unpack("<d", struct.pack("<Q", bits))[0]  # noqa: E731
    with np.errstate(all="ignore"):
        if mode == K.MAP_AFFINE_F64:
            r = np.float64(as_f(a)) * xs
            if y is not None:
                r = r + np.float64(as_f(b)) * y.numpy()
            return (r + np.float64(as_f(c))).view(np.uint64)
        assert mode == K.MAP_AFFINE_I64
        r = np.uint64(a) * xs.view(np.uint64)
        if y is not None:
            r = r + np.uint64(b) * y.numpy().view(np.uint64)
        return r + np.uint64(c)


def _random_affine(rng, floats: bool):
    names = ["x", "y", "z"] if floats else ["k", "m"]
    coef = (lambda: float(np.round(rng.normal() * 4, 3))) if floats else (lambda: int(rng.integers(-9, 10)))

    def term():
        c = col(names[rng.integers(len(names))])
        r = rng.random()
        if r < 0.3:
            return c
        if r < 0.4:
            return -c
        return c * coef() if r < 0.7 else lit(coef()) * c

    e = term()
    if rng.random() < 0.6:
        t2 = term()
        e = e + t2 if rng.random() < 0.5 else e - t2
    if rng.random() < 0.6:
        e = e + coef() if rng.random() < 0.5 else e - coef()
    return e


@pytest.mark.parametrize("floats", [True, False], ids=["f64", "i64"])
def test_fused_plan_equals_evaluator_bit_for_bit(floats):
    rng = np.random.default_rng(17 if floats else 18)
    t = _table(2000, 5)
    fused = 0
    for _ in range(150):
        e = _random_affine(rng, floats).alias("w")
        units = ColumnMap("k", e).fusion_units(t)
        if units is None:          # e.g. "x" alone with an alias: not an affine unit, not an error
            continue
        assert units[0][2] == K.MAP_COPY and units[0][0] is t.columns[3]
        got = _epilogue(units[1], t.num_rows)
        want, _ = _run(t, [e])
        w = want[0].to_numpy(dtype=np.float64 if floats else np.int64, na_value=np.nan if floats else 0)
        wb = w.view(np.uint64)
        if floats:   # a NaN (inf - inf) has no defined sign / payload: NaN == NaN, everything else bit for bit
            both_nan = np.isnan(got.view(np.float64)) & np.isnan(w)
            assert np.array_equal(got[~both_nan], wb[~both_nan]), str(e)
        else:
            assert np.array_equal(got, wb), str(e)
        fused += 1
    assert fused > 100


def test_what_does_not_fuse():
    t = _table(64, 1)
    for e in [col("x") / col("y"), (col("x") + 1) * col("y"), col("x") * col("y"), col("nul") * 2, col("x").cast(int),
              col("x") + col("k"), col("k") * 2.5, col("x") * 2 + col("y") * 3 + col("z"), col("x") > 1,
              (col("x") + col("y")).cast("long")]:
        assert ColumnMap("k", e.alias("w")).fusion_units(t) is None, str(e)
    units = ColumnMap("k", "x", col("m")).fusion_units(t)
    assert [u[2] for u in units] == [K.MAP_COPY] * 3 and [u[6] for u in units] == [pa.int64(), pa.float64(), pa.int64()]
    with pytest.raises(ValueError):
        from fugue_b200.column import functions as ff

        ColumnMap(ff.sum(col("x")))


def test_default_constant_keeps_signed_zeros():
    """``x * 2`` has no constant: the epilogue adds -0.0, the only value with ``v + c == v`` for every v."""
    t = _table(64, 2)
    (u,) = ColumnMap((col("x") * 2.0).alias("w")).fusion_units(t)
    assert u[5] == struct.unpack("<Q", struct.pack("<d", -0.0))[0]
    xs = t.columns[0].numpy()
    with np.errstate(all="ignore"):
        assert np.array_equal(_epilogue(u, 64), (xs * 2.0).view(np.uint64))
