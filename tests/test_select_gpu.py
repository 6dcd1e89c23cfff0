"""select / filter / assign / expression aggregates on the B200 engine.

* literal expectations of fugue_test/execution_suite.py:85-206 (test_filter, test_select, test_assign,
  test_aggregate) through the fa.* API;
* the device evaluator (fb_eval_expr through the C ABI) against numpy on random programs' building
  blocks, and whole selects against oracle/expressions.py on random nullable tables."""
import numpy as np
import pandas as pd
import pytest
from pytest import raises

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from fugue_b200 import api as fa
from fugue_b200 import kernels as K
from fugue_b200.column import SelectColumns, all_cols, col, functions as ff, lit, null
from fugue_b200.dataframe import ArrayDataFrame, df_eq
from oracle import expressions as OX


@pytest.fixture(scope="module")
def e():
    return fa.make_execution_engine("b200")


def _a():
    return ArrayDataFrame([[1, 2], [None, 2], [None, 1], [3, 4], [None, 4]], "a:double,b:int")


def test_filter(e):
    a = _a()
    df_eq(fa.filter(a, col("a").not_null(), engine=e), [[1, 2], [3, 4]], "a:double,b:int", throw=True)
    df_eq(fa.filter(a, col("a").not_null() & (col("b") < 3), engine=e), [[1, 2]], "a:double,b:int", throw=True)
    df_eq(fa.filter(a, col("a") + col("b") == 3, engine=e), [[1, 2]], "a:double,b:int", throw=True)
    with raises(ValueError):
        fa.filter(a, ff.max(col("a")) > 1, engine=e)


def test_select(e):
    a = _a()
    b = fa.select(a, col("b"), (col("b") + 1).alias("c").cast(str), engine=e)
    df_eq(b, [[2, "3"], [2, "3"], [1, "2"], [4, "5"], [4, "5"]], "b:int,c:str", throw=True)
    b = fa.select(a, col("b"), (col("b") + 1).alias("c").cast(str), distinct=True, engine=e)
    df_eq(b, [[2, "3"], [1, "2"], [4, "5"]], "b:int,c:str", throw=True)
    b = fa.select(a, all_cols(), where=col("a") + col("b") == 3, engine=e)
    df_eq(b, [[1, 2]], "a:double,b:int", throw=True)
    b = fa.select(a, col("a"), ff.sum(col("b")).cast(float).alias("b"), engine=e)
    df_eq(b, [[1, 2], [3, 4], [None, 7]], "a:double,b:double", throw=True)
    col_b = ff.sum(col("b"))
    b = fa.select(a, col("a"), col_b.cast(float).alias("c"), having=(col_b >= 7) | (col("a") == 1), engine=e)
    df_eq(b, [[1, 2], [None, 7]], "a:double,c:double", throw=True)
    b = fa.select(a, col("a"), lit(1, "o").cast(str), col_b.cast(float).alias("c"),
                  having=(col_b >= 7) | (col("a") == 1), engine=e)
    df_eq(b, [[1, "1", 2], [None, "1", 7]], "a:double,o:str,c:double", throw=True)


def test_assign(e):
    b = fa.assign(_a(), x=1, b=col("b").cast(str), c=(col("b") + 1).cast(int), engine=e)
    df_eq(b, [[1, "2", 1, 3], [None, "2", 1, 3], [None, "1", 1, 2], [3, "4", 1, 5], [None, "4", 1, 5]],
          "a:double,b:str,x:long,c:long", throw=True)


def test_aggregate_expressions(e):
    a = _a()
    b = fa.aggregate(a, b=ff.max(col("b")), c=(ff.max(col("b")) * 2).cast("int32").alias("c"), engine=e)
    df_eq(b, [[4, 8]], "b:int,c:int", throw=True)
    b = fa.aggregate(a, "a", b=ff.max(col("b")), c=(ff.max(col("b")) * 2).cast("int32").alias("c"), engine=e)
    df_eq(b, [[None, 4, 8], [1, 2, 4], [3, 4, 8]], "a:double,b:int,c:int", throw=True)
    with raises(ValueError):
        fa.aggregate(a, "a", b=ff.max(col("b")), x=1, engine=e)
    with raises(ValueError):
        fa.aggregate(a, "a", engine=e)


def test_strings_and_literals(e):
    s = ArrayDataFrame([["x", 1], ["y", 2], [None, 3], ["x", 4]], "k:str,v:long")
    df_eq(fa.filter(s, col("k") == "x", engine=e), [["x", 1], ["x", 4]], "k:str,v:long", throw=True)
    df_eq(fa.filter(s, col("k") != "x", engine=e), [["y", 2]], "k:str,v:long", throw=True)
    df_eq(fa.filter(s, col("k") == "zz", engine=e), [], "k:str,v:long", throw=True)
    df_eq(fa.filter(s, col("k").is_null() | (col("v") >= 4), engine=e), [[None, 3], ["x", 4]], "k:str,v:long",
          throw=True)
    b = fa.select(s, col("k"), lit("c", "t"), null().cast("double").alias("n"), (col("v") * 1.5).alias("w"), engine=e)
    df_eq(b, [["x", "c", None, 1.5], ["y", "c", None, 3.0], [None, "c", None, 4.5], ["x", "c", None, 6.0]],
          "k:str,t:str,n:double,w:double", throw=True)
    with raises(NotImplementedError):
        fa.filter(s, col("k") < "x", engine=e)


def test_eval_expr_kernel_blocks():
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    n = 100_003  # not a multiple of the 2048-row tile
    a = rng.integers(-1000, 1000, n).astype(np.int32)
    b = rng.standard_normal(n)
    bm = (rng.random(n) > 0.3).astype(np.uint8)
    ta, tb, tm = (torch.from_numpy(x).to(dev) for x in (a, b, bm))
    f64 = lambda v: int(np.float64(v).view(np.uint64))
    N, C, I, R = K.XK_NONE, K.XK_COL, K.XK_IMM, K.XK_REG
    prog = [
        (K.X_MOV, C, 0, 0, 0), (K.X_MUL_I, I, 0, 0, 3), (K.X_SUB_I, I, 0, 0, 7), (K.X_OUT, N, 0, 0, 0),  # a*3-7
        (K.X_MOV, C, 0, 0, 0), (K.X_I2F, N, 0, 0, 0), (K.X_DIV_F, C, 1, 0, 0), (K.X_OUT, N, 1, 0, 0),    # a/b
        (K.X_MOV, C, 0, 0, 0), (K.X_LT_I, I, 0, 0, 10), (K.X_ST, N, 2, 0, 0),                            # t2 = a<10
        (K.X_MOV, C, 1, 0, 0), (K.X_GT_F, I, 0, 0, f64(0.0)), (K.X_AND, R, 2, 0, 0), (K.X_OUT, N, 2, 0, 0),
        (K.X_MOV, C, 1, 0, 0), (K.X_COALESCE, I, 0, 0, f64(2.5)), (K.X_OUT, N, 3, 0, 0),
        (K.X_MOV, C, 0, K.XF_B_I2F, 0), (K.X_RSUB_F, I, 0, 0, f64(0.5)), (K.X_OUT, N, 4, 0, 0),          # 0.5 - a
    ]
    dts = [torch.int16, torch.float64, torch.uint8, torch.float64, torch.float32]
    outs, valids = K.eval_expr(n, dev, [ta, tb], [None, tm], prog, dts, [False, True, True, True, False])
    assert (outs[0].cpu().numpy() == (a.astype(np.int64) * 3 - 7).astype(np.int16)).all()
    v = bm.astype(bool)
    assert (valids[1].cpu().numpy().astype(bool) == v).all()
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = a.astype(np.float64) / b
    got = outs[1].cpu().numpy()
    assert (got[v].view(np.uint64) == ref[v].view(np.uint64)).all() and (got[~v] == 0).all()
    lt, pos = a < 10, b > 0
    is_false = (~lt) | (v & ~pos)
    assert (valids[2].cpu().numpy().astype(bool) == (is_false | v)).all()
    assert (outs[2].cpu().numpy().astype(bool) == (lt & v & pos)).all()
    assert (outs[3].cpu().numpy() == np.where(v, b, 2.5)).all() and valids[3].cpu().numpy().all()
    assert (outs[4].cpu().numpy() == (0.5 - a).astype(np.float32)).all()
    # the numpy model of the machine (tests/_expr_sim.py, used by the CPU compiler tests) agrees bit for bit
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _expr_sim as sim
    souts, svalid = sim.run(n, [a, b], [None, bm], prog, [K.expr_type_of(d) for d in dts])
    for o, so in zip(outs, souts):
        assert np.array_equal(o.cpu().numpy().view(np.uint8), so.view(np.uint8))
    for vv, sv in zip(valids, svalid):
        assert vv is None or np.array_equal(vv.cpu().numpy(), sv)
    for bad in ([(99, N, 0, 0, 0)], [(K.X_ADD_I, N, 0, 0, 0), (K.X_OUT, N, 0, 0, 0)], [(K.X_MOV, C, 7, 0, 0)],
                [(K.X_OUT, N, 3, 0, 0)], [(K.X_ST, N, K.EXPR_NREGS, 0, 0)]):
        with raises(Exception):
            K.eval_expr(n, dev, [ta], [None], bad, [torch.int64], [False])


def _random_table(rng, n):
    a = rng.integers(-50, 50, n).astype(np.int64)
    b = rng.integers(0, 7, n).astype(np.int32)
    x = rng.standard_normal(n)
    x[rng.random(n) < 0.2] = np.nan
    y = rng.standard_normal(n) * 10
    g = pd.array(np.where(rng.random(n) < 0.1, None, rng.integers(0, 20, n)), dtype="Int64")
    p = pd.array(np.where(rng.random(n) < 0.15, None, rng.random(n) < 0.5), dtype="boolean")
    return pd.DataFrame({"a": a, "b": b, "x": x, "y": y, "g": g, "p": p})


def _frames_equal(got: pd.DataFrame, want: pd.DataFrame, sort: bool):
    assert list(got.columns) == list(want.columns)
    assert len(got) == len(want), (len(got), len(want))
    if sort:
        got = got.sort_values(list(got.columns)).reset_index(drop=True)
        want = want.sort_values(list(want.columns)).reset_index(drop=True)
    for c in want.columns:
        w, gt = want[c], got[c]
        wn = w.isna().to_numpy()
        gn = gt.isna().to_numpy()
        assert (wn == gn).all(), c
        if pd.api.types.is_numeric_dtype(w.dtype) or pd.api.types.is_bool_dtype(w.dtype):
            wv = w.to_numpy(dtype="float64", na_value=0.0)[~wn]
            gv = pd.to_numeric(gt[~gn]).to_numpy(dtype="float64")
            if pd.api.types.is_float_dtype(w.dtype):
                assert np.allclose(wv, gv, rtol=1e-12, atol=1e-12, equal_nan=True), c
            else:
                assert (wv == gv).all(), c
        else:
            assert (w[~wn].astype(str).to_numpy() == gt[~gn].astype(str).to_numpy()).all(), c


def test_random_selects_match_oracle(e):
    rng = np.random.default_rng(11)
    pdf = _random_table(rng, 200_000)
    edf = e.to_df(pdf)
    cases = [
        SelectColumns(col("a"), (col("a") * col("b") - 3).alias("m"), (col("x") / col("y")).alias("d"),
                      (col("a") / col("b")).alias("q"), (-col("x")).alias("nx"), (-col("b")).alias("nb")),
        SelectColumns(((col("x") > 0) & col("p")).alias("k1"), ((col("x") > 0) | col("p")).alias("k2"),
                      (~col("p")).alias("k3"), (col("x").is_null() | (col("a") >= col("b"))).alias("k4"),
                      (col("g") == col("b")).alias("k5"), (col("g") != 3).alias("k6")),
        SelectColumns(ff.coalesce(col("x"), col("y")).alias("c1"), ff.coalesce(col("g"), -1).alias("c2"),
                      ff.coalesce(col("g"), col("x"), 0.5).alias("c3"), (col("a") + 1.5).cast(int).alias("c4"),
                      col("x").cast("int").alias("c5"), col("b").cast(float).alias("c6"),
                      (col("a") > 0).cast(int).alias("c7"), col("g").cast(bool).alias("c8")),
    ]
    for sel in cases:
        got = e.select(edf, sel).as_pandas()
        want = OX.select(pdf, sel)
        _frames_equal(got, want, sort=False)
    # filters (row order preserved)
    for cond in [(col("x") > 0.5) & (col("g") < 10), col("p") | col("x").is_null(), ~(col("a") * 2 <= col("b")),
                 (col("x") + col("y") > 0) | (col("g").is_null() & col("p"))]:
        got = e.filter(edf, cond).as_pandas()
        want = OX.filter_rows(pdf, cond)
        assert len(got) == len(want)
        _frames_equal(got, want, sort=False)


def test_random_aggregating_selects_match_oracle(e):
    rng = np.random.default_rng(12)
    pdf = _random_table(rng, 300_000)
    edf = e.to_df(pdf)
    s = ff.sum(col("x"))
    cases = [
        (SelectColumns(col("g"), col("b"), ff.sum(col("x") * col("y")).alias("sxy"), ff.count(all_cols()).alias("n"),
                       ff.count(col("x")).alias("nx"), ff.max(col("a") + col("b")).alias("mx"),
                       ff.avg(col("y")).alias("av")), None, None),
        (SelectColumns((col("a") - col("a") / 10 * 10 + col("b")).cast(int).alias("k"), ff.min(col("y")).alias("lo"),
                       (ff.max(col("y")) - ff.min(col("y"))).alias("span")), col("x").not_null(), None),
        (SelectColumns(col("g"), (s / ff.count(col("x"))).alias("mean"), lit(7, "seven")), None,
         (s > 0) & col("g").not_null()),
        (SelectColumns(ff.sum(col("a")).alias("sa"), ff.count(all_cols()).alias("n"), ff.avg(col("x")).alias("ax")),
         col("p"), None),
    ]
    for sel, where, having in cases:
        got = e.select(edf, sel, where=where, having=having).as_pandas()
        want = OX.select(pdf, sel, where=where, having=having)
        _frames_equal(got, want, sort=True)


def test_large_expression_is_split_over_launches(e):
    pdf = pd.DataFrame({"a": np.arange(5000, dtype=np.int64), "x": np.linspace(0, 1, 5000)})
    cols = [((col("a") + i) * (col("x") - i) + (col("a") - i) * 2).alias(f"c{i}") for i in range(20)]
    got = e.select(e.to_df(pdf), SelectColumns(*cols)).as_pandas()
    for i in range(20):
        ref = (pdf["a"] + i) * (pdf["x"] - i) + (pdf["a"] - i) * 2
        assert np.allclose(got[f"c{i}"].to_numpy(), ref.to_numpy(), rtol=1e-13)


def test_sql_text_reaches_the_device_path(e):
    rng = np.random.default_rng(21)
    n = 200_000
    pdf = pd.DataFrame({"key": rng.integers(0, 300, n), "v0": rng.standard_normal(n), "v1": rng.standard_normal(n),
                        "tag": rng.choice(["x", "y", "z"], n)})
    got = fa.raw_sql("SELECT key, SUM(v0 * 2) AS s, COUNT(*) AS c, AVG(v1) a FROM", pdf,
                     "WHERE v1 > 0 AND tag <> 'z' GROUP BY key HAVING COUNT(*) > 5 ORDER BY key DESC LIMIT 50",
                     engine=e, as_local=True)
    f = pdf[(pdf.v1 > 0) & (pdf.tag != "z")]
    want = f.assign(v02=f.v0 * 2).groupby("key").agg(s=("v02", "sum"), c=("v0", "size"), a=("v1", "mean")).reset_index()
    want = want[want.c > 5].sort_values("key", ascending=False).head(50).reset_index(drop=True)
    assert list(got.columns) == ["key", "s", "c", "a"]
    assert np.array_equal(got.key.to_numpy(), want.key.to_numpy()) and np.array_equal(got.c.to_numpy(), want.c.to_numpy())
    assert np.allclose(got.s.to_numpy(), want.s.to_numpy(), rtol=1e-9) and np.allclose(got.a.to_numpy(), want.a.to_numpy(), rtol=1e-9)
    # projection with expressions, CAST, IN / BETWEEN, IS NULL, DISTINCT
    got = fa.raw_sql("SELECT key, CAST(v0 * 10 AS long) AS q, v0 IS NULL AS isn FROM", pdf,
                     "WHERE key IN (1, 2, 3) AND v1 NOT BETWEEN -0.5 AND 0.5", engine=e, as_local=True)
    f = pdf[pdf.key.isin([1, 2, 3]) & ~pdf.v1.between(-0.5, 0.5)]
    assert np.array_equal(got.key.to_numpy(), f.key.to_numpy())
    assert np.array_equal(got.q.to_numpy(), np.trunc(f.v0.to_numpy() * 10).astype(np.int64))
    assert not got.isn.any()
    got = fa.raw_sql("SELECT DISTINCT tag, key - key AS z FROM", pdf, engine=e, as_local=True)
    assert sorted(got.tag.tolist()) == ["x", "y", "z"] and (got.z == 0).all()
    # GROUP BY key that is not selected rides along as a hidden column
    got = fa.raw_sql("SELECT MAX(v0) AS m FROM", pdf, "GROUP BY key", engine=e, as_local=True)
    assert np.allclose(np.sort(got.m.to_numpy()), np.sort(pdf.groupby("key").v0.max().to_numpy()))
    with raises(ValueError):
        fa.raw_sql("SELECT key, tag, SUM(v0) AS s FROM", pdf, "GROUP BY key", engine=e)
    with raises(NotImplementedError):
        fa.raw_sql("SELECT key FROM", pdf, "UNION SELECT key FROM", pdf, engine=e)


def test_first_last_aggregates(e):
    a = ArrayDataFrame([[1, None, "x"], [1, 5.0, None], [1, 7.0, "y"], [2, None, None], [None, 3.0, "z"], [None, 4.0, "w"]],
                       "k:long,v:double,s:str")
    b = fa.aggregate(a, "k", f=ff.first(col("v")), l=ff.last(col("v")), fs=ff.first(col("s")), ls=ff.last(col("s")),
                     engine=e)
    df_eq(b, [[1, 5.0, 7.0, "x", "y"], [2, None, None, None, None], [None, 3.0, 4.0, "z", "w"]],
          "k:long,f:double,l:double,fs:str,ls:str", throw=True)
    b = fa.select(a, (ff.last(col("v")) - ff.first(col("v"))).alias("d"), engine=e)
    df_eq(b, [[-1.0]], "d:double", throw=True)
    got = fa.raw_sql("SELECT k, FIRST(v) AS f FROM", a, "WHERE v IS NOT NULL GROUP BY k", engine=e, as_fugue=True)
    df_eq(got, [[1, 5.0], [None, 3.0]], "k:long,f:double", throw=True)


def test_full_size_properties_100m_rows(e):
    """Size-independent checks at the BASELINE row count: linearity of an integer projection, the
    three-way split of a nullable predicate, and filter == mask semantics."""
    from fugue_b200.dataframe import B200DataFrame
    from fugue_b200.table import B200Table

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(9)
    n = 100_000_000
    a = torch.randint(-1000, 1000, (n,), dtype=torch.int64, device=dev, generator=g)
    x = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    xv = (torch.rand(n, device=dev, generator=g) > 0.1).to(torch.uint8)        # 10 % NULLs in x
    t = B200DataFrame(B200Table("a:long,x:double", [a, x], [None, xv]))
    out = e.select(t, SelectColumns((col("a") * 2 + 1).alias("y"), (col("x") > 0).alias("p"),
                                    ff.coalesce(col("x"), 0.0).alias("c"))).native
    assert int(out.column("y").sum()) == 2 * int(a.sum()) + n                    # exact integer arithmetic
    pv = out.valid[out.schema.index_of_key("p")]
    assert pv is not None and torch.equal(pv, xv)                                # NULL propagates through >
    assert out.valid[out.schema.index_of_key("c")] is None                       # COALESCE(x, 0.0) is never NULL
    assert torch.equal(out.column("c"), torch.where(xv.bool(), x, torch.zeros_like(x)))
    n_true = e.filter(t, col("x") > 0).count()
    n_false = e.filter(t, ~(col("x") > 0)).count()
    n_null = e.filter(t, (col("x") > 0).is_null()).count()
    assert n_true + n_false + n_null == n and n_null == n - int(xv.sum())
    assert n_true == int(((x > 0) & xv.bool()).sum())
    kept = e.filter(t, (col("a") >= 0) & col("x").not_null()).native
    m = (a >= 0) & xv.bool()
    assert torch.equal(kept.column("a"), a[m]) and torch.equal(kept.column("x"), x[m])   # order preserved


def test_count_distinct(e):
    a = ArrayDataFrame([[1, 5.0, 1], [1, 5.0, 2], [1, None, 3], [2, 7.0, 4], [2, 8.0, 5], [None, 9.0, 6]],
                       "k:long,x:double,y:long")
    b = fa.select(a, col("k"), ff.count_distinct(col("x")).alias("d"), ff.sum(col("y")).alias("s"),
                  ff.avg(col("y")).alias("m"), engine=e)
    df_eq(b, [[1, 1, 6, 2.0], [2, 2, 9, 4.5], [None, 1, 6, 6.0]], "k:long,d:long,s:long,m:double", throw=True)
    b = fa.select(fa.union(a, a, distinct=False, engine=e), ff.count_distinct(all_cols()).alias("n"), engine=e)
    df_eq(b, [[6]], "n:long", throw=True)
    rng = np.random.default_rng(31)
    pdf = _random_table(rng, 150_000)
    edf = e.to_df(pdf)
    for sel, where in [
        (SelectColumns(col("g"), ff.count_distinct(col("b")).alias("d"), ff.count(all_cols()).alias("n"),
                       ff.max(col("y")).alias("mx")), None),
        (SelectColumns(ff.count_distinct(col("a") + col("b")).alias("d"), ff.avg(col("x")).alias("ax")), col("p")),
        (SelectColumns(col("b"), (ff.count_distinct(col("g")) * 2).alias("d2")), col("x").not_null()),
    ]:
        got = e.select(edf, sel, where=where).as_pandas()
        want = OX.select(pdf, sel, where=where)
        _frames_equal(got, want, sort=True)
    got = fa.raw_sql("SELECT b, COUNT(DISTINCT g) AS d FROM", pdf, "GROUP BY b", engine=e, as_local=True)
    want = pdf.groupby("b")["g"].nunique().reset_index(name="d")
    assert np.array_equal(got.sort_values("b")["d"].to_numpy(), want["d"].to_numpy())
    with raises(NotImplementedError):
        e.select(edf, SelectColumns(ff.count_distinct(col("a")).alias("x"), ff.count_distinct(col("b")).alias("y")))
