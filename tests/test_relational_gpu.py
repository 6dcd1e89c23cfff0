"""Set operations, dropna/fillna/sample and IO on the B200 engine: literal expectations of
fugue_test/execution_suite.py:545-760, 1100-1300 (union / subtract / intersect / distinct / dropna /
fillna / sample / save-load)."""
import os

import numpy as np
import pandas as pd
import pytest
from pytest import raises

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from fugue_b200 import api as fa
from fugue_b200.dataframe import ArrayDataFrame, df_eq


@pytest.fixture(scope="module")
def e():
    return fa.make_execution_engine("b200")


S = "a:double,b:double,c:int"


def test_union(e):
    a = fa.as_fugue_engine_df(e, [[1, 2, 3], [4, None, 6]], S)
    b = fa.as_fugue_engine_df(e, [[1, 2, 33], [4, None, 6]], S)
    df_eq(fa.union(a, b, engine=e), [[1, 2, 3], [4, None, 6], [1, 2, 33]], S, throw=True)
    c = fa.union(a, b, distinct=False, engine=e)
    df_eq(c, [[1, 2, 3], [4, None, 6], [1, 2, 33], [4, None, 6]], S, throw=True)
    d = fa.union(a, b, c, distinct=False, engine=e)
    assert d.count() == 8
    with raises(ValueError):
        fa.union(a, fa.as_fugue_engine_df(e, [[1]], "x:int"), engine=e)


def test_subtract_intersect_distinct(e):
    a = fa.as_fugue_engine_df(e, [[1, 2, 3], [1, 2, 3], [4, None, 6]], S)
    b = fa.as_fugue_engine_df(e, [[1, 2, 33], [4, None, 6]], S)
    df_eq(fa.subtract(a, b, engine=e), [[1, 2, 3]], S, throw=True)
    x = fa.as_fugue_engine_df(e, [[1, 2, 33]], S)
    y = fa.as_fugue_engine_df(e, [[4, None, 6]], S)
    df_eq(fa.subtract(a, x, y, engine=e), [[1, 2, 3]], S, throw=True)
    with raises(NotImplementedError):
        fa.subtract(a, b, distinct=False, engine=e)
    a = fa.as_fugue_engine_df(e, [[1, 2, 3], [4, None, 6], [4, None, 6]], S)
    b = fa.as_fugue_engine_df(e, [[1, 2, 33], [4, None, 6], [4, None, 6], [4, None, 6]], S)
    df_eq(fa.intersect(a, b, engine=e), [[4, None, 6]], S, throw=True)
    y = fa.as_fugue_engine_df(e, [[4, None, 6], [4, None, 6], [4, None, 6]], S)
    df_eq(fa.intersect(a, x, y, engine=e), [], S, throw=True)
    a = fa.as_fugue_engine_df(e, [[4, None, 6], [1, 2, 3], [4, None, 6]], S)
    df_eq(fa.distinct(a, engine=e), [[4, None, 6], [1, 2, 3]], S, throw=True)
    s = fa.as_fugue_engine_df(e, [["x", 1], ["y", 1], ["x", 1], [None, 2], [None, 2]], "k:str,v:int")
    df_eq(fa.distinct(s, engine=e), [["x", 1], ["y", 1], [None, 2]], "k:str,v:int", throw=True)


def test_distinct_large_matches_pandas(e):
    rng = np.random.default_rng(2)
    pdf = pd.DataFrame({"a": rng.integers(0, 50, 100_000), "b": rng.integers(0, 40, 100_000).astype("int32")})
    got = fa.distinct(pdf, engine=e, as_local=True)
    exp = pdf.drop_duplicates()
    assert len(got) == len(exp)
    assert sorted(map(tuple, got.values.tolist())) == sorted(map(tuple, exp.values.tolist()))


def test_dropna_fillna(e):
    s = "a:double,b:double,c:double"
    a = fa.as_fugue_engine_df(e, [[4, None, 6], [1, 2, 3], [4, None, None]], s)
    df_eq(fa.dropna(a, engine=e), [[1, 2, 3]], s, throw=True)
    df_eq(fa.dropna(a, how="all", engine=e), [[4, None, 6], [1, 2, 3], [4, None, None]], s, throw=True)
    df_eq(fa.dropna(a, how="any", thresh=2, engine=e), [[4, None, 6], [1, 2, 3]], s, throw=True)
    df_eq(fa.dropna(a, how="any", subset=["a", "c"], engine=e), [[4, None, 6], [1, 2, 3]], s, throw=True)
    df_eq(fa.dropna(a, how="any", thresh=1, subset=["a", "c"], engine=e),
          [[4, None, 6], [1, 2, 3], [4, None, None]], s, throw=True)
    df_eq(fa.fillna(a, value=1, engine=e), [[4, 1, 6], [1, 2, 3], [4, 1, 1]], s, throw=True)
    d = fa.fillna(a, {"b": 99, "c": -99}, engine=e)
    df_eq(d, [[4, 99, 6], [1, 2, 3], [4, 99, -99]], s, throw=True)
    df_eq(fa.fillna(a, value=-99, subset=["c"], engine=e), [[4, None, 6], [1, 2, 3], [4, None, -99]], s, throw=True)
    df_eq(fa.fillna(a, {"b": 99, "c": -99}, subset=["c"], engine=e), d, throw=True)
    raises(ValueError, lambda: fa.fillna(a, {"b": None, "c": "99"}, engine=e))
    raises(ValueError, lambda: fa.fillna(a, None, engine=e))
    t = fa.as_fugue_engine_df(e, [["x", 1], [None, 2]], "k:str,v:int")
    df_eq(fa.fillna(t, {"k": "zz"}, engine=e), [["x", 1], ["zz", 2]], "k:str,v:int", throw=True)


def test_sample(e):
    a = fa.as_fugue_engine_df(e, [[x] for x in range(100)], "a:int")
    with raises(ValueError):
        fa.sample(a, engine=e)
    with raises(ValueError):
        fa.sample(a, n=90, frac=0.9, engine=e)
    f = fa.sample(a, frac=0.8, replace=False, engine=e)
    g = fa.sample(a, frac=0.8, replace=True, engine=e)
    h = fa.sample(a, frac=0.8, seed=1, engine=e)
    h2 = fa.sample(a, frac=0.8, seed=1, engine=e)
    i = fa.sample(a, frac=0.8, seed=2, engine=e)
    assert not df_eq(f, g, throw=False)
    df_eq(h, h2, throw=True)
    assert not df_eq(h, i, throw=False)
    assert abs(len(i.as_array()) - 80) < 10
    assert len(set(x[0] for x in f.as_array())) == 80            # without replacement: no duplicates


def test_save_and_load_parquet_csv_json(e, tmp_path):
    b = ArrayDataFrame([[6, 1.5, "x"], [2, 7.25, None]], "c:int,a:double,s:str")
    path = os.path.join(tmp_path, "a", "b.parquet")
    fa.save(b, path, engine=e)
    c = fa.load(path, engine=e, as_fugue=True)
    df_eq(c, [[6, 1.5, "x"], [2, 7.25, None]], "c:int,a:double,s:str", throw=True)
    c = fa.load(path, columns=["a", "c"], engine=e, as_fugue=True)
    df_eq(c, [[1.5, 6], [7.25, 2]], "a:double,c:int", throw=True)
    fa.save(c, path, engine=e)                                   # overwrite
    with raises(FileExistsError):
        fa.save(c, path, mode="error", engine=e)
    p2 = os.path.join(tmp_path, "x.csv")
    fa.save(ArrayDataFrame([[1, 2], [3, 4]], "a:long,b:long"), p2, header=True, engine=e)
    df_eq(fa.load(p2, header=True, infer_schema=True, engine=e, as_fugue=True), [[1, 2], [3, 4]], "a:long,b:long",
          throw=True)
    df_eq(fa.load(p2, header=True, engine=e, as_fugue=True), [["1", "2"], ["3", "4"]], "a:str,b:str", throw=True)
    df_eq(fa.load(p2, header=True, columns="b:long,a:double", engine=e, as_fugue=True), [[2, 1.0], [4, 3.0]],
          "b:long,a:double", throw=True)
    p3 = os.path.join(tmp_path, "x.json")
    fa.save(ArrayDataFrame([[1, 2], [3, 4]], "a:long,b:long"), p3, engine=e)
    df_eq(fa.load(p3, engine=e, as_fugue=True), [[1, 2], [3, 4]], "a:long,b:long", throw=True)


def test_take(e):
    ps = dict(by=["a"], presort="b DESC,c DESC")
    ps2 = dict(by=["c"], presort="b ASC")
    s = "a:str,b:int,c:long"
    a = fa.as_fugue_engine_df(e, [["a", 2, 3], ["a", 3, 4], ["b", 1, 2], ["b", 2, 2], [None, 4, 2], [None, 2, 1]], s)
    df_eq(fa.take(a, n=1, presort="b desc", engine=e), [[None, 4, 2]], s, throw=True)
    df_eq(fa.take(a, n=2, presort="a desc", na_position="first", engine=e), [[None, 4, 2], [None, 2, 1]], s, throw=True)
    df_eq(fa.take(a, n=1, presort="a asc, b desc", partition=ps, engine=e),
          [["a", 3, 4], ["b", 2, 2], [None, 4, 2]], s, throw=True)
    df_eq(fa.take(a, n=1, presort=None, partition=ps2, engine=e),
          [["a", 2, 3], ["a", 3, 4], ["b", 1, 2], [None, 2, 1]], s, throw=True)
    df_eq(fa.take(a, n=2, presort="a desc", na_position="last", engine=e), [["b", 1, 2], ["b", 2, 2]], s, throw=True)
    df_eq(fa.take(a, n=2, presort="a", na_position="first", engine=e), [[None, 4, 2], [None, 2, 1]], s, throw=True)
    a = fa.as_fugue_engine_df(e, [["a", 2, 3], [None, 4, 2], [None, 2, 1]], s)
    j = fa.take(a, n=2, partition="a", presort=None, engine=e)
    df_eq(j, [["a", 2, 3], [None, 4, 2], [None, 2, 1]], s, throw=True)
    i = fa.take(a, n=1, partition="a", presort=None, engine=e)
    assert i.count() == 2
    raises(ValueError, lambda: fa.take(a, n=0.5, presort=None, engine=e))


def test_device_sort_matches_pandas(e):
    from collections import OrderedDict

    from fugue_b200 import sort as S

    rng = np.random.default_rng(8)
    n = 300_000
    pdf = pd.DataFrame({"i": rng.integers(-10**12, 10**12, n), "f": rng.standard_normal(n),
                        "s": rng.integers(-3, 3, n).astype("int32"), "u": rng.integers(0, 255, n).astype("uint8")})
    pdf.loc[rng.integers(0, n, 3000), "f"] = np.nan
    t = fa.as_fugue_engine_df(e, pdf, "i:long,f:double,s:int,u:ubyte").native
    for sorts, napos in [(OrderedDict(i=True), "last"), (OrderedDict(f=False), "first"),
                         (OrderedDict([("s", True), ("u", False), ("f", True)]), "last")]:
        got = S.sort_table(t, sorts, napos).to_pandas()
        exp = pdf.sort_values(list(sorts.keys()), ascending=list(sorts.values()), na_position=napos,
                              kind="stable").reset_index(drop=True)
        pd.testing.assert_frame_equal(got, exp, check_exact=True)


def test_presort_and_logical_partitions_for_device_functions(e):
    """select-top per logical partition written as a device function (fugue_test/execution_suite.py:
    225-256 does it with cursor.row on the host)."""
    from fugue_b200.table import B200Table
    from fugue_b200.sort import take_rows

    def first_of_each_logical_partition(t: B200Table) -> B200Table:
        return take_rows(t, t.logical_offsets[:-1])

    o = ArrayDataFrame([[1, 2], [None, 2], [None, 1], [3, 4], [None, 4]], "a:double,b:int")
    c = fa.transform(o, first_of_each_logical_partition, schema="*", partition=dict(by=["a"], presort="b"), engine=e)
    df_eq(c, [[None, 1], [1, 2], [3, 4]], "a:double,b:int", throw=True)
    c = fa.transform(o, first_of_each_logical_partition, schema="*", partition=dict(by=["a"], presort="b DESC", num=3),
                     engine=e)
    df_eq(c, [[None, 4], [1, 2], [3, 4]], "a:double,b:int", throw=True)
    rng = np.random.default_rng(12)
    pdf = pd.DataFrame({"k": rng.integers(0, 5000, 200_000), "v": rng.standard_normal(200_000)})
    got = fa.transform(pdf, first_of_each_logical_partition, schema="*", partition=dict(by="k", presort="v desc"),
                       engine=e, as_local=True)
    exp = pdf.sort_values("v", ascending=False).groupby("k").head(1)
    assert sorted(map(tuple, got.values.tolist())) == sorted(map(tuple, exp.values.tolist()))


@pytest.mark.gpu
def test_tma_pull_and_dma_runs_copy_any_8_byte_alignment():
    """The exchange's two run copiers (persistent TMA pull kernel, copy engines) on local memory:
    runs whose source / destination are only 8-byte aligned, odd lengths, empty and tiny runs."""
    import torch

    from fugue_b200 import kernels as K

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    src = torch.randint(-(2**62), 2**62, (3_000_000,), dtype=torch.int64, device=dev, generator=g)
    runs = [(0, 0, 1_000_001), (1_000_001, 1_000_001, 4096), (1_004_097, 1_004_098, 0), (1_004_097, 1_004_099, 1),
            (1_100_001, 1_100_000, 777_777), (2_000_000, 2_000_003, 2), (2_100_000, 2_100_001, 300_003)]
    for fn in (lambda s, d, n: K.pull_runs_tma(dev, s, d, n, 7), lambda s, d, n: K.copy_runs_dma(dev, s, d, n)):
        dst = torch.zeros_like(src)
        fn([src.data_ptr() + 8 * a for a, _, _ in runs], [dst.data_ptr() + 8 * b for _, b, _ in runs],
           [8 * n for _, _, n in runs])
        torch.cuda.synchronize()
        exp = torch.zeros_like(src)
        for a, b, n in runs:
            exp[b:b + n] = src[a:a + n]
        assert torch.equal(dst, exp)
