"""K6 hash group-by vs the oracle (pandas groupby(dropna=False).agg) and the reference's literal
expectations (fugue_test/execution_suite.py:177-206 test_aggregate, builtin_suite.py:937-949)."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from fugue_b200 import api as fa
from fugue_b200.column import all_cols, col, functions as ff
from fugue_b200.dataframe import ArrayDataFrame, df_eq
from oracle import native_engine as ora


@pytest.fixture(scope="module")
def engine():
    return fa.make_execution_engine("b200")


def test_aggregate_reference_literals(engine):
    a = ArrayDataFrame([[1, 2], [None, 2], [None, 1], [3, 4], [None, 4]], "a:double,b:int")
    b = fa.aggregate(a, b=ff.max(col("b")), engine=engine)
    df_eq(b, [[4]], "b:int", throw=True)
    b = fa.aggregate(a, "a", b=ff.max(col("b")), engine=engine)
    df_eq(b, [[None, 4], [1, 2], [3, 4]], "a:double,b:int", throw=True)
    with pytest.raises(ValueError):
        fa.aggregate(a, "a", b=ff.max(col("b")), x=1, engine=engine)
    with pytest.raises(ValueError):
        fa.aggregate(a, "a", engine=engine)


def test_sum_count_matches_oracle_1e6(engine):
    rng = np.random.default_rng(0)
    n = 1_000_000
    pdf = pd.DataFrame({"key": rng.integers(-(2**62), 2**62, 200_000)[rng.integers(0, 200_000, n)],
                        "v0": rng.standard_normal(n), "i": rng.integers(-1000, 1000, n)})
    res = fa.aggregate(pdf, "key", s=ff.sum(col("v0")), c=ff.count(all_cols()), si=ff.sum(col("i")),
                       mn=ff.min(col("v0")), mx=ff.max(col("i")), av=ff.avg(col("i")),
                       engine=engine, as_local=True)
    exp = ora.aggregate(pdf, ["key"], {"s": ("v0", "sum"), "c": ("*", "count"), "si": ("i", "sum"),
                                       "mn": ("v0", "min"), "mx": ("i", "max"), "av": ("i", "avg")})
    got = res.sort_values("key").reset_index(drop=True)
    exp = exp.sort_values("key").reset_index(drop=True)
    assert len(got) == len(exp)
    for c in ("key", "c", "si", "mx"):
        assert np.array_equal(got[c].to_numpy(), exp[c].to_numpy()), c       # integers: bit exact
    assert np.array_equal(got["mn"].to_numpy(), exp["mn"].to_numpy())          # min moves a value: exact
    rel = np.abs(got["s"].to_numpy() - exp["s"].to_numpy()) / np.maximum(np.abs(exp["s"].to_numpy()), 1e-300)
    assert rel.max() <= 1e-9, rel.max()                                         # fp64 SUM within 1e-9 relative
    assert np.allclose(got["av"].to_numpy(), exp["av"].to_numpy(), rtol=1e-12, atol=0)


def test_null_keys_null_values_and_special_key(engine):
    a = ArrayDataFrame([[None, 1.0], [None, None], [-1, 2.0], [-1, None], [5, None], [5, None], [7, 4.5]],
                       "k:long,v:double")
    res = fa.aggregate(a, "k", s=ff.sum(col("v")), c=ff.count(all_cols()), cv=ff.count(col("v")),
                       m=ff.max(col("v")), engine=engine)
    # key -1 is the all-ones bit pattern (the table's EMPTY marker) and must still be a group;
    # SUM/MAX over only-NULL values is NULL; COUNT(v) skips NULLs
    df_eq(res, [[None, 1.0, 2, 1, 1.0], [-1, 2.0, 2, 1, 2.0], [5, None, 2, 0, None], [7, 4.5, 1, 1, 4.5]],
          "k:long,s:double,c:long,cv:long,m:double", throw=True)


def test_low_cardinality_and_skew(engine):
    rng = np.random.default_rng(3)
    n = 2_000_000
    pdf = pd.DataFrame({"key": np.minimum(rng.zipf(1.3, n), 1000).astype("int64"), "v0": rng.standard_normal(n)})
    res = fa.aggregate(pdf, "key", s=ff.sum(col("v0")), c=ff.count(all_cols()), engine=engine, as_local=True)
    exp = ora.aggregate_sum_count(pdf, ["key"], "v0")
    got = res.sort_values("key").reset_index(drop=True)
    assert np.array_equal(got["key"].to_numpy(), exp["key"].to_numpy())
    assert np.array_equal(got["c"].to_numpy(), exp["c"].to_numpy())
    assert np.max(np.abs(got["s"].to_numpy() - exp["s"].to_numpy()) / np.abs(exp["s"].to_numpy())) <= 1e-9


def test_other_key_types_and_empty(engine):
    a = ArrayDataFrame([[1.5, 1], [0.0, 2], [-0.0, 3], [None, 4], [1.5, 5]], "k:double,v:int")
    res = fa.aggregate(a, "k", s=ff.sum(col("v")), engine=engine)
    df_eq(res, [[None, 4], [0.0, 5], [1.5, 6]], "k:double,s:long", throw=True)
    a = ArrayDataFrame([["x", 1], ["y", 2], ["x", 3], [None, 9]], "k:str,v:int")
    res = fa.aggregate(a, "k", s=ff.sum(col("v")), c=ff.count(all_cols()), engine=engine)
    df_eq(res, [["x", 4, 2], ["y", 2, 1], [None, 9, 1]], "k:str,s:long,c:long", throw=True)
    e = ArrayDataFrame([], "k:long,v:double")
    res = fa.aggregate(e, "k", s=ff.sum(col("v")), engine=engine)
    df_eq(res, [], "k:long,s:double", throw=True)
    a = ArrayDataFrame([[1, 2], [3, 4]], "k:int,v:short")
    res = fa.aggregate(a, "k", m=ff.min(col("v")), engine=engine)
    df_eq(res, [[1, 2], [3, 4]], "k:int,m:short", throw=True)


def test_full_size_properties_100m_rows_10m_keys():
    """BASELINE config 4 shape per GPU (100 M rows, 10 M distinct keys): checksum properties."""
    from fugue_b200 import kernels as K

    dev = torch.device("cuda", 0)
    n, nk = 100_000_000, 10_000_000
    g = torch.Generator(device=dev).manual_seed(1)
    idx = torch.randint(0, nk, (n,), dtype=torch.int64, device=dev, generator=g)
    keys = idx * 0x9E3779B97F4A7C15 % (1 << 62)           # a fixed bijection-ish scramble of dense ids
    v = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    gk, _, (s, c), ng = K.groupby_u64(keys, None, [v.view(torch.int64), None], [None, None],
                                      [K.AGG_SUM_F64, K.AGG_COUNT])
    assert ng == int(torch.unique(keys).numel())
    assert int(c.sum()) == n                                # COUNT(*) adds up to the row count
    total = float(s.view(torch.float64).sum())
    ref = float(v.sum())
    assert abs(total - ref) <= 1e-9 * max(1.0, float(v.abs().sum())) # sum of sums == sum
    assert int(torch.unique(gk).numel()) == ng              # every group appears once


def test_sql_engine_group_by_and_join(engine):
    """BASELINE configs 4/5 as SQL text (fugue/sql/_visitors.py:743-766 forwards exactly such text)."""
    from fugue_b200.sql import StructuredRawSQL

    rng = np.random.default_rng(5)
    n = 100_000
    pdf = pd.DataFrame({"key": rng.integers(0, 5000, n), "v0": rng.standard_normal(n)})
    res = fa.raw_sql("SELECT key, SUM(v0) AS s, COUNT(*) AS c FROM", pdf, "GROUP BY key", engine=engine,
                     as_local=True)
    exp = ora.aggregate_sum_count(pdf, ["key"], "v0")
    got = res.sort_values("key").reset_index(drop=True)
    assert list(got.columns) == ["key", "s", "c"]
    assert np.array_equal(got["key"], exp["key"]) and np.array_equal(got["c"], exp["c"])
    assert np.max(np.abs(got["s"] - exp["s"]) / np.abs(exp["s"])) <= 1e-9
    # pieces form (fugue/collections/sql.py:48-151) and a join
    l = pd.DataFrame({"key": rng.integers(0, 1000, 3000), "lv": rng.standard_normal(3000)})
    r = pd.DataFrame({"key": rng.permutation(1000), "rv": rng.standard_normal(1000)})
    st = StructuredRawSQL([(False, "SELECT * FROM"), (True, "a"), (False, "INNER JOIN"), (True, "b"),
                           (False, "ON a.key = b.key")])
    out = engine.sql_engine.select({"a": l, "b": r}, st)
    exp = ora.join(l, r, "inner")
    df_eq(out, exp.values.tolist(), "key:long,lv:double,rv:double", throw=True)
    out = fa.raw_sql("SELECT COUNT(*) AS n, MAX(lv) AS m FROM", l, engine=engine, as_local=True)
    assert out.values.tolist() == [[3000, l.lv.max()]]
    out = fa.raw_sql("SELECT key FROM", l, "WHERE key > 3", engine=engine, as_local=True)
    assert np.array_equal(out["key"].to_numpy(), l.key[l.key > 3].to_numpy())
    with pytest.raises(NotImplementedError):
        fa.raw_sql("SELECT key FROM (SELECT * FROM", l, ")", engine=engine)


def test_multi_column_group_by(engine):
    rng = np.random.default_rng(9)
    n = 300_000
    pdf = pd.DataFrame({"a": rng.integers(0, 300, n), "b": rng.integers(-5, 5, n).astype("int32"),
                        "c": np.round(rng.standard_normal(n), 1), "v": rng.standard_normal(n)})
    pdf.loc[rng.integers(0, n, 2000), "c"] = np.nan                  # NULLs in one of the key columns
    res = fa.aggregate(fa.as_fugue_engine_df(engine, pdf, "a:long,b:int,c:double,v:double"), ["a", "b", "c"],
                       s=ff.sum(col("v")), n=ff.count(all_cols()), engine=engine, as_local=True)
    got = res.as_pandas().sort_values(["a", "b", "c"]).reset_index(drop=True)
    exp = ora.aggregate(pdf, ["a", "b", "c"], {"s": ("v", "sum"), "n": ("*", "count")})
    exp = exp.sort_values(["a", "b", "c"]).reset_index(drop=True)
    assert len(got) == len(exp)
    for k in ("a", "b", "n"):
        assert np.array_equal(got[k].to_numpy(), exp[k].to_numpy()), k
    assert np.array_equal(got["c"].isna().to_numpy(), exp["c"].isna().to_numpy())
    assert np.allclose(got["c"].fillna(0), exp["c"].fillna(0), rtol=0, atol=0)
    assert np.max(np.abs(got["s"] - exp["s"]) / np.maximum(np.abs(exp["s"]), 1e-300)) <= 1e-9
    # string + int keys
    a = ArrayDataFrame([["x", 1, 1.0], ["x", 2, 2.0], ["y", 1, 3.0], ["x", 1, 4.0], [None, 1, 5.0]], "k:str,j:int,v:double")
    res = fa.aggregate(a, ["k", "j"], s=ff.sum(col("v")), engine=engine)
    df_eq(res, [["x", 1, 5.0], ["x", 2, 2.0], ["y", 1, 3.0], [None, 1, 5.0]], "k:str,j:int,s:double", throw=True)


def test_batched_table_initialisation_gives_the_same_groups(monkeypatch):
    """fb_groupby_u64 with partition offsets: init + aggregate in L2-sized batches of regions."""
    from fugue_b200 import kernels as K

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    n = 6_000_000
    keys = torch.randint(0, 700_000, (n,), dtype=torch.int64, device=dev, generator=g)
    v = torch.randint(-1000, 1000, (n,), dtype=torch.int64, device=dev, generator=g)
    ref = K.groupby_u64(keys, None, [v, None], [None, None], [K.AGG_SUM_I64, K.AGG_COUNT])
    monkeypatch.setattr(K, "GROUPBY_BATCHED", True)
    got = K.groupby_u64(keys, None, [v, None], [None, None], [K.AGG_SUM_I64, K.AGG_COUNT])
    assert ref[3] == got[3]
    o1, o2 = torch.argsort(ref[0][:ref[3]]), torch.argsort(got[0][:got[3]])
    assert torch.equal(ref[0][:ref[3]][o1], got[0][:got[3]][o2])
    for a, b in zip(ref[2], got[2]):
        assert torch.equal(a[:ref[3]][o1], b[:got[3]][o2])
