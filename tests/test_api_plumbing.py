"""The functional API (``fugue_b200.api`` = the ``fugue.api`` functions of the path) around a RECORDING stand-in
engine, on CPU: what each wrapper hands to the engine (argument meaning as in fugue/execution/api.py:181-1232) and
the return-type rule they all share (``_convert_df``: a Fugue DataFrame in or ``as_fugue`` -> a Fugue DataFrame out,
a native object in -> a native object out; ``as_local`` forces a local result)."""
from typing import Any, Dict, List

import pandas as pd
import pytest

from fugue_b200 import api as fa
from fugue_b200.column import SelectColumns, col, functions as ff
from fugue_b200.dataframe import ArrowDataFrame, DataFrame, as_fugue_df
from fugue_b200.lifecycle import EngineLifecycle
from fugue_b200.partition import PartitionSpec


class Recorder(EngineLifecycle):
    """Every engine method records (name, args, kwargs) and returns a fixed frame tagged with the call number."""

    is_distributed = False

    def __init__(self):
        self.conf: Dict[str, Any] = {}
        self.log: List[Any] = []

    def to_df(self, df, schema=None):
        return as_fugue_df(df, schema)

    def convert_yield_dataframe(self, df, as_local):
        self.log.append(("convert", as_local))
        return df

    def get_current_parallelism(self):
        return 7

    @property
    def sql_engine(self):
        return self

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)

        def method(*args, **kwargs):
            self.log.append((name, args, kwargs))
            return ArrowDataFrame([[len(self.log)]], "n:long")

        return method

    def calls(self, name):
        return [c for c in self.log if c[0] == name]


@pytest.fixture
def eng():
    return Recorder()


PDF = pd.DataFrame({"k": [1, 2], "v": [1.5, 2.5]})
PDF2 = pd.DataFrame({"k": [1, 2], "w": [3, 4]})
PDF3 = pd.DataFrame({"k": [1, 2], "z": [5, 6]})


def test_return_type_rule(eng):
    assert isinstance(fa.distinct(PDF, engine=eng), pd.DataFrame)
    assert isinstance(fa.distinct(ArrowDataFrame(PDF), engine=eng), DataFrame)
    assert isinstance(fa.distinct(PDF, engine=eng, as_fugue=True), DataFrame)
    assert isinstance(fa.distinct(PDF.pipe(lambda d: __import__("pyarrow").Table.from_pandas(d)), engine=eng), pd.DataFrame)
    fa.distinct(PDF, engine=eng, as_local=True)
    assert eng.calls("convert")[-1] == ("convert", True) and eng.calls("convert")[0] == ("convert", False)


def test_relational_wrappers_pass_their_arguments(eng):
    fa.aggregate(PDF, "k", engine=eng, s=ff.sum(col("v")), one=1)
    (_, (df, spec, cols), _), = eng.calls("aggregate")
    assert df.schema == "k:long,v:double" and spec == PartitionSpec(by=["k"])
    assert [str(c) for c in cols] == ["SUM(v) AS s", "1 AS one"]
    fa.aggregate(PDF, engine=eng, c=ff.count(col("*")))
    assert eng.calls("aggregate")[-1][1][1] is None                       # no keys: a global aggregate

    fa.select(PDF, "k", (col("v") * 2).alias("d"), where=col("v") > 1, having=None, distinct=True, engine=eng)
    (_, (df, cols), kw), = eng.calls("select")
    assert isinstance(cols, SelectColumns) and cols.is_distinct and [str(c) for c in cols.all_cols] == ["k", "*(v,2) AS d"]
    assert str(kw["where"]) == ">(v,1)" and kw["having"] is None

    fa.filter(PDF, col("k") == 1, engine=eng)
    assert str(eng.calls("filter")[0][1][1]) == "==(k,1)"
    fa.assign(PDF, engine=eng, x=1, y=col("v") + 1)
    assert [str(c) for c in eng.calls("assign")[0][1][1]] == ["1 AS x", "+(v,1) AS y"]


def test_joins_chain_left_to_right(eng):
    out = fa.join(PDF, PDF2, PDF3, how="inner", on=["k"], engine=eng)
    first, second = eng.calls("join")
    assert first[2] == dict(how="inner", on=["k"]) and first[1][0].schema == "k:long,v:double"
    assert second[1][0].schema == "n:long" and second[1][1].schema == "k:long,z:long"     # (a JOIN b) JOIN c
    assert isinstance(out, pd.DataFrame)
    assert isinstance(fa.join(PDF, ArrowDataFrame(PDF2), how="cross", engine=eng), DataFrame)  # any Fugue frame in
    for f, how in [(fa.inner_join, "inner"), (fa.semi_join, "semi"), (fa.anti_join, "anti"),
                   (fa.left_outer_join, "left_outer"), (fa.right_outer_join, "right_outer"),
                   (fa.full_outer_join, "full_outer"), (fa.cross_join, "cross")]:
        f(PDF, PDF2, engine=eng)
        assert eng.calls("join")[-1][2] == dict(how=how, on=None)


def test_set_and_row_operations(eng):
    fa.union(PDF, PDF, PDF, distinct=False, engine=eng)
    assert len(eng.calls("union")) == 2 and all(c[2] == dict(distinct=False) for c in eng.calls("union"))
    fa.subtract(PDF, PDF, engine=eng)
    fa.intersect(PDF, PDF, distinct=True, engine=eng)
    assert eng.calls("subtract")[0][2] == dict(distinct=True) and eng.calls("intersect")[0][2] == dict(distinct=True)
    fa.dropna(PDF, how="all", thresh=2, subset=["v"], engine=eng)
    assert eng.calls("dropna")[0][2] == dict(how="all", thresh=2, subset=["v"])
    fa.fillna(PDF, {"v": 0.0}, subset=None, engine=eng)
    assert eng.calls("fillna")[0][2] == dict(value={"v": 0.0}, subset=None)
    fa.sample(PDF, frac=0.5, replace=True, seed=3, engine=eng)
    assert eng.calls("sample")[0][2] == dict(n=None, frac=0.5, replace=True, seed=3)
    fa.take(PDF, 2, presort="v desc", na_position="first", partition=dict(by=["k"]), engine=eng)
    kw = eng.calls("take")[0][2]
    assert (kw["n"], kw["presort"], kw["na_position"]) == (2, "v desc", "first") and kw["partition_spec"] == PartitionSpec(by=["k"])
    fa.take(PDF, 1, engine=eng)
    assert eng.calls("take")[-1][2]["partition_spec"] is None


def test_engine_level_wrappers(eng):
    fa.repartition(PDF, dict(by=["k"], num=4), engine=eng)
    assert eng.calls("repartition")[0][1][1] == PartitionSpec(by=["k"], num=4)
    fa.persist(PDF, lazy=True, engine=eng, level="x")
    assert eng.calls("persist")[0][2] == dict(lazy=True, level="x")
    fa.broadcast(PDF, engine=eng)
    assert len(eng.calls("broadcast")) == 1
    fa.load("a.parquet", columns=["k"], engine=eng, opt=1)
    assert eng.calls("load_df")[0][1] == ("a.parquet",) and eng.calls("load_df")[0][2] == dict(format_hint=None, columns=["k"], opt=1)
    fa.save(PDF, "b.csv", mode="error", engine=eng, header=True)
    assert eng.calls("save_df")[0][1][1] == "b.csv" and eng.calls("save_df")[0][2] == dict(format_hint=None, mode="error", header=True)
    assert fa.get_current_parallelism(eng) == 7
    got = fa.run_engine_function(lambda e: e.distinct(e.to_df(PDF)), engine=eng, infer_by=[PDF])
    assert isinstance(got, pd.DataFrame) and fa.run_engine_function(lambda e: 5, engine=eng) == 5
    assert fa.as_fugue_engine_df(eng, PDF).schema == "k:long,v:double"


def test_raw_sql_pieces(eng):
    fa.raw_sql("SELECT k, SUM(v) AS s FROM", PDF, "GROUP BY k", engine=eng)
    (_, (dfs, st), _), = eng.calls("select")
    assert list(dfs) == ["_0"] and st.construct() == "SELECT k, SUM(v) AS s FROM _0 GROUP BY k"
    fa.raw_sql("SELECT * FROM", PDF, "INNER JOIN", PDF2, "ON a.k = b.k", engine=eng)
    dfs, st = eng.calls("select")[-1][1]
    assert list(dfs) == ["_0", "_1"] and st.construct({"_0": "a", "_1": "b"}) == "SELECT * FROM a INNER JOIN b ON a.k = b.k"
