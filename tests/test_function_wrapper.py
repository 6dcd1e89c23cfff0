"""Host logic of the function -> transformer adapter (``fugue_b200.api._FuncAsTransformer``), CPU only.

What the reference's adapter does for the north-star path (``_FuncAsTransformer.transform``
fugue/extensions/transformer/convert.py:343-348, ``DataFrameFunctionWrapper.run``
fugue/dataframe/function_wrapper.py:68-148 with the annotated-parameter classes :330-516): the first
parameter's annotation decides how a logical partition is handed to the user function, the return value is
turned back into a local dataframe of the declared schema, extra parameters come from ``params``, a
``PartitionCursor``-typed parameter receives the cursor, ``ignore_errors`` turns listed exceptions into an
empty partition.  The runner is exercised here on host dataframes - no engine, no GPU.
"""
from typing import Any, Dict, Iterable, Iterator, List

import pandas as pd
import pyarrow as pa
import pytest

from fugue_b200.api import _FuncAsTransformer
from fugue_b200.lifecycle import EngineLifecycle
from fugue_b200.dataframe import ArrayDataFrame, ArrowDataFrame
from fugue_b200.partition import PartitionCursor, PartitionSpec
from fugue_b200.schema import Schema

ROWS = [[0, 1.5], [0, 2.5], [1, 4.0]]
SCHEMA = "k:long,v:double"


def _run(func: Any, schema: Any = "*", params: Any = None, ignore: Any = None, rows: Any = ROWS) -> Any:
    df = ArrowDataFrame(ArrayDataFrame(rows, SCHEMA).as_arrow(), SCHEMA)
    tf = _FuncAsTransformer(func, schema, params)
    out_schema = tf.get_output_schema(df)
    spec = PartitionSpec(by=["k"])
    cursor = spec.get_cursor(df.schema, 0)
    cursor.set(lambda: df.peek_array(), 3, 0)
    res = tf.make_runner(out_schema, ignore or [])(cursor, df)
    assert res.schema == out_schema
    return tf, res.as_array(type_safe=True)


def test_pandas_in_out():
    def f(df: pd.DataFrame) -> pd.DataFrame:
        return df.assign(v=df.v * 2)

    tf, out = _run(f)
    assert tf.get_format_hint() == "pandas" and out == [[0, 3.0], [0, 5.0], [1, 8.0]]


def test_arrow_in_out():
    def f(df: pa.Table) -> pa.Table:
        return df.slice(1)

    tf, out = _run(f)
    assert tf.get_format_hint() == "pyarrow" and out == ROWS[1:]


def test_rows_and_dicts():
    def rows(df: List[List[Any]]) -> Iterable[List[Any]]:
        for r in df:
            yield [r[0], r[1] + 1]

    def lazy_rows(df: Iterable[List[Any]]) -> List[List[Any]]:
        return [[r[0], -r[1]] for r in df]

    def dicts(df: List[Dict[str, Any]]) -> Iterable[Dict[str, Any]]:
        for r in df:
            yield dict(v=r["v"], k=r["k"] + 10)  # key order of the dict does not matter, the schema does

    tf, out = _run(rows)
    assert tf.get_format_hint() is None and out == [[0, 2.5], [0, 3.5], [1, 5.0]]
    assert _run(lazy_rows)[1] == [[0, -1.5], [0, -2.5], [1, -4.0]]
    assert _run(dicts)[1] == [[10, 1.5], [10, 2.5], [11, 4.0]]


def test_iterables_of_frames():
    def chunks(df: Iterable[pd.DataFrame]) -> Iterable[pd.DataFrame]:
        for part in df:
            yield part.iloc[:1]
            yield part.iloc[:0]  # empty chunks are dropped
            yield part.iloc[1:]

    def tables(df: Iterator[pa.Table]) -> Iterable[pa.Table]:
        for part in df:
            yield part.slice(2)
            yield part.slice(0, 2)

    def nothing(df: Iterable[pd.DataFrame]) -> Iterable[pd.DataFrame]:
        for part in df:
            yield part.iloc[:0]

    assert _run(chunks)[1] == ROWS
    assert _run(tables)[1] == [ROWS[2], ROWS[0], ROWS[1]]
    assert _run(nothing)[1] == []


def test_schema_sources():
    # schema: *,w:double
    def commented(df: pd.DataFrame) -> pd.DataFrame:
        return df.assign(w=df.v + 1)

    def bare(df: pd.DataFrame) -> pd.DataFrame:
        return df

    # schema: k:long
    # a remark in between
    #schema:*,w:double # the lowest hint wins, the remark after the second '#' is dropped
    # another remark
    def lowest(df: pd.DataFrame) -> pd.DataFrame:
        return df.assign(w=df.v + 1)

    # schema:
    def empty_hint(df: pd.DataFrame) -> pd.DataFrame:
        return df

    # schema : *, w : double # trailing remark
    # # # schema : k:long    (commented out twice: not a hint)
    def nested_hash(df: pd.DataFrame) -> pd.DataFrame:
        return df.assign(w=df.v + 1)

    assert _run(commented, schema=None)[1] == [[0, 1.5, 2.5], [0, 2.5, 3.5], [1, 4.0, 5.0]]
    assert _run(lowest, schema=None)[1] == [[0, 1.5, 2.5], [0, 2.5, 3.5], [1, 4.0, 5.0]]
    with pytest.raises(SyntaxError):
        _FuncAsTransformer(empty_hint, None, None)
    assert _run(nested_hash, schema=None)[1][0] == [0, 1.5, 2.5]
    assert _run(commented, schema=Schema("k:long,v:double,w:double"))[1][2] == [1, 4.0, 5.0]
    with pytest.raises(ValueError):
        _FuncAsTransformer(bare, None, None)         # no schema argument, no '# schema:' comment
    with pytest.raises(TypeError):
        _FuncAsTransformer(lambda df: df, "*", None)  # un-annotated dataframe parameter
    with pytest.raises(TypeError):
        _FuncAsTransformer(123, "*", None)


def test_params_cursor_and_ignored_errors():
    def f(df: pd.DataFrame, c: PartitionCursor, add: float, scale: float = 1.0) -> pd.DataFrame:
        assert c.partition_no == 3 and c.key_value_array == [0] and c.row == [0, 1.5]
        return df.assign(v=(df.v + add) * scale)

    assert _run(f, params=dict(add=0.5, scale=2.0))[1] == [[0, 4.0], [0, 6.0], [1, 9.0]]

    def boom(df: pd.DataFrame) -> pd.DataFrame:
        raise NotImplementedError("nope")

    assert _run(boom, ignore=[NotImplementedError])[1] == []   # processors.py:330-338
    with pytest.raises(NotImplementedError):
        _run(boom, ignore=[ValueError])
    with pytest.raises(NotImplementedError):
        _run(boom)


def test_one_row_as_a_dict():
    def first(rows: list[dict[str, Any]]) -> dict[str, Any]:      # built-in generics; one row out
        return dict(v=rows[0]["v"], k=rows[0]["k"])

    tf, out = _run(first)
    assert out == [[0, 1.5]] and tf.get_format_hint() is None


def test_none_and_empty_outputs():
    def none(df: pd.DataFrame) -> None:
        return None

    def empty_rows(df: List[List[Any]]) -> List[List[Any]]:
        return []

    assert _run(none)[1] == [] and _run(empty_rows)[1] == []


def test_format_hints():
    """Input annotation first, then the return annotation (tests/fugue/dataframe/test_function_wrapper.py:93-107)."""
    from fugue_b200.table import B200Table

    def arrow_in(df: pa.Table) -> None: ...
    def pandas_in(df: pd.DataFrame) -> pa.Table: ...
    def frames_in(df: Iterable[pd.DataFrame]) -> pa.Table: ...
    def tables_in(df: Iterator[pa.Table]) -> List[List[Any]]: ...
    def frames_out(df: List[List[Any]]) -> Iterator[pd.DataFrame]: ...
    def table_out(df: Iterable[Dict[str, Any]]) -> pa.Table: ...
    def rows(df: List[List[Any]]) -> List[List[Any]]: ...
    def device_in(df: B200Table) -> B200Table: ...
    def device_out(df: pd.DataFrame) -> B200Table: ...
    def device_out_rows(df: List[List[Any]]) -> B200Table: ...

    hint = lambda f: _FuncAsTransformer(f, "*", None).get_format_hint()  # noqa: E731
    assert [hint(f) for f in (arrow_in, pandas_in, frames_in, tables_in, frames_out, table_out, rows)] == \
        ["pyarrow", "pandas", "pandas", "pyarrow", "pandas", "pyarrow", None]
    assert hint(device_in) == "b200" and hint(device_out) == "pandas" and hint(device_out_rows) is None


class _HostEngine(EngineLifecycle):
    """A stand-in engine for ``fa.out_transform`` on CPU: its map engine calls the runner once per key group of
    a pandas frame - the part of ``map_dataframe`` that is host logic in every engine."""

    def __init__(self):
        self.conf: Dict[str, Any] = {}
        self.calls: List[Any] = []

    is_distributed = False

    @property
    def map_engine(self):
        return self

    def persist(self, df, lazy=False, **kwargs):
        self.calls.append("persist")
        return df

    def convert_yield_dataframe(self, df, as_local):
        return df.as_local() if as_local else df

    def save_df(self, df, path, format_hint=None, mode="overwrite", **kwargs):
        from fugue_b200 import io as IO

        IO.save_df(self.to_df(df), path, format_hint, mode, **kwargs)

    def load_df(self, path, format_hint=None, columns=None, **kwargs):
        from fugue_b200 import io as IO

        return IO.load_df(path, format_hint, columns, **kwargs)

    def to_df(self, df, schema=None):
        from fugue_b200.dataframe import as_fugue_df

        return as_fugue_df(df, schema)

    def map_dataframe(self, df, map_func, output_schema, partition_spec, on_init=None, map_func_format_hint=None):
        pdf = df.as_pandas()
        groups = [pdf] if not partition_spec.partition_by else \
            [g for _, g in pdf.groupby(partition_spec.partition_by, sort=True)]
        outs: List[Any] = []
        for no, g in enumerate(groups):
            part = ArrowDataFrame(g.reset_index(drop=True), df.schema)
            cursor = partition_spec.get_cursor(df.schema, 0)
            cursor.set(lambda: part.peek_array(), no, 0)
            out = map_func(cursor, part)
            assert out.schema == output_schema
            outs.append(out.as_pandas())
            self.calls.append(map_func_format_hint)
        return ArrowDataFrame(pd.concat(outs, ignore_index=True) if outs else None, output_schema)


def test_out_transform_runs_for_side_effects_only():
    from fugue_b200 import api as fa

    seen = []

    def collect(df: pd.DataFrame, tag: str) -> None:
        seen.append((tag, df.k.iloc[0], len(df)))

    def lazy(df: List[List[Any]]) -> Iterable[List[Any]]:     # a generator has to be drained to have its effects
        for r in df:
            seen.append(("row", r[0], r[1]))
            yield r

    def boom(df: pd.DataFrame) -> None:
        raise KeyError("x")

    eng = _HostEngine()
    pdf = pd.DataFrame(ROWS, columns=["k", "v"])
    assert fa.out_transform(pdf, collect, params=dict(tag="t"), partition=dict(by=["k"]), engine=eng) is None
    assert seen == [("t", 0, 2), ("t", 1, 1)] and eng.calls == ["pandas", "pandas"]
    seen.clear()
    fa.out_transform(pdf, lazy, engine=eng)
    assert seen == [("row", 0, 1.5), ("row", 0, 2.5), ("row", 1, 4.0)]
    fa.out_transform(pdf, boom, ignore_errors=[KeyError], engine=eng)      # swallowed per partition
    with pytest.raises(KeyError):
        fa.out_transform(pdf, boom, engine=eng)
    with pytest.raises(ValueError):
        fa.out_transform("x.csv", collect, engine=eng)
    with pytest.raises(NotImplementedError):
        fa.out_transform(pdf, collect, callback=print, engine=eng)


def test_transform_control_flow_on_a_host_engine(tmp_path):
    """``fa.transform`` around the map engine (fugue/workflow/api.py:34-184): schema resolution, params, the
    return-type rule, persist, save_path / checkpoint, a parquet path as input - with the host stand-in engine."""
    from fugue_b200 import api as fa
    from fugue_b200.dataframe import DataFrame

    def add(df: pd.DataFrame, n: float = 1.0) -> pd.DataFrame:
        return df.assign(w=df.v + n)

    eng = _HostEngine()
    pdf = pd.DataFrame(ROWS, columns=["k", "v"])
    want = [[0, 1.5, 3.5], [0, 2.5, 4.5], [1, 4.0, 6.0]]
    out = fa.transform(pdf, add, schema="*,w:double", params=dict(n=2.0), partition=dict(by=["k"]), engine=eng)
    assert isinstance(out, pd.DataFrame) and out.values.tolist() == want         # native in -> native out
    fdf = fa.transform(ArrowDataFrame(pdf), add, schema="*,w:double", params=dict(n=2.0), engine=eng)
    assert isinstance(fdf, DataFrame) and fdf.as_array() == want                  # Fugue frame in -> Fugue frame out
    assert isinstance(fa.transform(pdf, add, schema="*,w:double", engine=eng, as_fugue=True), DataFrame)
    eng.calls.clear()
    fa.transform(pdf, add, schema="*,w:double", engine=eng, persist=True)
    assert "persist" in eng.calls
    # the declared schema is enforced on what the function returns (PandasDataFrame(output, schema) in the
    # reference's adapter, function_wrapper.py:442-460): a double column declared long is cast
    assert fa.transform(pdf, add, schema="*,w:long", engine=eng).values.tolist() == [[0, 1.5, 2], [0, 2.5, 3], [1, 4.0, 5]]
    with pytest.raises(Exception):
        fa.transform(pdf, add, schema="*,missing:long", engine=eng)               # a declared column is not produced

    p1 = str(tmp_path / "out" / "a.parquet")
    assert fa.transform(pdf, add, schema="*,w:double", engine=eng, save_path=p1) == p1
    assert pd.read_parquet(p1).values.tolist() == [[0, 1.5, 2.5], [0, 2.5, 3.5], [1, 4.0, 5.0]]
    def twice(df: pd.DataFrame) -> pd.DataFrame:
        return df.assign(ww=df.w * 2)

    again = fa.transform(p1, twice, schema="*,ww:double", engine=eng)              # a parquet path as input
    assert list(again.columns) == ["k", "v", "w", "ww"] and len(again) == 3
    ck = fa.transform(pdf, add, schema="*,w:double", engine=eng, save_path=str(tmp_path / "b.parquet"),
                      checkpoint=True)
    assert ck.values.tolist() == [[0, 1.5, 2.5], [0, 2.5, 3.5], [1, 4.0, 5.0]]      # continued from the file
    for bad in (dict(save_path=str(tmp_path / "x.csv")), dict(checkpoint=True)):   # csv path; no checkpoint dir
        with pytest.raises(ValueError):
            fa.transform(pdf, add, schema="*,w:double", engine=eng, **bad)
    eng.conf["fugue.workflow.checkpoint.path"] = str(tmp_path / "ckpt")
    assert len(fa.transform(pdf, add, schema="*,w:double", engine=eng, checkpoint=True)) == 3
    assert len(list((tmp_path / "ckpt").iterdir())) == 1
    with pytest.raises(ValueError):
        fa.transform("in.csv", add, schema="*,w:double", engine=eng)


def test_chunked_frames_are_matched_by_column_name():
    """fugue_test/builtin_suite.py:426-488 (the mapInPandas shape): chunks may come back with their columns in any
    order, or not at all; a narrower Arrow type is widened to the declared one."""
    from fugue_b200 import api as fa

    # schema: *,c:int
    def mt_pandas(dfs: Iterable[pd.DataFrame], empty: bool = False) -> Iterator[pd.DataFrame]:
        for df in dfs:
            if not empty:
                df = df.assign(c=2)
                yield df[list(reversed(list(df.columns)))]

    # schema: *
    def mt_arrow(dfs: Iterable[pa.Table], empty: bool = False) -> Iterator[pa.Table]:
        for df in dfs:
            if not empty:
                yield df.select(list(reversed(df.schema.names)))

    # schema: a:long
    def mt_arrow_2(dfs: Iterable[pa.Table]) -> Iterator[pa.Table]:
        for df in dfs:
            yield df.drop(["b"])

    eng = _HostEngine()
    a = ArrowDataFrame([[1, 2], [3, 4]], "a:int,b:int")
    got = fa.transform(a, mt_pandas, engine=eng)
    assert got.schema == "a:int,b:int,c:int" and got.as_array() == [[1, 2, 2], [3, 4, 2]]
    assert fa.transform(a, mt_arrow, engine=eng).as_array() == [[1, 2], [3, 4]]
    narrow = fa.transform(a, mt_arrow_2, engine=eng)
    assert narrow.schema == "a:long" and narrow.as_array() == [[1], [3]]
    for f, schema in ((mt_pandas, "a:int,b:int,c:int"), (mt_arrow, "a:int,b:int")):
        for part in (None, dict(by=["a"])):
            out = fa.transform(a, f, params=dict(empty=True), partition=part, engine=eng)
            assert out.schema == schema and out.count() == 0


def test_validation_rules_from_comment_hints():
    """fugue/extensions/_utils.py:36-150 and the uses in fugue_test/builtin_suite.py:638-730, 1403-1460: rules next
    to the schema hint are checked before the function runs - partitioning rules without data, input rules
    against the input's schema."""
    from fugue_b200 import api as fa

    # schema: *
    # partitionby_has: k
    # presort_is: v desc
    # input_has: k, v:double
    def strict(df: pd.DataFrame) -> pd.DataFrame:
        return df

    # partitionby_is: a, b
    # presort_has: c
    # input_is: k:long,v:double
    def other(df: pd.DataFrame) -> None:
        pass

    # partitionby_has:
    def empty_rule(df: pd.DataFrame) -> None:
        pass

    tf = _FuncAsTransformer(strict, None, None)
    assert tf._rules == dict(partitionby_has=["k"], presort_is=[("v", False)], input_has=["k", "v:double"])
    assert _FuncAsTransformer(other, "*", None)._rules == dict(
        partitionby_is=["a", "b"], presort_has=[("c", True)], input_is="k:long,v:double")
    with pytest.raises(SyntaxError):
        _FuncAsTransformer(empty_rule, "*", None)

    eng = _HostEngine()
    pdf = pd.DataFrame(ROWS, columns=["k", "v"])
    good = dict(by=["k"], presort="v desc")
    assert len(fa.transform(pdf, strict, partition=good, engine=eng)) == 3
    for bad in (None, dict(by=["v"]), dict(by=["k"]), dict(by=["k"], presort="v"), dict(by=["k"], presort="v desc, x")):
        with pytest.raises(fa.FugueWorkflowCompileValidationError):
            fa.transform(pdf, strict, partition=bad, engine=eng)
    for frame in (pdf.rename(columns={"v": "w"}), pdf.assign(v=[1, 2, 3])):      # v missing; v is not a double
        with pytest.raises(fa.FugueWorkflowRuntimeValidationError):
            fa.transform(frame, strict, partition=good, engine=eng)

    ok = dict(by=["a", "b"], presort="c")
    wide = pd.DataFrame({"k": [1], "v": [1.0]})
    with pytest.raises(fa.FugueWorkflowCompileValidationError):
        fa.out_transform(wide, other, partition=dict(by=["a"], presort="c"), engine=eng)   # partitionby_is: exact set
    with pytest.raises(fa.FugueWorkflowRuntimeValidationError):
        fa.out_transform(wide.assign(extra=1), other, partition=ok, engine=eng)             # input_is: exact schema
