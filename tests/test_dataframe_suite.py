"""Conformance of the host frames (ArrayDataFrame / ArrowDataFrame / PandasDataFrame) with the
behaviours the reference's dataframe suite pins (fugue_test/dataframe_suite.py:25-462 - every new DataFrame
type has to pass it).  The expected literals are the suite's; the harness is a table of cases run against every
constructor.  CPU only: these frames are what crosses the boundary into and out of the device engine
(``to_df`` ingest, ``as_local`` egress), the device frame's own conformance runs in the GPU tests.
"""
from datetime import date, datetime
from typing import Any, Callable, List

import numpy as np
import pandas as pd
import pytest

from fugue_b200 import api as fa
from fugue_b200.dataframe import (ArrayDataFrame, ArrowDataFrame, DataFrame, FugueDataFrameOperationError,
                                  FugueDatasetEmptyError, PandasDataFrame, df_eq)

TS = pd.Timestamp("2020-01-01")


@pytest.fixture(params=[ArrayDataFrame, ArrowDataFrame, PandasDataFrame], ids=lambda c: c.__name__)
def mk(request: Any) -> Callable[[Any, Any], DataFrame]:
    return lambda data, schema: request.param(data, schema)


def _both_readers() -> List[Callable[..., List[List[Any]]]]:
    return [lambda df, *a: fa.as_array(df, *a, type_safe=True),
            lambda df, *a: list(fa.as_array_iterable(df, *a, type_safe=True))]


def test_native_and_fugue_views(mk):
    df = mk([[1]], "a:int")
    assert fa.is_df(df) and fa.as_fugue_df(df) is df
    native = fa.get_native_as_df(df)
    assert fa.is_df(native) and not isinstance(native, DataFrame) and fa.get_native_as_df(native) is native


def test_peek_and_emptiness(mk):
    empty = mk([], "x:str,y:double")
    for peek in (fa.peek_array, fa.peek_dict):
        with pytest.raises(FugueDatasetEmptyError):
            peek(empty)
    df = mk([["a", 1.0], ["b", 2.0]], "x:str,y:double")
    assert fa.count(df) == 2 and not fa.is_empty(df) and fa.is_empty(empty)
    assert fa.peek_array(df) == ["a", 1.0] and fa.peek_dict(df) == dict(x="a", y=1.0)
    assert fa.as_pandas(df).values.tolist() == [["a", 1.0], ["b", 2.0]]
    assert fa.as_pandas(empty).values.tolist() == [] and fa.is_local(fa.as_pandas(empty))


def test_as_local_keeps_metadata(mk):
    for f in (fa.as_local, fa.as_local_bounded):
        with pytest.raises(NotImplementedError):
            f(10)
    df = mk([["a", 1.0]], "x:str,y:double")
    df.reset_metadata({"a": 1})
    for f in (fa.as_local, fa.as_local_bounded):
        ldf = f(df)
        assert fa.is_local(ldf) and fa.is_bounded(ldf) and ldf.metadata == {"a": 1}


@pytest.mark.parametrize("data", [[], [["a", 1]]])
def test_drop_and_select_columns(mk, data):
    dropped = fa.drop_columns(mk(data, "a:str,b:int"), ["a"])
    selected = fa.select_columns(mk(data, "a:str,b:int"), ["b"])
    for df in (dropped, selected):
        assert fa.get_schema(df) == "b:int" and fa.get_column_names(df) == ["b"]
        assert fa.as_array(df, type_safe=True) == [r[1:] for r in data]
        with pytest.raises(FugueDataFrameOperationError):
            fa.drop_columns(df, ["b"])       # a frame can't lose its last column
        with pytest.raises(FugueDataFrameOperationError):
            fa.drop_columns(df, ["x"])       # unknown column
        with pytest.raises(FugueDataFrameOperationError):
            fa.select_columns(df, [])
        with pytest.raises(FugueDataFrameOperationError):
            fa.select_columns(df, ["a"])
    assert df_eq(fa.select_columns(mk([["a", 1, 2]], "a:str,b:int,c:int"), ["c", "a"]), [[2, "a"]], "c:int,a:str",
                 throw=True)


@pytest.mark.parametrize("data", [[], [["a", 1]]])
def test_rename(mk, data):
    df = mk(data, "a:str,b:int")
    assert df_eq(fa.rename(df, columns=dict(a="aa")), data, "aa:str,b:int", throw=True)
    assert fa.get_schema(df) == "a:str,b:int"                          # the input is a value, not mutated
    assert df_eq(fa.rename(df, columns={}), data, "a:str,b:int", throw=True)
    with pytest.raises(FugueDataFrameOperationError):
        fa.rename(df, columns=dict(aa="ab"))


def test_rows_come_back_as_plain_python(mk):
    for read in _both_readers():
        assert read(mk([], "a:str,b:int")) == []
        assert read(mk([["a", 1]], "a:str,b:int")) == [["a", 1]]
        assert read(mk([["a", 1]], "a:str,b:int"), ["a", "b"]) == [["a", 1]]
        assert read(mk([["a", 1]], "a:str,b:int"), ["b", "a"]) == [[1, "a"]]
        for one in (1.0, np.float64(1.0)):
            (row,) = read(mk([[one, 1]], "a:double,b:int"))
            assert row == [1.0, 1] and type(row[0]) is float and type(row[1]) is int
        (row,) = read(mk([[TS, 1]], "a:datetime,b:int"))
        assert row == [datetime(2020, 1, 1), 1] and type(row[0]) is datetime
        assert read(mk([[pd.NaT, 1]], "a:datetime,b:int")) == [[None, 1]]          # missing markers are NULL
        assert read(mk([[float("nan"), 1]], "a:double,b:int")) == [[None, 1]]
        assert read(mk([[float("inf"), 1]], "a:double,b:int")) == [[float("inf"), 1]]


def test_dicts(mk):
    for read in (fa.as_dicts, lambda df, *a: list(fa.as_dict_iterable(df, *a))):
        assert read(mk([[pd.NaT, 1]], "a:datetime,b:int")) == [dict(a=None, b=1)]
        assert read(mk([[pd.NaT, 1]], "a:datetime,b:int"), ["b"]) == [dict(b=1)]
        assert read(mk([[TS, 1]], "a:datetime,b:int")) == [dict(a=datetime(2020, 1, 1), b=1)]
        assert read(mk([[TS, 1]], "a:datetime,b:int"), ["b"]) == [dict(b=1)]


@pytest.mark.parametrize("schema,data,expected", [
    ("a:[int]", [[[30, 40]]], None),
    ("x:{a:int}", [[{"a": 1}], [{"a": 2}]], None),
    ("x:<str,int>", [[[("a", 1), ("b", 3)]], [[("b", 2)]]], None),
    ("a:{a:str,b:[int]}", [[dict(a="1", b=[3, 4], d=1.0)], [dict(b=[30, 40])]],
     [[dict(a="1", b=[3, 4])], [dict(a=None, b=[30, 40])]]),       # unknown keys dropped, missing keys NULL
    ("a:[{a:str,b:[int]}]", [[[dict(b=[30, 40])]]], [[[dict(a=None, b=[30, 40])]]]),
    ("a:bytes", [[b"\x01\x05"]], None),
    ("a:[float]", [[[float("nan"), 2.0]]], [[[None, 2.0]]]),
    ("a:{b:bool}", [[dict(b=True)]], None),
    ("a:[{b:[long]}]", [[[dict(b=[30, 40])]]], None),
])
def test_nested_and_binary_types(mk, schema, data, expected):
    df = mk(data, schema)
    assert fa.get_schema(df) == schema
    expected = data if expected is None else expected
    assert fa.as_array(df, type_safe=True) == expected
    assert ArrowDataFrame(fa.as_arrow(df)).as_array() == expected


def test_as_arrow(mk):
    back = lambda df: list(ArrowDataFrame(fa.as_arrow(df)).as_dict_iterable())  # noqa: E731
    assert back(mk([], "a:int,b:int")) == [] and fa.is_local(fa.as_arrow(mk([], "a:int,b:int")))
    assert back(mk([[pd.NaT, 1]], "a:datetime,b:int")) == [dict(a=None, b=1)]
    assert back(mk([[TS, 1]], "a:datetime,b:int")) == [dict(a=datetime(2020, 1, 1), b=1)]


def test_head(mk):
    empty = mk([], "a:str,b:int")
    assert fa.as_array(fa.head(empty, 1)) == [] and fa.as_array(fa.head(empty, 1, ["b"])) == []
    one = mk([["a", 1]], "a:str,b:int")
    assert fa.as_array(fa.head(one, 1)) == [["a", 1]] and fa.as_array(fa.head(one, 1, ["b", "a"])) == [[1, "a"]]
    assert fa.as_array(fa.head(one, 0)) == []
    four = mk([[0, 1], [0, 2], [1, 1], [1, 3]], "a:int,b:int")
    assert fa.count(fa.head(four, 2)) == 2 and fa.count(fa.head(four, 10)) == 4
    assert fa.is_local(fa.head(four, 10)) and fa.is_bounded(fa.head(four, 10))


def test_show(mk, capsys):
    df = mk([["a", 1]], "a:str,b:int")
    df.reset_metadata({"k": "v"})
    fa.show(df, with_count=True, title="T")
    out = capsys.readouterr().out
    assert "T" in out and "a:str,b:int" in out and "['a', 1]" in out and "Total count: 1" in out and "'k': 'v'" in out
    fa.show(mk([], "a:str,b:int"))
    assert "(empty)" in capsys.readouterr().out


# (input schema, rows, columns to alter, schema after, rows after [alternatives allowed])
ALTER_CASES = [
    ("a:str,b:int", [], "a:str,b:str", "a:str,b:str", [[]]),
    ("a:str,b:int", [["a", 1], ["c", None]], "b:int,a:str", "a:str,b:int", [[["a", 1], ["c", None]]]),
    ("a:str,b:bool", [["a", True], ["b", False], ["c", None]], "b:str", "a:str,b:str",
     [[["a", "True"], ["b", "False"], ["c", None]], [["a", "true"], ["b", "false"], ["c", None]]]),
    ("a:str,b:int", [["a", 1], ["c", None]], "b:str", "a:str,b:str",
     [[["a", "1"], ["c", None]], [["a", "1.0"], ["c", None]]]),
    ("a:str,b:int", [["a", 1], ["c", None]], "b:double", "a:str,b:double", [[["a", 1], ["c", None]]]),
    ("a:str,b:double", [["a", 1.1], ["b", None]], "b:str", "a:str,b:str", [[["a", "1.1"], ["b", None]]]),
    ("a:str,b:double", [["a", 1.0], ["b", None]], "b:int", "a:str,b:int", [[["a", 1], ["b", None]]]),
    ("a:str,b:date", [["a", date(2020, 1, 1)], ["b", date(2020, 1, 2)], ["c", None]], "b:str", "a:str,b:str",
     [[["a", "2020-01-01"], ["b", "2020-01-02"], ["c", None]]]),
    ("a:str,b:datetime", [["a", datetime(2020, 1, 1, 3, 4, 5)], ["b", datetime(2020, 1, 2, 16, 7, 8)], ["c", None]],
     "b:str", "a:str,b:str", [[["a", "2020-01-01 03:04:05"], ["b", "2020-01-02 16:07:08"], ["c", None]]]),
    ("a:str,b:str", [["a", "trUe"], ["b", "False"], ["c", None]], "b:bool,a:str", "a:str,b:bool",
     [[["a", True], ["b", False], ["c", None]]]),
    ("a:str,b:str", [["a", "1"]], "b:int,a:str", "a:str,b:int", [[["a", 1]]]),
    ("a:str,b:str", [["a", "1.1"], ["b", "2"], ["c", None]], "b:double", "a:str,b:double",
     [[["a", 1.1], ["b", 2.0], ["c", None]]]),
    ("a:str,b:str", [["1", "2020-01-01"], ["2", "2020-01-02"], ["3", None]], "b:date,a:int", "a:int,b:date",
     [[[1, date(2020, 1, 1)], [2, date(2020, 1, 2)], [3, None]]]),
    ("a:str,b:str", [["1", "2020-01-01 01:02:03"], ["2", "2020-01-02 01:02:03"], ["3", None]], "b:datetime,a:int",
     "a:int,b:datetime", [[[1, datetime(2020, 1, 1, 1, 2, 3)], [2, datetime(2020, 1, 2, 1, 2, 3)], [3, None]]]),
]


@pytest.mark.parametrize("schema,rows,alter,after,accepted", ALTER_CASES, ids=[f"{c[0]}->{c[2]}" for c in ALTER_CASES])
def test_alter_columns(mk, schema, rows, alter, after, accepted):
    out = fa.alter_columns(mk(rows, schema), alter, as_fugue=True)
    assert fa.get_schema(out) == after                                  # column ORDER of the input is kept
    assert fa.as_array(out, type_safe=True) in accepted


def test_alter_columns_rejects_what_cannot_be_cast(mk):
    with pytest.raises(Exception):
        fa.as_array(fa.alter_columns(mk([["1", "x"], ["2", "y"], ["3", None]], "a:str,b:str"), "b:int"))
    with pytest.raises(FugueDataFrameOperationError):
        fa.alter_columns(mk([["1", "x"]], "a:str,b:str"), "c:int")      # not a column of the frame
    sub_second = mk([[datetime(2020, 1, 1, 3, 4, 5, 250000)]], "t:datetime")
    assert fa.as_array(fa.alter_columns(sub_second, "t:str")) == [["2020-01-01 03:04:05.250000"]]


def test_native_frames_keep_any_column_names():
    pdf = pd.DataFrame([[0, 1, 2]], columns=["0", "1", "2"])
    assert fa.get_column_names(pdf) == ["0", "1", "2"]
    assert fa.get_column_names(fa.rename(pdf, {"0": "_0", "1": "_1", "2": "_2"})) == ["_0", "_1", "_2"]
    named = pd.DataFrame([[0, 1, 2]], columns=["a", "b", "c"])
    assert fa.get_column_names(fa.rename(named, {})) == ["a", "b", "c"]


def test_row_values_are_converted_to_the_schema(mk):
    """Rows are untyped Python data (tests/fugue/dataframe/test_array_dataframe.py:25-60, 86-140): text is parsed,
    numbers become text, floats are truncated into integer columns, nested values may be JSON text."""
    import json

    data = [["a", 1], ["b", 2]]
    assert mk(data, "a:str,b:str").as_array(type_safe=True) == [["a", "1"], ["b", "2"]]
    assert mk(data, "a:str,b:double").as_array(type_safe=True) == [["a", 1.0], ["b", 2.0]]
    mixed = mk([["a", 1], ["b", "2"]], "x:str,y:double")
    assert mixed.count() == 2 and mixed.peek_array() == ["a", 1.0] and mixed.peek_dict() == dict(x="a", y=1.0)
    (row,) = mk([[1.0, 1.1]], "a:double,b:int").as_array(type_safe=True)
    assert row == [1.0, 1] and type(row[1]) is int
    assert mk([[1.0, 1.1]], "a:double,b:int").as_array(["b", "a"], type_safe=True) == [[1, 1.0]]
    assert mk([["2020-01-01", 1.1]], "a:datetime,b:int").as_array(type_safe=True) == [[datetime(2020, 1, 1), 1]]
    assert mk([["TRUE", "no"], [1, 0]], "p:bool,q:bool").as_array() == [[True, False], [True, False]]
    with pytest.raises(Exception):
        mk([["maybe"]], "p:bool")
    with pytest.raises(Exception):
        mk([["x1"]], "p:int")
    nested = [[dict(a=1, b=[3, 4], d=1.0)], [json.dumps(dict(b=[30, "40"]))]]
    assert mk(nested, "a:{a:str,b:[int]}").as_array(type_safe=True) == \
        [[dict(a="1", b=[3, 4])], [dict(a=None, b=[30, 40])]]
    assert mk([[[json.dumps(dict(b=[30, "40"]))]]], "a:[{a:str,b:[int]}]").as_array(type_safe=True) == \
        [[[dict(a=None, b=[30, 40])]]]


def test_frames_built_from_frames(mk):
    """A frame as the data argument: same schema, a cast, a reorder, or a projection by a list of names
    (tests/fugue/dataframe/test_array_dataframe.py:36-58).  One divergence: the reference's ArrayDataFrame keeps
    the rows as given, so its int 1 under ``b:double`` later prints as "1"; this frame is typed, 1.0 -> "1.0"."""
    src = mk([["a", 1], ["b", 2]], "a:str,b:double")
    assert mk(src, None).as_array(type_safe=True) == [["a", 1.0], ["b", 2.0]]
    assert mk(src, "a:str,b:float64").schema == "a:str,b:double"
    assert mk(src, "b:str,a:str").as_array(type_safe=True) == [["1.0", "a"], ["2.0", "b"]]
    only_b = mk(src, ["b"])
    assert only_b.schema == "b:double" and only_b.as_array(type_safe=True) == [[1.0], [2.0]]
    assert mk(src, ["b:str"]).as_array(type_safe=True) == [["1.0"], ["2.0"]]
    with pytest.raises(Exception):
        mk(src, ["nope"])
    with pytest.raises(Exception):
        mk(123, None)
    assert mk([], "x:str,y:double").empty and mk(None, "x:str,y:double").empty


def test_pandas_input_with_a_schema():
    """tests/fugue/dataframe/test_pandas_dataframe.py:45-93 (what the reference says about object identity of
    the wrapped pandas frame does not apply: this frame is Arrow-backed)."""
    pdf = pd.DataFrame([["a", 1], ["b", 2]], columns=["a", "b"])
    assert PandasDataFrame(schema="a:str,b:int").count() == 0
    assert PandasDataFrame(pdf, "a:str,b:str").as_array() == [["a", "1"], ["b", "2"]]
    assert PandasDataFrame(pdf, "a:str,b:int").as_array() == [["a", 1], ["b", 2]]
    assert PandasDataFrame(pdf, "a:str,b:double").as_array() == [["a", 1.0], ["b", 2.0]]
    assert PandasDataFrame(pdf["b"], "b:str").as_array() == [["1"], ["2"]]          # a Series is a one-column frame
    assert PandasDataFrame(pdf["b"], "b:double").as_array() == [[1.0], [2.0]]
    xy = pd.DataFrame([["a", 1], ["b", 2]], columns=["x", "y"])
    assert PandasDataFrame(xy).schema == "x:str,y:long"
    reordered = PandasDataFrame(xy, "y:str,x:str")
    assert reordered.as_array() == [["1", "a"], ["2", "b"]] and PandasDataFrame(reordered).as_array() == reordered.as_array()
    assert PandasDataFrame([["a", "1"], ["b", "2"]], "x:str,y:double").peek_array() == ["a", 1.0]
    with pytest.raises(Exception):
        PandasDataFrame(123)


def test_misc_frame_protocol(mk):
    """tests/fugue/dataframe/test_dataframe.py:12-73."""
    import copy
    import json

    for f in (fa.as_fugue_df, fa.get_native_as_df):
        with pytest.raises(NotImplementedError):
            f(10)
    df = mk([["a", 1], ["b", 2]], "a:str,b:str")
    assert copy.copy(df) is df and copy.deepcopy(df) is df
    info = json.loads(df.get_info_str())
    assert info["schema"] == "a:str,b:str" and info["metadata"] == {} and info["type"].endswith(type(df).__name__)
    assert repr(df) == df._repr_html_() and "a:str,b:str" in repr(df)
