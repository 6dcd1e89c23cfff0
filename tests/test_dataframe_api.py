"""The dataframe / dataset utility functions of ``fugue.api`` (fugue/dataframe/api.py,
fugue/dataset/api.py) on local frames: same names, same return-type rule (a Fugue DataFrame in ->
a Fugue DataFrame out, a native object in -> the same kind of native object out)."""
import pandas as pd
import pyarrow as pa
import pytest

from fugue_b200 import api as fa
from fugue_b200.dataframe import (ArrayDataFrame, ArrowDataFrame, DataFrame, FugueDataFrameOperationError,
                                  FugueDatasetEmptyError)
from fugue_b200.schema import Schema


def test_conversions_and_peeks():
    pdf = pd.DataFrame({"a": [1, 2, 3], "b": ["x", None, "z"]})
    assert fa.is_df(pdf) and fa.is_df(pa.Table.from_pandas(pdf)) and not fa.is_df([[1]])
    assert fa.get_column_names(pdf) == ["a", "b"] and fa.get_schema(pdf) == Schema("a:long,b:str")
    assert fa.as_array(pdf) == [[1, "x"], [2, None], [3, "z"]]
    assert fa.as_array(pdf, columns=["b"]) == [["x"], [None], ["z"]]
    assert list(fa.as_array_iterable(pdf))[0] == [1, "x"]
    assert fa.as_dicts(pdf)[1] == {"a": 2, "b": None} and next(iter(fa.as_dict_iterable(pdf))) == {"a": 1, "b": "x"}
    assert fa.peek_array(pdf) == [1, "x"] and fa.peek_dict(pdf) == {"a": 1, "b": "x"}
    assert fa.count(pdf) == 3 and not fa.is_empty(pdf) and fa.is_local(pdf) and fa.is_bounded(pdf)
    assert fa.get_num_partitions(pdf) == 1
    assert isinstance(fa.as_arrow(pdf), pa.Table) and isinstance(fa.as_pandas(pa.Table.from_pandas(pdf)), pd.DataFrame)
    with pytest.raises(FugueDatasetEmptyError):
        fa.peek_array(ArrayDataFrame([], "a:int"))
    fdf = ArrowDataFrame(pdf)
    assert fa.get_native_as_df(fdf) is fdf.native and fa.get_native_as_df(pdf) is pdf
    assert fa.as_local(pdf) is not None and isinstance(fa.as_local_bounded(fdf), DataFrame)


def test_column_operations_keep_the_input_kind():
    pdf = pd.DataFrame({"a": [1, 2, 3], "b": [1.5, 2.5, None], "c": ["x", "y", "z"]})
    r = fa.rename(pdf, {"a": "aa"})
    assert isinstance(r, pd.DataFrame) and list(r.columns) == ["aa", "b", "c"]
    assert fa.rename(pdf, {}) is pdf
    with pytest.raises(FugueDataFrameOperationError):
        fa.rename(pdf, {"zz": "a"})
    d = fa.drop_columns(pa.Table.from_pandas(pdf), ["b"])
    assert isinstance(d, pa.Table) and d.schema.names == ["a", "c"]
    with pytest.raises(FugueDataFrameOperationError):
        fa.drop_columns(pdf, ["a", "b", "c"])
    with pytest.raises(FugueDataFrameOperationError):
        fa.drop_columns(pdf, ["nope"])
    s = fa.select_columns(ArrowDataFrame(pdf), ["c", "a"])
    assert isinstance(s, DataFrame) and s.schema == Schema("c:str,a:long")
    with pytest.raises(FugueDataFrameOperationError):
        fa.select_columns(pdf, [])
    h = fa.head(pdf, 2, columns=["b", "a"])
    assert isinstance(h, pd.DataFrame) and h.values.tolist() == [[1.5, 1], [2.5, 2]]
    assert isinstance(fa.head(pdf, 1, as_fugue=True), DataFrame)


def test_alter_columns():
    fdf = ArrayDataFrame([[1, "2", 1.5], [None, "3", None]], "a:int,b:str,c:double")
    r = fa.alter_columns(fdf, "b:long,a:double")
    assert r.schema == Schema("a:double,b:long,c:double")              # order of the original schema kept
    assert r.as_array() == [[1.0, 2, 1.5], [None, 3, None]]
    assert fa.alter_columns(fdf, "a:int") is fdf                        # nothing to change
    assert fa.alter_columns(fdf, "c:str", as_fugue=True).as_array() == [[1, "2", "1.5"], [None, "3", None]]
    with pytest.raises(FugueDataFrameOperationError):
        fa.alter_columns(fdf, "zz:int")
    with pytest.raises(FugueDataFrameOperationError):
        fa.alter_columns(ArrayDataFrame([["x"]], "b:str"), "b:long")    # not castable
    pdf = pd.DataFrame({"a": [1, 2]})
    out = fa.alter_columns(pdf, "a:double")
    assert isinstance(out, pd.DataFrame) and out["a"].dtype == "float64"
