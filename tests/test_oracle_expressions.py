"""Pins oracle/expressions.py to the literal tables of the reference's own conformance tests
(fugue_test/execution_suite.py:85-206: test_filter, test_select, test_assign, test_aggregate)."""
import numpy as np
import pandas as pd

from fugue_b200.column import SelectColumns, all_cols, col, functions as ff, lit
from oracle import expressions as OX

A = pd.DataFrame({"a": [1, np.nan, np.nan, 3, np.nan], "b": np.array([2, 2, 1, 4, 4], dtype="int32")})


def rows(df):
    def cell(x):
        if x is pd.NA or x is None or (isinstance(x, float) and np.isnan(x)):
            return None
        return x.item() if hasattr(x, "item") else x

    return sorted(([cell(x) for x in r] for r in df.itertuples(index=False)),
                  key=lambda r: [(v is None, v) for v in r])


def expect(data):
    return sorted(data, key=lambda r: [(v is None, v) for v in r])


def test_filter():
    assert rows(OX.filter_rows(A, col("a").not_null())) == expect([[1, 2], [3, 4]])
    assert rows(OX.filter_rows(A, col("a").not_null() & (col("b") < 3))) == [[1, 2]]
    assert rows(OX.filter_rows(A, col("a") + col("b") == 3)) == [[1, 2]]


def test_select():
    b = OX.select(A, SelectColumns(col("b"), (col("b") + 1).alias("c").cast(str)))
    assert rows(b) == expect([[2, "3"], [2, "3"], [1, "2"], [4, "5"], [4, "5"]])
    b = OX.select(A, SelectColumns(col("b"), (col("b") + 1).alias("c").cast(str), arg_distinct=True))
    assert rows(b) == expect([[2, "3"], [1, "2"], [4, "5"]])
    assert rows(OX.select(A, SelectColumns(all_cols()), where=col("a") + col("b") == 3)) == [[1, 2]]
    b = OX.select(A, SelectColumns(col("a"), ff.sum(col("b")).cast(float).alias("b")))
    assert rows(b) == expect([[1, 2], [3, 4], [None, 7]])
    col_b = ff.sum(col("b"))
    b = OX.select(A, SelectColumns(col("a"), col_b.cast(float).alias("c")), having=(col_b >= 7) | (col("a") == 1))
    assert rows(b) == expect([[1, 2], [None, 7]])
    b = OX.select(A, SelectColumns(col("a"), lit(1, "o").cast(str), col_b.cast(float).alias("c")),
                  having=(col_b >= 7) | (col("a") == 1))
    assert rows(b) == expect([[1, "1", 2], [None, "1", 7]])


def test_assign():
    b = OX.assign(A, [lit(1).alias("x"), col("b").cast(str).alias("b"), (col("b") + 1).cast(int).alias("c")])
    assert list(b.columns) == ["a", "b", "x", "c"]
    assert rows(b) == expect([[1, "2", 1, 3], [None, "2", 1, 3], [None, "1", 1, 2], [3, "4", 1, 5],
                              [None, "4", 1, 5]])


def test_aggregate_expressions():
    b = OX.select(A, SelectColumns(ff.max(col("b")).alias("b"), (ff.max(col("b")) * 2).cast("int32").alias("c")))
    assert rows(b) == [[4, 8]]
    b = OX.select(A, SelectColumns(col("a"), ff.max(col("b")).alias("b"),
                                   (ff.max(col("b")) * 2).cast("int32").alias("c")))
    assert rows(b) == expect([[None, 4, 8], [1, 2, 4], [3, 4, 8]])


def test_three_valued_logic():
    t = pd.DataFrame({"p": pd.array([True, True, True, False, False, False, None, None, None], dtype="boolean"),
                      "q": pd.array([True, False, None] * 3, dtype="boolean")})
    r = OX.select(t, SelectColumns((col("p") & col("q")).alias("a"), (col("p") | col("q")).alias("o"),
                                   (~col("p")).alias("n")))
    assert r["a"].tolist() == [True, False, pd.NA, False, False, False, pd.NA, False, pd.NA]
    assert r["o"].tolist() == [True, True, True, True, False, pd.NA, True, pd.NA, pd.NA]
    assert r["n"].tolist() == [False, False, False, True, True, True, pd.NA, pd.NA, pd.NA]


def test_count_distinct():
    t = pd.DataFrame({"k": [1, 1, 1, 2, 2, None], "x": [5.0, 5.0, np.nan, 7.0, 8.0, 9.0], "y": [1, 2, 3, 4, 5, 6]})
    r = OX.select(t, SelectColumns(col("k"), ff.count_distinct(col("x")).alias("d"), ff.sum(col("y")).alias("s")))
    assert rows(r) == expect([[1, 1, 6], [2, 2, 9], [None, 1, 6]])
    r = OX.select(pd.concat([t, t]), SelectColumns(ff.count_distinct(all_cols()).alias("n")))
    assert rows(r) == [[6]]
