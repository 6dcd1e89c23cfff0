"""Pins oracle/native_engine.py (the pandas restatement of NativeExecutionEngine) against the
literal expectations of the reference's own conformance suite:
  fugue_test/execution_suite.py:208-256 (test_map), :258-278 (NULL keys), :400-543 (joins),
  :177-206 (test_aggregate)."""
import numpy as np
import pandas as pd

from oracle import native_engine as ora


def _rows(df):
    return sorted([[None if (isinstance(x, float) and np.isnan(x)) else x for x in r] for r in df.values.tolist()],
                  key=lambda r: [(-1e30 if x is None else x) if not isinstance(x, str) else hash(x) for x in r])


def _df(rows, cols):
    return pd.DataFrame(rows, columns=cols)


def test_map_noop_and_select_top():
    o = _df([[1, 2], [None, 2], [None, 1], [3, 4], [None, 4]], ["a", "b"])
    noop = lambda c, d: d
    top = lambda c, d: pd.DataFrame([c.row], columns=list(d.columns))
    assert _rows(ora.map_dataframe(o, noop, ["a", "b"])) == _rows(o)
    assert _rows(ora.map_dataframe(o, noop, ["a", "b"], ["a"], {"b": True})) == _rows(o)
    r = ora.map_dataframe(o, top, ["a", "b"], ["a"], {"b": True})
    assert _rows(r) == _rows(_df([[None, 1], [1, 2], [3, 4]], ["a", "b"]))
    r = ora.map_dataframe(o, top, ["a", "b"], ["a"], {"b": False})
    assert _rows(r) == _rows(_df([[None, 4], [1, 2], [3, 4]], ["a", "b"]))


def test_map_multi_key_with_nulls():
    o = _df([[1, None, 1], [1, None, 0], [None, None, 2]], ["a", "b", "c"])
    top = lambda c, d: pd.DataFrame([c.row], columns=list(d.columns))
    r = ora.map_dataframe(o, top, ["a", "b", "c"], ["a", "b"], {"c": True})
    assert _rows(r) == _rows(_df([[1, None, 0], [None, None, 2]], ["a", "b", "c"]))


def test_map_cursor_partition_numbers_and_keys():
    o = _df([[2, "x"], [1, "y"], [2, "z"]], ["k", "v"])
    seen = []

    def f(cursor, d):
        seen.append((cursor.partition_no, cursor.key_value_array, cursor.key_value_dict, len(d)))
        return d

    ora.map_dataframe(o, f, ["k", "v"], ["k"])
    assert seen == [(1, [1], {"k": 1}, 1), (2, [2], {"k": 2}, 2)]   # ascending keys, 1-based numbering


def test_join_truth_tables():
    a = _df([[1, 2], [3, 4]], ["a", "b"])
    b = _df([[6, 1], [2, 7]], ["c", "a"])
    assert ora.join(a, b, "inner", ["a"]).values.tolist() == [[1, 2, 6]]
    assert ora.join(b, a, "inner").values.tolist() == [[6, 1, 2]]
    assert ora.join(a, b, "semi", ["a"]).values.tolist() == [[1, 2]]
    assert ora.join(a, b, "anti", ["a"]).values.tolist() == [[3, 4]]
    assert _rows(ora.join(a, _df([[6], [7]], ["c"]), "cross")) == [[1, 2, 6], [1, 2, 7], [3, 4, 6], [3, 4, 7]]
    a = _df([[1, "2"], [3, "4"]], ["a", "b"])
    b = _df([["6", 1], ["2", 7]], ["c", "a"])
    assert _rows(ora.join(a, b, "left_outer", ["a"])) == _rows(_df([[1, "2", "6"], [3, "4", None]], list("abc")))
    assert _rows(ora.join(a, b, "right_outer", ["a"])) == _rows(_df([[1, "2", "6"], [7, None, "2"]], list("abc")))
    assert len(ora.join(a, b, "full_outer", ["a"])) == 3
    # SQL will not match null values (execution_suite.py:533-543)
    a = _df([[1, 2, 3], [4, None, 6]], ["a", "b", "c"])
    b = _df([[1, 2, 33], [4, None, 63]], ["a", "b", "d"])
    assert ora.join(a, b, "inner").values.tolist() == [[1, 2.0, 3, 33]]


def test_aggregate_with_null_key_group():
    a = _df([[1, 2], [None, 2], [None, 1], [3, 4], [None, 4]], ["a", "b"])
    r = ora.aggregate(a, ["a"], {"b": ("b", "max")})
    assert _rows(r) == _rows(_df([[None, 4], [1, 2], [3, 4]], ["a", "b"]))
    r = ora.aggregate(a, [], {"b": ("b", "max")})
    assert r.values.tolist() == [[4]]
    r = ora.aggregate_sum_count(_df([[1, 1.5], [1, 2.5], [2, 1.0]], ["key", "v0"]), ["key"], "v0")
    assert r.values.tolist() == [[1, 4.0, 2], [2, 1.0, 1]]
