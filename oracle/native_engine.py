"""Oracle: pandas restatement of the reference NativeExecutionEngine on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line where it matters:

* ``PandasMapEngine.map_dataframe``  fugue/execution/native_execution_engine.py:81-169
    - no keys / coarse branch  :104-154
    - keyed branch ``_map``    :156-169  (presort :157-160, cursor.set :162,
      ``safe_groupby_apply`` :166-168 = triad ``PandasUtils.safe_groupby_apply``:
      groupby(keys, dropna=False) + apply, ascending key order, NULL keys grouped)
* ``NativeExecutionEngine.join``     fugue/execution/native_execution_engine.py:230-241
    + ``get_join_schemas``           fugue/dataframe/utils.py:152-226
    + triad ``PandasUtils.join`` (third party, triad>=1.0.0 setup.py:34): pd.merge
      on the key columns after dropping rows with NULL keys (SQL semantics, pinned by
      fugue_test/execution_suite.py:533-543)
* ``ExecutionEngine.aggregate``      fugue/execution/execution_engine.py:889-939
    -> ``SELECT keys, AGG(..) .. GROUP BY keys`` (fugue/column/sql.py:275-334) run by
    qpd on pandas (third party) == pandas groupby(dropna=False).agg

pandas 3.0.2 (this image) drops the key columns from the frame handed to
``groupby.apply`` callbacks, which the reference (written for pandas 2.x) relies
on, so groups are *iterated* (iteration still yields complete sub-frames, in
ascending key order, stable inside a group) - SURVEY.md F4.
"""
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd


class OracleCursor:
    """Restates fugue/collections/partition.py:404-469 (PartitionCursor)."""

    def __init__(self, columns: List[str], keys: List[str], physical_partition_no: int = 0):
        self._columns = list(columns)
        self._key_index = [self._columns.index(k) for k in keys]
        self._physical_partition_no = physical_partition_no
        self._partition_no = 0
        self._slice_no = 0
        self._item: Any = None

    def set(self, row: Any, partition_no: int, slice_no: int) -> None:
        self._item = (lambda: list(row())) if callable(row) else list(row)
        self._partition_no = partition_no
        self._slice_no = slice_no

    @property
    def row(self) -> List[Any]:
        if callable(self._item):
            self._item = self._item()
        return self._item

    @property
    def partition_no(self) -> int:
        return self._partition_no

    @property
    def physical_partition_no(self) -> int:
        return self._physical_partition_no

    @property
    def slice_no(self) -> int:
        return self._slice_no

    @property
    def key_value_array(self) -> List[Any]:
        return [self.row[i] for i in self._key_index]

    @property
    def key_value_dict(self) -> Dict[str, Any]:
        return {self._columns[i]: self.row[i] for i in self._key_index}

    def __getitem__(self, key: str) -> Any:
        return self.row[self._columns.index(key)]


def _first_row(pdf: pd.DataFrame) -> List[Any]:
    # PandasDataFrame.peek_array  fugue/dataframe/pandas_dataframe.py:110-112
    return pdf.iloc[0].values.tolist()


def map_dataframe(
    pdf: pd.DataFrame,
    map_func: Callable[[OracleCursor, pd.DataFrame], pd.DataFrame],
    output_columns: Sequence[str],
    partition_by: Sequence[str] = (),
    presort: Optional[Dict[str, bool]] = None,
    num_partitions: int = 0,
    on_init: Optional[Callable[[int, pd.DataFrame], Any]] = None,
) -> pd.DataFrame:
    """native_execution_engine.py:81-169 on plain pandas frames."""
    presort = dict(presort or {})
    presort_keys = list(presort.keys())
    presort_asc = list(presort.values())
    keys = list(partition_by)
    cursor = OracleCursor(list(pdf.columns), keys, 0)
    if on_init is not None:
        on_init(0, pdf)
    if len(keys) == 0:  # :104-154
        if presort_keys:
            pdf = pdf.sort_values(presort_keys, ascending=presort_asc).reset_index(drop=True)
        if num_partitions > 0:
            outs: List[pd.DataFrame] = []
            for p, sub in enumerate(np.array_split(pdf, num_partitions)):
                if len(sub) > 0:
                    sub = sub.reset_index(drop=True)
                    cursor.set(lambda s=sub: _first_row(s), p, 0)
                    outs.append(map_func(cursor, sub))
            res = pd.concat(outs, ignore_index=True) if outs else pdf.iloc[:0]
        else:
            cursor.set(lambda: _first_row(pdf), 0, 0)
            res = map_func(cursor, pdf)
        return res[list(output_columns)].reset_index(drop=True)

    outs = []
    if len(pdf) > 0:
        by: Any = keys if len(keys) > 1 else keys  # always a list -> tuple group names
        for _, sub in pdf.groupby(by, dropna=False, sort=True):  # :166-168
            if presort_keys:  # :157-160
                sub = sub.sort_values(presort_keys, ascending=presort_asc)
            sub = sub.reset_index(drop=True)
            cursor.set(lambda s=sub: _first_row(s), cursor.partition_no + 1, 0)  # :162
            outs.append(map_func(cursor, sub))  # :163
    if not outs:
        return pd.DataFrame({c: [] for c in output_columns})
    return pd.concat(outs, ignore_index=True)[list(output_columns)]


def join_schemas(cols1: Sequence[str], cols2: Sequence[str], how: str,
                 on: Optional[Sequence[str]]) -> Tuple[List[str], List[str]]:
    """fugue/dataframe/utils.py:152-226 on column-name lists."""
    how = how.lower()
    if how not in ["semi", "left_semi", "anti", "left_anti", "inner", "left_outer",
                   "right_outer", "full_outer", "cross"]:
        raise ValueError(f"{how} is not a valid join type")
    on = list(on) if on is not None else []
    if how != "cross" and len(on) == 0:
        other = set(cols2)
        on = [c for c in cols1 if c in other]
        if not on:
            raise KeyError(f"no common columns between {cols1} and {cols2}")
    c2 = list(cols2)
    if how in ["semi", "left_semi", "anti", "left_anti"]:
        c2 = [c for c in c2 if c in on]
    if how == "cross":
        if set(cols1) & set(c2):
            raise KeyError("invalid cross join, two dataframes have common columns")
    return on, list(cols1) + [c for c in c2 if c not in cols1]


def join(pdf1: pd.DataFrame, pdf2: pd.DataFrame, how: str,
         on: Optional[Sequence[str]] = None) -> pd.DataFrame:
    """native_execution_engine.py:230-241; NULL keys never match."""
    keys, out_cols = join_schemas(list(pdf1.columns), list(pdf2.columns), how, on)
    how = how.lower()
    if how == "cross":
        return pdf1.merge(pdf2, how="cross")[out_cols].reset_index(drop=True)
    d1 = pdf1
    d2 = pdf2
    if how in ("inner", "semi", "left_semi"):
        d1 = pdf1.dropna(subset=keys)
    d2n = pdf2.dropna(subset=keys)
    if how == "inner":
        res = d1.merge(d2n, how="inner", on=keys)
    elif how in ("semi", "left_semi"):
        res = d1.merge(d2n[keys].drop_duplicates(), how="inner", on=keys)
    elif how in ("anti", "left_anti"):
        m = d1.merge(d2n[keys].drop_duplicates(), how="left", on=keys, indicator=True)
        res = m[m["_merge"] == "left_only"].drop(columns=["_merge"])
    elif how == "left_outer":
        res = d1.merge(d2n, how="left", on=keys)
    elif how == "right_outer":
        res = pdf1.dropna(subset=keys).merge(pdf2, how="right", on=keys)
    elif how == "full_outer":
        a = pdf1.dropna(subset=keys).merge(d2n, how="outer", on=keys)
        n1 = pdf1[pdf1[keys].isna().any(axis=1)]
        n2 = pdf2[pdf2[keys].isna().any(axis=1)]
        res = pd.concat([a, n1, n2], ignore_index=True)
    else:  # pragma: no cover
        raise ValueError(how)
    return res[out_cols].reset_index(drop=True)


def aggregate_sum_count(pdf: pd.DataFrame, keys: Sequence[str], value: str,
                        sum_name: str = "s", count_name: str = "c") -> pd.DataFrame:
    """``SELECT keys, SUM(value) AS s, COUNT(*) AS c FROM t GROUP BY keys``.
    NULL key is a group (fugue_test/execution_suite.py:195-200)."""
    g = pdf.groupby(list(keys), dropna=False, sort=True)
    res = g.agg(**{sum_name: (value, "sum"), count_name: (value, "size")}).reset_index()
    res[count_name] = res[count_name].astype("int64")
    return res


def aggregate(pdf: pd.DataFrame, keys: Sequence[str],
              aggs: Dict[str, Tuple[str, str]]) -> pd.DataFrame:
    """General form: aggs = {out_name: (column or '*', func)} with func in
    sum/count/min/max/avg (SQL semantics: NULL values skipped, COUNT(*) counts rows)."""
    fm = {"sum": "sum", "min": "min", "max": "max", "avg": "mean", "mean": "mean"}
    named = {}
    for name, (colname, func) in aggs.items():
        func = func.lower()
        if func == "count":
            if colname == "*":
                named[name] = (pdf.columns[0], "size")
            else:
                named[name] = (colname, "count")
        else:
            named[name] = (colname, fm[func])
    if len(keys) == 0:
        row = {}
        for name, (colname, f) in named.items():
            row[name] = [len(pdf) if f == "size" else getattr(pdf[colname], f)()]
        return pd.DataFrame(row)
    g = pdf.groupby(list(keys), dropna=False, sort=True)
    return g.agg(**named).reset_index()
