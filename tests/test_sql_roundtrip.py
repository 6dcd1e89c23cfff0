"""SQL text round trip on random trees: what the printer (``fugue_b200.column.to_sql`` - pinned to the
reference's SQL generator by tests/test_column_golden.py) writes, the SELECT parser of the SQL engine
(``fugue_b200.sql._parse_select``) must read back as the same tree.  This is the path ``ExecutionEngine.
aggregate / select`` take through ``SQLEngine.select`` in the reference (fugue/column/sql.py:275-347 ->
execution_engine.py:209-238).  CPU only, seeded."""
import numpy as np
import pytest

from fugue_b200.column import ColumnExpr, Kind, SelectColumns, col, functions as ff, select_sql, to_sql
from fugue_b200.sql import _parse_select
from oracle import expressions as OX
from test_expr_compiler import _random, _same
from test_expr_random import _boolean, _literal_only, _numeric

PDF = _random(n=300, seed=11)


def _same_meaning(parsed, original, text):
    """Same tree, or - where the text is ambiguous about shape (``--5`` is ``-(-5)`` or ``-(-(5))``) - the same
    values on every row of a random frame."""
    if str(parsed) == str(original):
        return
    if _literal_only(original):
        return
    a = OX.select(PDF, SelectColumns(parsed.alias("r")))["r"]
    b = OX.select(PDF, SelectColumns(original.alias("r")))["r"]
    _same(a, b, text)


def _parse(items: str, rest: str):
    return _parse_select(items, rest, f"SELECT {items} FROM {rest}")


def _has_inner_named_cast(e, root=True) -> bool:
    """A bare column with a cast INSIDE an expression: the reference's generator (and therefore this printer, which
    is pinned to it by the golden vectors) writes ``CAST(a AS long) AS a`` there too (fugue/column/sql.py:405-431,
    vector "CAST(a AS double) AS a+b AS ci") - not valid SQL, no engine could read it; such trees are left out."""
    if not isinstance(e, ColumnExpr):
        return False
    if not root and e.kind == Kind.NAMED and e.as_type is not None:
        return True
    return any(_has_inner_named_cast(a, False) for a in e.args)


@pytest.mark.parametrize("seed", range(4))
def test_printed_expressions_parse_back(seed):
    rng = np.random.default_rng(77 + seed)
    for _ in range(150):
        e = (_numeric if rng.random() < 0.5 else _boolean)(rng, int(rng.integers(1, 5))).alias("r")
        if _has_inner_named_cast(e):
            continue
        text = to_sql(e)
        st = _parse(text, "t")
        assert len(st.columns) == 1 and st.columns[0].output_name == "r", text
        _same_meaning(st.columns[0], e, text)
        where = _boolean(rng, int(rng.integers(1, 4)))
        if _has_inner_named_cast(where, False):
            continue
        st = _parse("x", "t WHERE " + to_sql(where))
        _same_meaning(st.where, where, to_sql(where))


def _fold(text: str) -> str:
    """``-(4)`` (negated literal) and ``-4`` (negative literal) print alike: compare them as one."""
    import re

    return re.sub(r"-\((\d+(?:\.\d+)?)\)", r"-\1", text)


def test_printed_aggregating_selects_parse_back():
    rng = np.random.default_rng(5)
    for _ in range(100):
        keys = [col(n) for n in rng.choice(["a", "b", "g"], size=int(rng.integers(0, 3)), replace=False)]
        aggs = []
        for i in range(int(rng.integers(1, 4))):
            f = [ff.sum, ff.min, ff.max, ff.avg, ff.count][rng.integers(5)]
            aggs.append(f(_numeric(rng, int(rng.integers(0, 3)))).alias(f"m{i}"))
        cols = SelectColumns(*keys, *aggs)
        where = _boolean(rng, 2) if rng.random() < 0.5 else None
        if any(_has_inner_named_cast(x, False) for x in aggs + ([where] if where is not None else [])):
            continue
        having = (aggs[0].alias("") > int(rng.integers(0, 9))) if rng.random() < 0.5 else None
        sql = select_sql(cols, "t", where, having)
        assert sql.startswith("SELECT ")
        items, rest = sql[len("SELECT "):].split(" FROM ", 1)
        st = _parse(items, rest)
        assert [_fold(str(c)) for c in st.columns] == [_fold(str(c)) for c in cols.all_cols], sql
        assert [str(g) for g in st.group_by] == [str(k) for k in keys], sql
        assert (st.where is None) == (where is None), sql
        if where is not None:
            _same_meaning(st.where, where, sql)
        assert (st.having is None) == (having is None), sql
        assert having is None or _fold(str(st.having)) == _fold(str(having)), sql
