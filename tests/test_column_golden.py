"""fugue_b200.column against golden vectors produced by the REFERENCE's own column DSL code run in the build
container (tests/golden/make_column_golden.py -> tests/golden/column_dsl_vectors.json): string forms, generated
SQL, alias / type inference and SELECT classification of 63 expressions and 9 statements."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from column_catalogue import describe_all  # noqa: E402

from fugue_b200 import column as bc  # noqa: E402
from fugue_b200.schema import Schema  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "column_dsl_vectors.json")


def test_column_dsl_matches_the_reference_vectors():
    class Printer:  # the catalogue speaks to the reference's SQLExpressionGenerator; ours are two functions
        def __init__(self, enable_cast=True):
            self.enable_cast = enable_cast

        def generate(self, e):
            return bc.to_sql(e, self.enable_cast)

        def select(self, cols, table, where=None, having=None):
            return [(False, bc.select_sql(cols, table, where, having, self.enable_cast))]

    ns = types.SimpleNamespace(col=bc.col, lit=bc.lit, null=bc.null, all_cols=bc.all_cols, function=bc.function,
                               f=bc.functions, SelectColumns=bc.SelectColumns,
                               SQLExpressionGenerator=Printer, Schema=Schema)
    got = describe_all(ns)
    want = json.load(open(GOLDEN))
    assert set(got["expressions"]) == set(want["expressions"]) and set(got["selects"]) == set(want["selects"])
    bad = []
    for kind in ("expressions", "selects"):
        for name, exp in want[kind].items():
            for field, val in exp.items():
                if got[kind][name].get(field) != val:
                    bad.append((kind, name, field, got[kind][name].get(field), val))
    assert not bad, "\n".join(f"{k}.{n}.{f}: got {g!r}, reference {w!r}" for k, n, f, g, w in bad)
