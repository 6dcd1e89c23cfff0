"""The slice of Fugue's column-expression DSL that reaches ``ExecutionEngine.aggregate``.

Mirrors fugue/column/expressions.py (``col``, ``all_cols``, ``alias``) and
fugue/column/functions.py:13-370 (``sum/count/min/max/avg/mean``): enough to write
``fa.aggregate(df, "key", s=f.sum(col("v0")), c=f.count(all_cols()))`` exactly as the
reference does (fugue/execution/api.py:1175-1232).
"""
from typing import Any, Optional


class ColumnExpr:
    def __init__(self, name: str, as_name: str = ""):
        self.name = name
        self.as_name = as_name

    def alias(self, as_name: str) -> "ColumnExpr":
        return ColumnExpr(self.name, as_name)

    @property
    def output_name(self) -> str:
        return self.as_name or self.name

    def __repr__(self) -> str:
        return self.name + (f" AS {self.as_name}" if self.as_name else "")


class _WildcardExpr(ColumnExpr):
    def __init__(self) -> None:
        super().__init__("*")


class AggFuncExpr(ColumnExpr):
    """``func(arg)`` with func in SUM / COUNT / MIN / MAX / AVG."""

    def __init__(self, func: str, arg: ColumnExpr, as_name: str = ""):
        super().__init__(f"{func.upper()}({arg.name})", as_name)
        self.func = func.upper()
        self.arg = arg

    def alias(self, as_name: str) -> "AggFuncExpr":
        return AggFuncExpr(self.func, self.arg, as_name)

    @property
    def output_name(self) -> str:
        # fugue/column/functions.py: an aggregation of a named column keeps that name by default
        return self.as_name or ("" if self.arg.name == "*" else self.arg.name)

    def _unsupported(self, *a: Any, **k: Any) -> Any:
        raise NotImplementedError("arithmetic / casts on aggregates are outside the hot path")

    __mul__ = __add__ = __sub__ = __truediv__ = cast = _unsupported


def col(obj: Any, alias: str = "") -> ColumnExpr:
    if isinstance(obj, ColumnExpr):
        return obj.alias(alias) if alias else obj
    if obj == "*":
        return all_cols()
    return ColumnExpr(str(obj), alias)


def all_cols() -> ColumnExpr:
    return _WildcardExpr()


class functions:  # used as ``import fugue_b200.column as fc; f = fc.functions``
    @staticmethod
    def sum(c: Any) -> AggFuncExpr:
        return AggFuncExpr("SUM", col(c))

    @staticmethod
    def count(c: Any) -> AggFuncExpr:
        return AggFuncExpr("COUNT", col(c))

    @staticmethod
    def min(c: Any) -> AggFuncExpr:
        return AggFuncExpr("MIN", col(c))

    @staticmethod
    def max(c: Any) -> AggFuncExpr:
        return AggFuncExpr("MAX", col(c))

    @staticmethod
    def avg(c: Any) -> AggFuncExpr:
        return AggFuncExpr("AVG", col(c))

    mean = avg
