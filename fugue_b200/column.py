"""Fugue's column-expression DSL, as consumed by ``ExecutionEngine.select/filter/assign/aggregate``.

Same public surface and observable behaviour as the reference (so code written against
``fugue.column`` runs unchanged), own implementation:

* expressions   fugue/column/expressions.py:8-856  (``col, lit, null, all_cols, function``, operators,
                ``alias / cast / infer_alias / infer_type / is_null / not_null``, ``str()`` format)
* functions     fugue/column/functions.py:13-370   (``coalesce, min, max, count, count_distinct, avg,
                sum, first, last, is_agg``)
* SelectColumns fugue/column/sql.py:38-246         (classification of SELECT columns, group-key inference)
* SQLExpressionGenerator fugue/column/sql.py:249-497 (the SQL text Fugue itself emits for these trees)

The B200 engine does not go through SQL text: ``fugue_b200/expr.py`` compiles these trees into programs
for the device evaluator (``fb_eval_expr``).  The generator is kept because it is part of the reference
surface (``fugue_plugin`` hands unrecognised statements to the host SQL engine with it).
"""
import hashlib
from typing import Any, Callable, Dict, Iterable, List, Optional, Set, Tuple

import pyarrow as pa

from .schema import Schema, parse_type, type_to_expr


def to_pa_datatype(obj: Any) -> pa.DataType:
    """python type / type expression / pyarrow type -> pyarrow type (triad ``to_pa_datatype`` rules:
    ``int`` -> int64, ``float`` -> float64, ``str`` -> string, ``bool`` -> bool; strings follow the
    schema expression syntax, so ``"int"`` -> int32)."""
    import datetime

    if isinstance(obj, pa.DataType):
        return obj
    if obj is int:
        return pa.int64()
    if obj is float:
        return pa.float64()
    if obj is str:
        return pa.string()
    if obj is bool:
        return pa.bool_()
    if obj is datetime.datetime:
        return pa.timestamp("us")
    if obj is datetime.date:
        return pa.date32()
    if isinstance(obj, str):
        return parse_type(obj)
    raise TypeError(f"can't convert {obj!r} to a data type")


def _quote_name(name: str) -> str:
    ok = name != "" and (name[0].isalpha() or name[0] == "_") and all(c.isalnum() or c == "_" for c in name)
    return name if ok else "`" + name.replace("`", "``") + "`"


class ColumnExpr:
    """Base of all column expressions; build them with :func:`col`, :func:`lit`, :func:`null`,
    :func:`all_cols`, :func:`function` and the operators."""

    def __init__(self) -> None:
        self._as_name = ""
        self._as_type: Optional[pa.DataType] = None

    @property
    def name(self) -> str:
        return ""

    @property
    def as_name(self) -> str:
        return self._as_name

    @property
    def as_type(self) -> Optional[pa.DataType]:
        return self._as_type

    @property
    def output_name(self) -> str:
        return self.as_name if self.as_name != "" else self.name

    def alias(self, as_name: str) -> "ColumnExpr":
        raise NotImplementedError

    def cast(self, data_type: Any) -> "ColumnExpr":
        raise NotImplementedError

    def infer_alias(self) -> "ColumnExpr":
        return self

    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        return self.as_type

    @property
    def body_str(self) -> str:
        raise NotImplementedError

    def __str__(self) -> str:
        res = self.body_str
        if self.as_type is not None:
            res = f"CAST({res} AS {type_to_expr(self.as_type)})"
        if self.as_name != "":
            res = res + " AS " + self.as_name
        return res

    __repr__ = __str__

    def is_null(self) -> "ColumnExpr":
        return _UnaryOpExpr("IS_NULL", self)

    def not_null(self) -> "ColumnExpr":
        return _UnaryOpExpr("NOT_NULL", self)

    def __neg__(self) -> "ColumnExpr":
        return _InvertOpExpr("-", self)

    def __pos__(self) -> "ColumnExpr":
        return self

    def __invert__(self) -> "ColumnExpr":
        return _NotOpExpr("~", self)

    def __add__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("+", self, other)

    def __radd__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("+", other, self)

    def __sub__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("-", self, other)

    def __rsub__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("-", other, self)

    def __mul__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("*", self, other)

    def __rmul__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("*", other, self)

    def __truediv__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("/", self, other)

    def __rtruediv__(self, other: Any) -> "ColumnExpr":
        return _BinaryOpExpr("/", other, self)

    def __and__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr("&", self, other)

    def __rand__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr("&", other, self)

    def __or__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr("|", self, other)

    def __ror__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr("|", other, self)

    def __lt__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr("<", self, other)

    def __gt__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr(">", self, other)

    def __le__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr("<=", self, other)

    def __ge__(self, other: Any) -> "ColumnExpr":
        return _BoolBinaryOpExpr(">=", self, other)

    def __eq__(self, other: Any) -> "ColumnExpr":  # type: ignore
        return _BoolBinaryOpExpr("==", self, other)

    def __ne__(self, other: Any) -> "ColumnExpr":  # type: ignore
        return _BoolBinaryOpExpr("!=", self, other)

    __hash__ = object.__hash__  # == builds an expression, identity is the only usable hash

    def __bool__(self) -> bool:
        raise TypeError("a column expression has no truth value; use & | ~ to combine conditions")

    def _uuid_keys(self) -> List[Any]:
        raise NotImplementedError

    def __uuid__(self) -> str:
        return to_uuid(type(self).__name__, self.as_name,
                       None if self.as_type is None else str(self.as_type), self._uuid_keys())


def to_uuid(*args: Any) -> str:
    """Deterministic id of nested python values / objects exposing ``__uuid__`` (triad ``to_uuid``)."""
    h = hashlib.md5()

    def feed(v: Any) -> None:
        if hasattr(v, "__uuid__"):
            h.update(b"u" + v.__uuid__().encode())
        elif isinstance(v, (list, tuple)):
            h.update(b"[")
            for x in v:
                feed(x)
            h.update(b"]")
        elif isinstance(v, dict):
            h.update(b"{")
            for k, x in v.items():
                feed(k)
                feed(x)
            h.update(b"}")
        else:
            h.update((type(v).__name__ + ":" + repr(v)).encode())

    feed(args)
    return h.hexdigest()


class _NamedColumnExpr(ColumnExpr):
    def __init__(self, name: Any):
        super().__init__()
        self._name = name

    @property
    def body_str(self) -> str:
        return self._name

    @property
    def name(self) -> str:
        return self._name

    def _derive(self, as_name: str, as_type: Optional[pa.DataType]) -> "ColumnExpr":
        other = _NamedColumnExpr(self._name)
        other._as_name, other._as_type = as_name, as_type
        return other

    def alias(self, as_name: str) -> ColumnExpr:
        return self._derive(as_name, self.as_type)

    def cast(self, data_type: Any) -> ColumnExpr:
        return self._derive(self.as_name, None if data_type is None else to_pa_datatype(data_type))

    def infer_alias(self) -> ColumnExpr:
        if self.as_name == "" and self.as_type is not None:
            return self.alias(self.output_name)
        return self

    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        if self.name not in schema:
            return self.as_type
        return self.as_type or schema[self.name].type

    def _uuid_keys(self) -> List[Any]:
        return [self._name]


class _WildcardExpr(ColumnExpr):
    @property
    def body_str(self) -> str:
        return "*"

    @property
    def name(self) -> str:
        return "*"

    @property
    def output_name(self) -> str:
        raise NotImplementedError("wildcard column doesn't have an output name")

    def alias(self, as_name: str) -> ColumnExpr:
        raise NotImplementedError("wildcard column can't have an alias")

    def cast(self, data_type: Any) -> ColumnExpr:
        raise NotImplementedError("wildcard column can't be cast")

    def infer_alias(self) -> ColumnExpr:
        return self

    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        return None

    def __uuid__(self) -> str:
        return to_uuid("*")


class _LiteralColumnExpr(ColumnExpr):
    _VALID_TYPES = (int, bool, float, str)

    def __init__(self, value: Any):
        super().__init__()
        if not (value is None or isinstance(value, _LiteralColumnExpr._VALID_TYPES)):
            raise NotImplementedError(f"{value}, type: {type(value)}")
        self._value = value

    @property
    def body_str(self) -> str:
        v = self._value
        if v is None:
            return "NULL"
        if isinstance(v, str):
            return "'" + v.replace("\\", "\\\\").replace("'", "\\'") + "'"
        if isinstance(v, bool):
            return "TRUE" if v else "FALSE"
        return str(v)

    @property
    def value(self) -> Any:
        return self._value

    def is_null(self) -> ColumnExpr:
        return _LiteralColumnExpr(self._value is None)

    def not_null(self) -> ColumnExpr:
        return _LiteralColumnExpr(self._value is not None)

    def _derive(self, as_name: str, as_type: Optional[pa.DataType]) -> ColumnExpr:
        other = _LiteralColumnExpr(self._value)
        other._as_name, other._as_type = as_name, as_type
        return other

    def alias(self, as_name: str) -> ColumnExpr:
        return self._derive(as_name, self.as_type)

    def cast(self, data_type: Any) -> ColumnExpr:
        return self._derive(self.as_name, None if data_type is None else to_pa_datatype(data_type))

    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        if self._value is None:
            return self.as_type
        return self.as_type or to_pa_datatype(type(self._value))

    def _uuid_keys(self) -> List[Any]:
        return [self._value]


class _FuncExpr(ColumnExpr):
    def __init__(self, func: str, *args: Any, arg_distinct: bool = False, **kwargs: Any):
        super().__init__()
        self._func = func
        self._distinct = arg_distinct
        self._args = list(args)
        self._kwargs = dict(kwargs)

    @property
    def body_str(self) -> str:
        def show(v: Any) -> str:
            if isinstance(v, bool):
                return "TRUE" if v else "FALSE"
            if isinstance(v, str):
                return f"'{v}'"
            return str(v)

        parts = [show(x) for x in self._args] + [k + "=" + show(v) for k, v in self._kwargs.items()]
        return f"{self._func}({'DISTINCT ' if self._distinct else ''}{','.join(parts)})"

    @property
    def func(self) -> str:
        return self._func

    @property
    def is_distinct(self) -> bool:
        return self._distinct

    @property
    def args(self) -> List[Any]:
        return self._args

    @property
    def kwargs(self) -> Dict[str, Any]:
        return self._kwargs

    def _copy(self) -> "_FuncExpr":
        return _FuncExpr(self._func, *self._args, **self._kwargs)

    def _derive(self, as_name: str, as_type: Optional[pa.DataType]) -> ColumnExpr:
        other = self._copy()
        other._distinct = self._distinct
        other._as_name, other._as_type = as_name, as_type
        return other

    def alias(self, as_name: str) -> ColumnExpr:
        return self._derive(as_name, self.as_type)

    def cast(self, data_type: Any) -> ColumnExpr:
        return self._derive(self.as_name, None if data_type is None else to_pa_datatype(data_type))

    def _uuid_keys(self) -> List[Any]:
        return [self._func, self._distinct, self._args, self._kwargs]


class _UnaryOpExpr(_FuncExpr):
    def __init__(self, op: str, column: ColumnExpr, arg_distinct: bool = False):
        super().__init__(op, column, arg_distinct=arg_distinct)

    @property
    def col(self) -> ColumnExpr:
        return self._args[0]

    @property
    def op(self) -> str:
        return self._func

    def infer_alias(self) -> ColumnExpr:
        return self if self.output_name != "" else self.alias(self.col.infer_alias().output_name)

    def _copy(self) -> _FuncExpr:
        return type(self)(self._func, self._args[0])


class _InvertOpExpr(_UnaryOpExpr):  # arithmetic negation
    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        if self.as_type is not None:
            return self.as_type
        tp = self.col.infer_type(schema)
        if tp is not None and (pa.types.is_signed_integer(tp) or pa.types.is_floating(tp)):
            return tp
        return None


class _NotOpExpr(_UnaryOpExpr):  # logical NOT
    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        if self.as_type is not None:
            return self.as_type
        tp = self.col.infer_type(schema)
        if tp is not None and pa.types.is_boolean(tp):
            return tp
        return None


class _BinaryOpExpr(_FuncExpr):
    def __init__(self, op: str, left: Any, right: Any, arg_distinct: bool = False):
        super().__init__(op, _to_col(left), _to_col(right), arg_distinct=arg_distinct)

    @property
    def left(self) -> ColumnExpr:
        return self._args[0]

    @property
    def right(self) -> ColumnExpr:
        return self._args[1]

    @property
    def op(self) -> str:
        return self._func

    def _copy(self) -> _FuncExpr:
        return type(self)(self._func, self._args[0], self._args[1])


class _BoolBinaryOpExpr(_BinaryOpExpr):
    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        return self.as_type or pa.bool_()


class AggFuncExpr(_FuncExpr):
    """``FUNC([DISTINCT] arg)`` (reference: ``_UnaryAggFuncExpr``, fugue/column/functions.py:343-356)."""

    def __init__(self, func: str, arg: ColumnExpr, as_name: str = "", arg_distinct: bool = False):
        super().__init__(func.upper(), arg, arg_distinct=arg_distinct)
        self._as_name = as_name

    @property
    def arg(self) -> ColumnExpr:
        return self._args[0]

    def infer_alias(self) -> ColumnExpr:
        return self if self.output_name != "" else self.alias(self.arg.infer_alias().output_name)

    def _copy(self) -> _FuncExpr:
        return type(self)(self._func, self._args[0], arg_distinct=self._distinct)


class _SameTypeAggFuncExpr(AggFuncExpr):  # MIN / MAX / FIRST / LAST keep the argument type
    def infer_type(self, schema: Schema) -> Optional[pa.DataType]:
        return self.as_type or self.arg.infer_type(schema)


_UnaryAggFuncExpr = AggFuncExpr  # the reference's name for it


def lit(obj: Any, alias: str = "") -> ColumnExpr:
    res = _LiteralColumnExpr(obj)
    return res if alias == "" else res.alias(alias)


def null() -> ColumnExpr:
    return lit(None)


def col(obj: Any, alias: str = "") -> ColumnExpr:
    if isinstance(obj, ColumnExpr):
        return obj if alias == "" else obj.alias(alias)
    if isinstance(obj, str):
        if obj == "*":
            return all_cols()
        res = _NamedColumnExpr(obj)
        return res if alias == "" else res.alias(alias)
    raise NotImplementedError(obj)


def all_cols() -> ColumnExpr:
    return _WildcardExpr()


def function(name: str, *args: Any, arg_distinct: bool = False, **kwargs: Any) -> ColumnExpr:
    return _FuncExpr(name, *args, arg_distinct=arg_distinct, **kwargs)


def _to_col(obj: Any) -> ColumnExpr:
    return obj if isinstance(obj, ColumnExpr) else lit(obj)


def _get_column_mentions(column: Any) -> Iterable[str]:
    if isinstance(column, _NamedColumnExpr):
        yield column.name
    elif isinstance(column, _FuncExpr):
        for a in column.args:
            yield from _get_column_mentions(a)
        for a in column.kwargs.values():
            yield from _get_column_mentions(a)


def is_agg(column: Any) -> bool:
    """True when the expression contains an aggregation anywhere (functions.py:314-340)."""
    if isinstance(column, AggFuncExpr):
        return True
    if isinstance(column, _FuncExpr):
        return any(is_agg(x) for x in column.args) or any(is_agg(x) for x in column.kwargs.values())
    return False


class functions:
    """``import fugue_b200.column as fc; f = fc.functions`` == ``import fugue.column.functions as f``."""

    @staticmethod
    def coalesce(*args: Any) -> ColumnExpr:
        return function("COALESCE", *[_to_col(x) for x in args])

    @staticmethod
    def min(c: Any) -> ColumnExpr:  # noqa: A003
        return _SameTypeAggFuncExpr("MIN", col(c))

    @staticmethod
    def max(c: Any) -> ColumnExpr:  # noqa: A003
        return _SameTypeAggFuncExpr("MAX", col(c))

    @staticmethod
    def first(c: Any) -> ColumnExpr:
        return _SameTypeAggFuncExpr("FIRST", col(c))

    @staticmethod
    def last(c: Any) -> ColumnExpr:
        return _SameTypeAggFuncExpr("LAST", col(c))

    @staticmethod
    def count(c: Any) -> ColumnExpr:
        return AggFuncExpr("COUNT", col(c))

    @staticmethod
    def count_distinct(c: Any) -> ColumnExpr:
        return AggFuncExpr("COUNT", col(c), arg_distinct=True)

    @staticmethod
    def avg(c: Any) -> ColumnExpr:
        return AggFuncExpr("AVG", col(c))

    @staticmethod
    def sum(c: Any) -> ColumnExpr:  # noqa: A003
        return AggFuncExpr("SUM", col(c))

    mean = avg
    is_agg = staticmethod(is_agg)


# ---------------------------------------------------------------------------------------------
# SELECT column collections and the SQL text Fugue emits for them
# ---------------------------------------------------------------------------------------------
class SelectColumns:
    """The columns of one ``SELECT`` (fugue/column/sql.py:38-246): literals, plain columns, non-aggregate
    functions, aggregations; group keys are every non-aggregate, non-literal column when an aggregation
    is present."""

    def __init__(self, *cols: ColumnExpr, arg_distinct: bool = False):
        self._distinct = arg_distinct
        self._all: List[ColumnExpr] = []
        self._literals: List[ColumnExpr] = []
        self._cols: List[ColumnExpr] = []
        self._non_agg_funcs: List[ColumnExpr] = []
        self._agg_funcs: List[ColumnExpr] = []
        self._group_keys: List[ColumnExpr] = []
        self._has_wildcard = False
        keys: List[ColumnExpr] = []
        for c in cols:
            c = c.infer_alias()
            self._all.append(c)
            if isinstance(c, _LiteralColumnExpr):
                self._literals.append(c)
                continue
            agg = False
            if isinstance(c, _WildcardExpr):
                if self._has_wildcard:
                    raise ValueError("'*' can be used at most once")
                self._has_wildcard = True
                self._cols.append(c)
            elif isinstance(c, _NamedColumnExpr):
                self._cols.append(c)
            elif isinstance(c, _FuncExpr):
                agg = is_agg(c)
                (self._agg_funcs if agg else self._non_agg_funcs).append(c)
            if not agg:
                keys.append(c if isinstance(c, _WildcardExpr) else c.alias("").cast(None))
        if len(self._agg_funcs) > 0:
            self._group_keys = keys
            if self._has_wildcard:
                raise ValueError(f"'*' can't be used in aggregation: {self}")

    def __str__(self) -> str:
        return "[" + ", ".join(str(x) for x in self._all) + "]"

    def __uuid__(self) -> str:
        return to_uuid(self._distinct, self._all)

    @property
    def is_distinct(self) -> bool:
        return self._distinct

    def replace_wildcard(self, schema: Schema) -> "SelectColumns":
        out: List[ColumnExpr] = []
        for c in self._all:
            if isinstance(c, _WildcardExpr):
                out.extend(col(n) for n in schema.names)
            else:
                out.append(c)
        return SelectColumns(*out, arg_distinct=self._distinct)

    def assert_all_with_names(self) -> "SelectColumns":
        names: Set[str] = set()
        for x in self._all:
            if isinstance(x, _WildcardExpr):
                continue
            if isinstance(x, _NamedColumnExpr) and self._has_wildcard and x.as_name == "":
                raise ValueError(f"with '*', all other columns must have an alias: {self}")
            if x.output_name == "":
                raise ValueError(f"{x} does not have an alias: {self}")
            if x.output_name in names:
                raise ValueError(f"{x} can't be reused in select: {self}")
            names.add(x.output_name)
        return self

    def assert_no_wildcard(self) -> "SelectColumns":
        assert not self._has_wildcard
        return self

    def assert_no_agg(self) -> "SelectColumns":
        assert not self.has_agg
        return self

    @property
    def all_cols(self) -> List[ColumnExpr]:
        return self._all

    @property
    def literals(self) -> List[ColumnExpr]:
        return self._literals

    @property
    def simple_cols(self) -> List[ColumnExpr]:
        return self._cols

    @property
    def non_agg_funcs(self) -> List[ColumnExpr]:
        return self._non_agg_funcs

    @property
    def agg_funcs(self) -> List[ColumnExpr]:
        return self._agg_funcs

    @property
    def group_keys(self) -> List[ColumnExpr]:
        return self._group_keys

    @property
    def has_agg(self) -> bool:
        return len(self._agg_funcs) > 0

    @property
    def has_literals(self) -> bool:
        return len(self._literals) > 0

    @property
    def simple(self) -> bool:
        return len(self._cols) == len(self._all)


_SQL_OPERATORS: Dict[str, str] = {"+": "+", "-": "-", "*": "*", "/": "/", "&": " AND ", "|": " OR ",
                                  "<": "<", ">": ">", "<=": "<=", ">=": ">=", "==": "=", "!=": "!="}


class SQLExpressionGenerator:
    """Expression trees -> the SQL text of fugue/column/sql.py:249-497 (``(is_table, text)`` pieces for
    whole statements, as ``StructuredRawSQL`` wants them)."""

    def __init__(self, enable_cast: bool = True):
        self._enable_cast = enable_cast
        self._func_handler: Dict[str, Callable[[_FuncExpr], Iterable[str]]] = {}

    def add_func_handler(self, name: str, handler: Callable[[_FuncExpr], Iterable[str]]) -> "SQLExpressionGenerator":
        self._func_handler[name] = handler
        return self

    def generate(self, expr: ColumnExpr) -> str:
        return "".join(self._gen(expr, False)).strip()

    def where(self, condition: ColumnExpr, table: str) -> Iterable[Tuple[bool, str]]:
        if is_agg(condition):
            raise ValueError(f"{condition} has aggregation functions")
        yield (False, "SELECT * FROM")
        yield (True, table)
        yield (False, "WHERE " + self.generate(condition.alias("")))

    def select(self, columns: SelectColumns, table: str, where: Optional[ColumnExpr] = None,
               having: Optional[ColumnExpr] = None) -> Iterable[Tuple[bool, str]]:
        columns.assert_all_with_names()
        if where is not None and is_agg(where):
            raise ValueError(f"{where} has aggregation functions")
        where_sql = "" if where is None else "WHERE " + self.generate(where.alias(""))
        having_sql = "" if having is None else "HAVING " + self.generate(having.alias(""))
        distinct = "DISTINCT " if columns.is_distinct else ""
        if not columns.has_agg:
            yield (False, f"SELECT {distinct}{', '.join(self.generate(x) for x in columns.all_cols)} FROM")
            yield (True, table)
            yield (False, where_sql)
            return
        columns.assert_no_wildcard()
        if columns.has_literals:  # literals are attached around the aggregation
            inner = [x for x in columns.all_cols if not isinstance(x, _LiteralColumnExpr)]
            names = [self.generate(x) if isinstance(x, _LiteralColumnExpr) else x.output_name
                     for x in columns.all_cols]
            yield (False, f"SELECT {', '.join(names)} FROM (")
            yield from self.select(SelectColumns(*inner), table, where=where, having=having)
            yield (False, ")")
            return
        yield (False, f"SELECT {distinct}{', '.join(self.generate(x) for x in columns.all_cols)} FROM")
        yield (True, table)
        yield (False, where_sql)
        if len(columns.group_keys) > 0:
            yield (False, "GROUP BY " + ", ".join(self.generate(x) for x in columns.group_keys))
        yield (False, having_sql)

    def correct_select_schema(self, input_schema: Schema, select: SelectColumns,
                              output_schema: Schema) -> Optional[Schema]:
        cols = select.replace_wildcard(input_schema).assert_all_with_names()
        fields = []
        for c in cols.all_cols:
            tp = c.infer_type(input_schema)
            if tp is not None and tp != output_schema[c.output_name].type:
                fields.append(pa.field(c.output_name, tp))
        return Schema(fields) if fields else None

    def type_to_expr(self, data_type: pa.DataType) -> str:
        return type_to_expr(data_type)

    def _gen(self, expr: ColumnExpr, bracket: bool) -> Iterable[str]:
        casting = self._enable_cast and expr.as_type is not None
        if casting:
            yield "CAST("
        if isinstance(expr, _LiteralColumnExpr):
            yield expr.body_str
        elif isinstance(expr, _NamedColumnExpr):
            yield _quote_name(expr.name)
        elif isinstance(expr, _WildcardExpr):
            yield "*"
        elif isinstance(expr, _FuncExpr):
            if expr.func in self._func_handler:
                yield from self._func_handler[expr.func](expr)
            elif isinstance(expr, _UnaryOpExpr):
                inner = "".join(self._gen(expr.col, True))
                yield {"-": "-" + inner, "~": "NOT " + inner, "IS_NULL": inner + " IS NULL",
                       "NOT_NULL": inner + " IS NOT NULL"}[expr.op]
            elif isinstance(expr, _BinaryOpExpr):
                if expr.op not in _SQL_OPERATORS:
                    raise NotImplementedError(expr)
                body = "".join(self._gen(expr.left, True)) + _SQL_OPERATORS[expr.op] + \
                    "".join(self._gen(expr.right, True))
                yield "(" + body + ")" if bracket else body
            else:
                def piece(v: Any) -> str:
                    return "".join(self._gen(v if isinstance(v, ColumnExpr) else lit(v), False))

                parts = [piece(x) for x in expr.args] + [k + "=" + piece(v) for k, v in expr.kwargs.items()]
                yield f"{expr.func}({'DISTINCT ' if expr.is_distinct else ''}{','.join(parts)})"
        if casting:
            yield " AS " + self.type_to_expr(expr.as_type) + ")"
        if expr.as_name != "":
            yield " AS " + _quote_name(expr.as_name)
        elif expr.as_type is not None and expr.name != "":
            yield " AS " + _quote_name(expr.name)
