"""Adapter that plugs the B200 engine into a real Fugue install (``fugue`` + ``triad`` importable).

Import-guarded: ``fugue`` cannot be imported in the build image, so nothing in
the package depends on this module and it is exercised only where Fugue is installed
(INTEGRATION.md).  Registration follows the in-tree backends:
  fugue_duckdb/registry.py:30-75, fugue_dask/registry.py:25-70, fugue_polars/registry.py:19-38,
entry-point group ``fugue.plugins`` (fugue/constants.py:7, setup.py:99-107).

Design: subclass ``NativeExecutionEngine`` for everything outside the hot path (exactly how
``fugue_duckdb`` reuses ``PandasMapEngine``, fugue_duckdb/execution_engine.py:200-201) and
override the facets SURVEY.md section 8 puts on the path: ``map_engine`` (map_dataframe),
``repartition``, ``join``, ``aggregate``, ``select`` (hence ``filter`` / ``assign``) and the SQL facet
(``create_default_sql_engine`` / ``register_sql_engine("b200")``, so FugueSQL SELECTs reach the device),
translating the reference's column expressions node by node into the engine's IR (``translate_expr``).
None of these falls back to the host engine.
"""
from contextlib import contextmanager
from typing import Any, Callable, Dict, Iterator, List, Optional

import pyarrow as pa

import fugue  # noqa: F401  (ImportError here means: no Fugue, no adapter)
from fugue import (ArrowDataFrame as FArrowDataFrame, DataFrame as FDataFrame,
                   LocalDataFrame as FLocalDataFrame, MapEngine as FMapEngine,
                   NativeExecutionEngine, PartitionCursor as FPartitionCursor,
                   PartitionSpec as FPartitionSpec, SQLEngine as FSQLEngine)
from fugue.dataframe.dataframe import LocalBoundedDataFrame as FLocalBoundedDataFrame
from fugue.dev import LocalDataFrameParam, fugue_annotated_param
from fugue.execution.factory import register_execution_engine, register_sql_engine
from fugue.plugins import (as_fugue_dataset, count, get_column_names, get_num_partitions, get_schema,
                           infer_execution_engine, is_bounded, is_df, is_empty, is_local)
from triad import Schema as TSchema

from . import api as _api  # noqa: F401
from . import column as _ir
from .dataframe import ArrowDataFrame as _ArrowDF, B200DataFrame as _B200DF, DataFrame as _DF
from .execution_engine import B200ExecutionEngine as _Engine
from .partition import PartitionSpec as _Spec
from .schema import Schema as _Schema
from .sql import B200SQLEngine as _SQLEngine, StructuredRawSQL as _RawSQL
from .table import B200Table


class FugueB200DataFrame(FDataFrame):
    """``fugue.dataframe.DataFrame`` over a ``B200Table`` (abstract members: dataframe.py:29-299)."""

    def __init__(self, table: B200Table):
        self._table = table
        super().__init__(TSchema(table.schema.pa_schema))

    @property
    def native(self) -> B200Table:
        return self._table

    def native_as_df(self) -> B200Table:
        return self._table

    @property
    def is_local(self) -> bool:
        return False

    @property
    def is_bounded(self) -> bool:
        return True

    @property
    def num_partitions(self) -> int:
        return self._table.num_partitions

    @property
    def empty(self) -> bool:
        return self._table.num_rows == 0

    def count(self) -> int:
        return self._table.num_rows

    def as_arrow(self, type_safe: bool = False) -> pa.Table:
        return self._table.to_arrow()

    def as_pandas(self):
        return self._table.to_pandas()

    def as_local_bounded(self) -> FLocalBoundedDataFrame:
        res = FArrowDataFrame(self.as_arrow())
        if self.has_metadata:
            res.reset_metadata(self.metadata)
        return res

    def peek_array(self) -> List[Any]:
        return _B200DF(self._table).peek_array()

    def as_array(self, columns: Optional[List[str]] = None, type_safe: bool = False) -> List[Any]:
        return self.as_local_bounded().as_array(columns, type_safe=type_safe)

    def as_array_iterable(self, columns: Optional[List[str]] = None, type_safe: bool = False):
        yield from self.as_array(columns, type_safe)

    def head(self, n: int, columns: Optional[List[str]] = None) -> FLocalBoundedDataFrame:
        t = self._table.slice(0, min(n, self._table.num_rows))
        return FArrowDataFrame((t.select(columns) if columns else t).to_arrow())

    def _drop_cols(self, cols: List[str]) -> FDataFrame:
        return FugueB200DataFrame(self._table.select([c for c in self.columns if c not in cols]))

    def _select_cols(self, cols: List[Any]) -> FDataFrame:
        return FugueB200DataFrame(self._table.select(cols))

    def rename(self, columns: Dict[str, str]) -> FDataFrame:
        return FugueB200DataFrame(self._table.rename(columns))

    def alter_columns(self, columns: Any) -> FDataFrame:
        new_schema = self._get_altered_schema(columns)
        if new_schema == self.schema:
            return self
        return FugueB200DataFrame(B200Table.from_arrow(self.as_arrow().cast(new_schema.pa_schema)))


def _to_spec(spec: FPartitionSpec) -> _Spec:
    return _Spec(dict(spec.jsondict))


class FugueB200MapEngine(FMapEngine):
    """``MapEngine.map_dataframe`` (fugue/execution/execution_engine.py:283-315) on the device."""

    @property
    def execution_engine_constraint(self):
        return FugueB200ExecutionEngine

    @property
    def is_distributed(self) -> bool:
        return False

    def map_dataframe(self, df: FDataFrame, map_func: Callable, output_schema: Any,
                      partition_spec: FPartitionSpec, on_init: Optional[Callable] = None,
                      map_func_format_hint: Optional[str] = None) -> FDataFrame:
        eng: "FugueB200ExecutionEngine" = self.execution_engine  # type: ignore
        inner = eng.b200
        edf = eng._to_device(df)
        out_schema = _Schema(TSchema(output_schema).pa_schema)
        fcursor = partition_spec.get_cursor(df.schema, 0)

        def adapt(cursor: Any, part: _DF) -> _DF:
            # hand Fugue's own cursor/dataframe types to Fugue's map_func
            fcursor.set(lambda: cursor.row, cursor.partition_no, cursor.slice_no)
            fdf = FugueB200DataFrame(part.native) if isinstance(part, _B200DF) \
                else FArrowDataFrame(part.as_arrow())
            res = map_func(fcursor, fdf)
            if isinstance(res, FugueB200DataFrame):
                return _B200DF(res.native)
            return _ArrowDF(res.as_arrow())

        finit = None if on_init is None else (lambda no, d: on_init(no, FugueB200DataFrame(d.native)))
        res = inner.map_engine.map_dataframe(edf, adapt, out_schema, _to_spec(partition_spec), finit,
                                             map_func_format_hint)
        return FugueB200DataFrame(res.native)


class FugueB200ExecutionEngine(NativeExecutionEngine):
    def __init__(self, conf: Any = None, **kwargs: Any):
        super().__init__(conf)
        self._b200 = _Engine(dict(self.conf), **kwargs)

    def __repr__(self) -> str:
        return "FugueB200ExecutionEngine"

    @property
    def b200(self) -> _Engine:
        return self._b200

    def create_default_map_engine(self) -> FMapEngine:
        return FugueB200MapEngine(self)

    def _to_device(self, df: Any) -> _B200DF:
        if isinstance(df, FugueB200DataFrame):
            return _B200DF(df.native)
        if isinstance(df, B200Table):
            return _B200DF(df)
        local = super().to_df(df)
        return self._b200.to_df(_ArrowDF(local.as_arrow()))

    def to_df(self, df: Any, schema: Any = None) -> FDataFrame:
        if isinstance(df, FugueB200DataFrame):
            return df
        if isinstance(df, B200Table):
            return FugueB200DataFrame(df)
        return super().to_df(df, schema)  # host data stays on the host until a device op needs it

    def repartition(self, df: FDataFrame, partition_spec: FPartitionSpec) -> FDataFrame:
        if len(partition_spec.partition_by) == 0:
            return df
        return FugueB200DataFrame(self._b200.repartition(self._to_device(df), _to_spec(partition_spec)).native)

    def persist(self, df: FDataFrame, lazy: bool = False, **kwargs: Any) -> FDataFrame:
        return self.to_df(df)

    def create_default_sql_engine(self) -> FSQLEngine:
        """FugueSQL ``SELECT`` statements (``RunSQLSelect`` -> ``SQLEngine.select``,
        fugue/extensions/_builtins/processors.py:148-154) run on the device too."""
        return FugueB200SQLEngine(self)

    # The operators SURVEY.md section 8 puts on the path never fall back to the host engine: what the
    # device path can't do raises NotImplementedError (north_star: no CPU fallback on join / aggregate).
    def join(self, df1: FDataFrame, df2: FDataFrame, how: str, on: Optional[List[str]] = None) -> FDataFrame:
        res = self._b200.join(self._to_device(df1), self._to_device(df2), how, on)
        return FugueB200DataFrame(res.native)

    # ``filter`` and ``assign`` of the base class are written in terms of ``select``
    # (fugue/execution/execution_engine.py:808-887), so these two overrides put all four on the device
    def select(self, df: FDataFrame, cols: Any, where: Any = None, having: Any = None) -> FDataFrame:
        mine = _ir.SelectColumns(*[translate_expr(c) for c in cols.all_cols], arg_distinct=cols.is_distinct)
        res = self._b200.select(self._to_device(df), mine,
                                where=None if where is None else translate_expr(where),
                                having=None if having is None else translate_expr(having))
        return FugueB200DataFrame(res.native)

    def aggregate(self, df: FDataFrame, partition_spec: Optional[FPartitionSpec], agg_cols: List[Any]) -> FDataFrame:
        res = self._b200.aggregate(self._to_device(df),
                                   None if partition_spec is None else _to_spec(partition_spec),
                                   [translate_expr(c) for c in agg_cols])
        return FugueB200DataFrame(res.native)


class FugueB200SQLEngine(FSQLEngine):
    """``SQLEngine.select`` (fugue/execution/execution_engine.py:209-238) for the B200 engine: the
    statement's table references are bound to device tables and the text goes to the engine's own
    SELECT parser (fugue_b200/sql.py: single-table SELECT / WHERE / GROUP BY / HAVING / ORDER BY /
    LIMIT and two-table equi-joins -> device select / group-by / join kernels).  Statements outside
    that grammar raise NotImplementedError; there is no host SQL engine behind it."""

    @property
    def execution_engine_constraint(self):
        return FugueB200ExecutionEngine

    @property
    def is_distributed(self) -> bool:
        return False

    @property
    def dialect(self) -> Optional[str]:
        return None  # no sqlglot transpile (fugue/collections/sql.py:99-103)

    def select(self, dfs: Any, statement: Any) -> FDataFrame:
        eng: "FugueB200ExecutionEngine" = self.execution_engine  # type: ignore
        named, text = self.encode(dfs, statement)
        tables = {k: eng._to_device(v) for k, v in named.items()}
        res = _SQLEngine(eng.b200).select(tables, _RawSQL([(False, text)]))
        return FugueB200DataFrame(res.native if isinstance(res, _B200DF) else eng._to_device(
            FArrowDataFrame(res.as_arrow())).native)


# ---- the reference's expression trees -> the engine's IR ---------------------------------------------
_UNARY_BUILDERS = {"-": lambda x: -x, "~": lambda x: ~x, "IS_NULL": lambda x: x.is_null(),
                   "NOT_NULL": lambda x: x.not_null()}


def translate_expr(e: Any) -> Any:
    """``fugue.column`` expression tree (fugue/column/expressions.py, functions.py) -> ``fugue_b200.column``
    IR, node by node.  The reference encodes the node kind in its class; here it becomes ``Kind``."""
    from fugue.column import expressions as fe
    from fugue.column import functions as ff

    if not isinstance(e, fe.ColumnExpr):
        return e  # plain python value used as a function argument
    if isinstance(e, fe._WildcardExpr):
        return _ir.all_cols()
    if isinstance(e, fe._NamedColumnExpr):
        out: Any = _ir.col(e.name)
    elif isinstance(e, fe._LiteralColumnExpr):
        out = _ir.lit(e.value)
    elif isinstance(e, ff._UnaryAggFuncExpr):
        out = _ir.agg(e.func, translate_expr(e.args[0]), arg_distinct=e.is_distinct)
    elif isinstance(e, fe._UnaryOpExpr):
        if e.op not in _UNARY_BUILDERS:
            raise NotImplementedError(f"unary operator {e.op}")
        out = _UNARY_BUILDERS[e.op](translate_expr(e.col))
    elif isinstance(e, fe._BinaryOpExpr):
        out = _ir.binary(e.op, translate_expr(e.left), translate_expr(e.right))
    elif isinstance(e, fe._FuncExpr):
        out = _ir.function(e.func, *[translate_expr(x) for x in e.args], arg_distinct=e.is_distinct,
                           **{k: translate_expr(v) for k, v in e.kwargs.items()})
    else:
        raise NotImplementedError(f"can't translate {type(e).__name__}")
    if e.as_type is not None:
        out = out.cast(e.as_type)
    return out.alias(e.as_name) if e.as_name != "" else out


# ---- registration (patterns: fugue_duckdb/registry.py:30-75, fugue_dask/registry.py:25-70,
#      fugue/dataframe/arrow_dataframe.py:263-331) --------------------------------------------------------
@infer_execution_engine.candidate(
    lambda objs: any(isinstance(o, (B200Table, FugueB200DataFrame)) for o in objs))
def _infer_b200(objs: Any) -> Any:
    return "b200"


@as_fugue_dataset.candidate(lambda df, **kwargs: isinstance(df, B200Table))
def _b200_as_fugue_df(df: B200Table, **kwargs: Any) -> FugueB200DataFrame:
    return FugueB200DataFrame(df)


@fugue_annotated_param(B200Table)
class _B200TableParam(LocalDataFrameParam):
    """Functions typed on ``B200Table`` are called once per device table (format hint "b200")."""

    def to_input_data(self, df: FDataFrame, ctx: Any) -> Any:
        assert isinstance(df, FugueB200DataFrame), "B200Table functions run on the b200 engine"
        return df.native

    def to_output_df(self, output: Any, schema: Any, ctx: Any) -> FDataFrame:
        assert isinstance(output, B200Table)
        return FugueB200DataFrame(output)

    def count(self, df: Any) -> int:
        return df.num_rows

    def format_hint(self) -> Optional[str]:
        return "b200"


def _is_table(df: Any, *args: Any, **kwargs: Any) -> bool:
    return isinstance(df, B200Table)


@is_df.candidate(_is_table)
def _b200_is_df(df: B200Table) -> bool:
    return True


@count.candidate(_is_table)
def _b200_count(df: B200Table) -> int:
    return df.num_rows


@is_bounded.candidate(_is_table)
def _b200_is_bounded(df: B200Table) -> bool:
    return True


@is_empty.candidate(_is_table)
def _b200_is_empty(df: B200Table) -> bool:
    return df.num_rows == 0


@is_local.candidate(_is_table)
def _b200_is_local(df: B200Table) -> bool:
    return False  # lives in HBM


@get_num_partitions.candidate(_is_table)
def _b200_num_partitions(df: B200Table) -> int:
    return df.num_partitions


@get_column_names.candidate(_is_table)
def _b200_column_names(df: B200Table) -> List[Any]:
    return list(df.schema.names)


@get_schema.candidate(_is_table)
def _b200_schema(df: B200Table) -> TSchema:
    return TSchema(df.schema.pa_schema)


def _make_test_backend() -> Any:
    """``@ft.fugue_test_backend`` class (fugue/test/plugins.py:99-136, 226-312): lets the reference's own
    conformance suites (``fugue_test/execution_suite.py``, ``builtin_suite.py``) run against this engine with
    ``@ft.fugue_test_suite("b200", mark_test=True)``.  ``fugue.test`` needs pytest; skipped where absent."""
    try:
        import fugue.test as ft
    except Exception:  # pragma: no cover
        return None

    @ft.fugue_test_backend
    class B200TestBackend(ft.FugueTestBackend):
        name = "b200"
        default_fugue_conf: Dict[str, Any] = {"fugue.b200.default.partitions": 16}

        @classmethod
        @contextmanager
        def session_context(cls, session_conf: Dict[str, Any]) -> Iterator[Any]:
            yield "b200"  # the engine name is the session object: make_execution_engine("b200", conf)

    return B200TestBackend


def register() -> None:
    register_execution_engine("b200", lambda conf, **kwargs: FugueB200ExecutionEngine(conf, **kwargs),
                              on_dup="ignore")
    register_sql_engine("b200", lambda engine: FugueB200SQLEngine(engine), on_dup="ignore")


register()
B200TestBackend = _make_test_backend()
