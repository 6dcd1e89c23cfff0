"""Adapter that plugs the B200 engine into a real Fugue install (``fugue`` + ``triad`` importable).

Import-guarded: ``fugue`` cannot be imported in the build image, so nothing in
the package depends on this module and it is exercised only where Fugue is installed
(INTEGRATION.md).  Registration follows the in-tree backends:
  fugue_duckdb/registry.py:30-75, fugue_dask/registry.py:25-70, fugue_polars/registry.py:19-38,
entry-point group ``fugue.plugins`` (fugue/constants.py:7, setup.py:99-107).

Design: subclass ``NativeExecutionEngine`` for everything outside the hot path (exactly how
``fugue_duckdb`` reuses ``PandasMapEngine``, fugue_duckdb/execution_engine.py:200-201) and
override the facets SURVEY.md section 8 puts on the path: ``map_engine`` (map_dataframe),
``repartition``, ``join``, ``aggregate`` and ``select`` (hence ``filter`` / ``assign``), translating the
reference's column expressions node by node into ``fugue_b200.column``.
"""
from typing import Any, Callable, Dict, List, Optional

import pyarrow as pa

import fugue  # noqa: F401  (ImportError here means: no Fugue, no adapter)
from fugue import (ArrowDataFrame as FArrowDataFrame, DataFrame as FDataFrame,
                   LocalDataFrame as FLocalDataFrame, MapEngine as FMapEngine,
                   NativeExecutionEngine, PartitionCursor as FPartitionCursor,
                   PartitionSpec as FPartitionSpec)
from fugue.dataframe.dataframe import LocalBoundedDataFrame as FLocalBoundedDataFrame
from fugue.dev import LocalDataFrameParam, fugue_annotated_param
from fugue.execution.factory import register_execution_engine
from fugue.plugins import as_fugue_dataset, infer_execution_engine
from triad import Schema as TSchema

from . import api as _api
from .dataframe import ArrowDataFrame as _ArrowDF, B200DataFrame as _B200DF, DataFrame as _DF
from .execution_engine import B200ExecutionEngine as _Engine
from .partition import PartitionSpec as _Spec
from .schema import Schema as _Schema
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

    def join(self, df1: FDataFrame, df2: FDataFrame, how: str, on: Optional[List[str]] = None) -> FDataFrame:
        try:  # every join type runs on the device (fugue_b200/join.py); exotic column types fall back
            res = self._b200.join(self._to_device(df1), self._to_device(df2), how, on)
            return FugueB200DataFrame(res.native)
        except NotImplementedError:
            return super().join(self.to_df(df1).as_local(), self.to_df(df2).as_local(), how, on)

    # ``filter`` and ``assign`` of the base class are written in terms of ``select``
    # (fugue/execution/execution_engine.py:808-887), so these two overrides put all four on the device
    def select(self, df: FDataFrame, cols: Any, where: Any = None, having: Any = None) -> FDataFrame:
        from .column import SelectColumns as _SelectColumns

        try:
            mine = _SelectColumns(*[_translate_expr(c) for c in cols.all_cols], arg_distinct=cols.is_distinct)
            res = self._b200.select(self._to_device(df), mine,
                                    where=None if where is None else _translate_expr(where),
                                    having=None if having is None else _translate_expr(having))
            return FugueB200DataFrame(res.native)
        except NotImplementedError:
            return super().select(self.to_df(df).as_local(), cols, where=where, having=having)

    def aggregate(self, df: FDataFrame, partition_spec: Optional[FPartitionSpec], agg_cols: List[Any]) -> FDataFrame:
        try:
            res = self._b200.aggregate(self._to_device(df),
                                       None if partition_spec is None else _to_spec(partition_spec),
                                       [_translate_expr(c) for c in agg_cols])
            return FugueB200DataFrame(res.native)
        except NotImplementedError:
            return super().aggregate(self.to_df(df).as_local(), partition_spec, agg_cols)


def _translate_expr(e: Any) -> Any:
    """``fugue.column`` expression tree -> the same tree in ``fugue_b200.column`` (the two DSLs have the
    same node kinds: named / wildcard / literal / function / unary / binary / aggregation)."""
    from fugue.column import expressions as fe
    from fugue.column import functions as ff

    from . import column as bc

    if not isinstance(e, fe.ColumnExpr):
        return e
    if isinstance(e, fe._WildcardExpr):
        return bc.all_cols()
    if isinstance(e, fe._NamedColumnExpr):
        res: Any = bc.col(e.name)
    elif isinstance(e, fe._LiteralColumnExpr):
        res = bc.lit(e.value)
    elif isinstance(e, ff._UnaryAggFuncExpr):
        same = isinstance(e, ff._SameTypeUnaryAggFuncExpr)
        cls = bc._SameTypeAggFuncExpr if same else bc.AggFuncExpr
        res = cls(e.func, _translate_expr(e.args[0]), arg_distinct=e.is_distinct)
    elif isinstance(e, fe._UnaryOpExpr):
        arg = _translate_expr(e.col)
        res = {"-": lambda: -arg, "~": lambda: ~arg, "IS_NULL": arg.is_null, "NOT_NULL": arg.not_null}[e.op]()
    elif isinstance(e, fe._BinaryOpExpr):
        kind = bc._BoolBinaryOpExpr if isinstance(e, fe._BoolBinaryOpExpr) else bc._BinaryOpExpr
        res = kind(e.op, _translate_expr(e.left), _translate_expr(e.right))
    elif isinstance(e, fe._FuncExpr):
        res = bc.function(e.func, *[_translate_expr(a) for a in e.args], arg_distinct=e.is_distinct,
                          **{k: _translate_expr(v) for k, v in e.kwargs.items()})
    else:
        raise NotImplementedError(f"can't translate {type(e).__name__}")
    if e.as_type is not None:
        res = res.cast(e.as_type)
    return res.alias(e.as_name) if e.as_name != "" else res


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


def register() -> None:
    register_execution_engine("b200", lambda conf, **kwargs: FugueB200ExecutionEngine(conf, **kwargs),
                              on_dup="ignore")


register()
