"""Compile column expressions (``fugue_b200.column``) into programs for the device evaluator.

``project`` / ``filter_table`` are what ``B200ExecutionEngine.select / filter / assign`` run on: the
reference generates SQL text for these trees and lets qpd/pandas evaluate it operator by operator
(fugue/execution/execution_engine.py:736-887, fugue/column/sql.py:275-347); here a whole SELECT list is
one register-machine program and one pass over HBM (``fb_eval_expr``, fugue_b200/csrc/fb_expr.cu).

Typing rules (the reference leaves them to the SQL engine; these are the pandas/qpd ones):
``+ - *`` on integers -> int64, with any float -> float64; ``/`` -> float64 (true division);
comparisons / ``& | ~`` / ``IS NULL`` -> bool with SQL three-valued logic; an explicit ``cast`` wins.
String columns are dictionary encoded: they can be passed through, tested for NULL and compared
(``==`` / ``!=``) with a string literal; ``cast(str)`` of a numeric result builds a dictionary.
"""
import struct
from typing import Any, Dict, List, Optional, Sequence, Tuple

import pyarrow as pa
import torch

from . import kernels as K
from .column import (AggFuncExpr, ColumnExpr, _BinaryOpExpr, _FuncExpr, _LiteralColumnExpr, _NamedColumnExpr,
                     _UnaryOpExpr, _WildcardExpr)
from .schema import Schema
from .table import B200Table, _storage_dtype


def _is_str(tp: Optional[pa.DataType]) -> bool:
    return tp is not None and (pa.types.is_string(tp) or pa.types.is_large_string(tp))


def _cls_of(tp: pa.DataType) -> str:
    if pa.types.is_boolean(tp):
        return "b"
    if pa.types.is_floating(tp):
        return "f"
    if _is_str(tp):
        return "s"
    return "i"  # integers, dates, timestamps: int64 arithmetic


def _f64_bits(v: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(v)))[0]


class _OutOfResources(Exception):
    pass


class _Program:
    """One ``fb_eval_expr`` launch: instructions, referenced columns, pending outputs."""

    def __init__(self, table: B200Table):
        self.t = table
        self.ins: List[Tuple[int, int, int, int, int]] = []
        self.cols: List[int] = []          # table column indices, in load order
        self.free = list(range(K.EXPR_NREGS - 1, -1, -1))
        self.outs: List[Tuple[int, torch.dtype, bool]] = []  # (reg, dtype, want_valid)

    # -- resources
    def alloc(self) -> int:
        if not self.free:
            raise _OutOfResources("registers")
        return self.free.pop()

    def release(self, r: int) -> None:
        self.free.append(r)

    def emit(self, op: int, dst: int, a: int = 0, b: int = 0, imm: int = 0) -> None:
        if len(self.ins) >= K.EXPR_MAX_INS:
            raise _OutOfResources("instructions")
        self.ins.append((op, dst, a, b, imm))

    def col_slot(self, ci: int) -> int:
        if ci in self.cols:
            return self.cols.index(ci)
        if len(self.cols) >= K.EXPR_MAX_COLS:
            raise _OutOfResources("columns")
        self.cols.append(ci)
        return len(self.cols) - 1

    # -- compilation: returns (register, class, nullable)
    def compile(self, e: ColumnExpr, top: bool = False) -> Tuple[int, str, bool]:
        r, cls, nullable = self._node(e)
        if e.as_type is not None:
            if _is_str(e.as_type) and not top and cls != "s":
                raise NotImplementedError(f"cast to str inside an expression (only on a whole column): {e}")
            r, cls = self._cast(r, cls, _cls_of(e.as_type), e)
        return r, cls, nullable

    def _cast(self, r: int, cls: str, want: str, e: ColumnExpr) -> Tuple[int, str]:
        if want == "s":
            return r, cls  # applied to the stored column (top level only, checked by the caller)
        if cls == "s":
            raise NotImplementedError(f"cast of a string expression to {e.as_type}: {e}")
        if want == cls:
            return r, cls
        if want == "f":
            self.emit(K.X_I2F, r, r)
        elif want == "i":
            if cls == "f":
                self.emit(K.X_F2I, r, r)
        elif want == "b":
            self.emit(K.X_TOBOOL_F if cls == "f" else K.X_TOBOOL_I, r, r)
        return r, want

    def _to_float(self, r: int, cls: str) -> None:
        if cls != "f":
            self.emit(K.X_I2F, r, r)

    def _to_bool(self, r: int, cls: str) -> None:
        if cls == "f":
            self.emit(K.X_TOBOOL_F, r, r)
        elif cls == "i":
            self.emit(K.X_TOBOOL_I, r, r)

    def _node(self, e: ColumnExpr) -> Tuple[int, str, bool]:  # noqa: C901
        t = self.t
        if isinstance(e, _NamedColumnExpr):
            if e.name not in t.schema:
                raise KeyError(f"column {e.name} is not in {t.schema}")
            ci = t.schema.index_of_key(e.name)
            cls = _cls_of(t.schema.types[ci])
            r = self.alloc()
            self.emit(K.X_LOAD, r, self.col_slot(ci))
            return r, cls, t.valid[ci] is not None
        if isinstance(e, _LiteralColumnExpr):
            v = e.value
            r = self.alloc()
            if v is None:
                self.emit(K.X_NULL, r)
                return r, "n", True
            if isinstance(v, bool):
                self.emit(K.X_LIT, r, imm=int(v))
                return r, "b", False
            if isinstance(v, int):
                self.emit(K.X_LIT, r, imm=v & ((1 << 64) - 1))
                return r, "i", False
            if isinstance(v, float):
                self.emit(K.X_LIT, r, imm=_f64_bits(v))
                return r, "f", False
            raise NotImplementedError(f"string literal {e} outside a comparison with a string column")
        if isinstance(e, _WildcardExpr):
            raise ValueError("'*' can't be evaluated as a value")
        if isinstance(e, AggFuncExpr):
            raise ValueError(f"aggregation {e} in a row-wise expression")
        if isinstance(e, _UnaryOpExpr):
            if e.op in ("IS_NULL", "NOT_NULL"):
                r, _, _ = self._string_or_value(e.col)
                self.emit(K.X_IS_NULL if e.op == "IS_NULL" else K.X_NOT_NULL, r, r)
                return r, "b", False
            r, cls, nullable = self.compile(e.col)
            if cls == "s":
                raise NotImplementedError(f"{e.op} on a string expression: {e}")
            if e.op == "-":
                if cls == "n":
                    return r, cls, True
                self.emit(K.X_NEG_F if cls == "f" else K.X_NEG_I, r, r)
                return r, ("i" if cls == "b" else cls), nullable
            if e.op == "~":
                self._to_bool(r, cls)
                self.emit(K.X_NOT, r, r)
                return r, "b", nullable
            raise NotImplementedError(f"unary operator {e.op}")
        if isinstance(e, _BinaryOpExpr):
            return self._binary(e)
        if isinstance(e, _FuncExpr):
            if e.func.upper() == "COALESCE":
                return self._coalesce(e)
            raise NotImplementedError(f"function {e.func} has no device implementation")
        raise NotImplementedError(f"can't evaluate {e!r}")

    def _string_or_value(self, e: ColumnExpr) -> Tuple[int, str, bool]:
        """Operand of IS NULL / NOT NULL: string columns are fine here (only validity is read)."""
        return self.compile(e)

    def _binary(self, e: _BinaryOpExpr) -> Tuple[int, str, bool]:  # noqa: C901
        op = e.op
        str_cmp = self._string_compare(e)
        if str_cmp is not None:
            return str_cmp
        ra, ca, na = self.compile(e.left)
        rb, cb, nb = self.compile(e.right)
        if "s" in (ca, cb):
            raise NotImplementedError(f"operator {op} on string operands: {e}")
        nullable = na or nb
        if op in ("&", "|"):
            if ca != "n":
                self._to_bool(ra, ca)
            if cb != "n":
                self._to_bool(rb, cb)
            self.emit(K.X_AND if op == "&" else K.X_OR, ra, ra, rb)
            self.release(rb)
            return ra, "b", nullable
        is_f = "f" in (ca, cb) or op == "/"
        if is_f:
            if ca != "n":
                self._to_float(ra, ca)
            if cb != "n":
                self._to_float(rb, cb)
        table = {
            "+": (K.X_ADD_I, K.X_ADD_F, False), "-": (K.X_SUB_I, K.X_SUB_F, False),
            "*": (K.X_MUL_I, K.X_MUL_F, False), "/": (None, K.X_DIV_F, False),
            "<": (K.X_LT_I, K.X_LT_F, False), "<=": (K.X_LE_I, K.X_LE_F, False),
            ">": (K.X_LT_I, K.X_LT_F, True), ">=": (K.X_LE_I, K.X_LE_F, True),
            "==": (K.X_EQ_I, K.X_EQ_F, False), "!=": (K.X_NE_I, K.X_NE_F, False),
        }
        if op not in table:
            raise NotImplementedError(f"operator {op}")
        oi, of, swap = table[op]
        code = of if is_f else oi
        if swap:
            self.emit(code, ra, rb, ra)
        else:
            self.emit(code, ra, ra, rb)
        self.release(rb)
        if op in ("+", "-", "*", "/"):
            return ra, ("f" if is_f else "i"), nullable
        return ra, "b", nullable

    def _string_compare(self, e: _BinaryOpExpr) -> Optional[Tuple[int, str, bool]]:
        """``strcol == 'lit'`` / ``!=``: compare dictionary codes."""
        t = self.t
        sides = [e.left, e.right]
        named = [isinstance(s, _NamedColumnExpr) and s.as_type is None and s.name in t.schema
                 and _is_str(t.schema[s.name].type) for s in sides]
        lits = [isinstance(s, _LiteralColumnExpr) and isinstance(s.value, str) for s in sides]
        if not (any(named) or any(lits)):
            return None
        if e.op not in ("==", "!="):
            raise NotImplementedError(f"operator {e.op} on strings (only == and != run on the device): {e}")
        if named[0] and lits[1]:
            c, lit_ = sides[0], sides[1]
        elif named[1] and lits[0]:
            c, lit_ = sides[1], sides[0]
        else:
            raise NotImplementedError(f"string comparison needs a string column and a literal: {e}")
        ci = t.schema.index_of_key(c.name)
        d = t.dictionaries[c.name]
        code = d.index(lit_.value).as_py() if len(d) > 0 else -1  # -1: value not in the dictionary
        r = self.alloc()
        self.emit(K.X_LOAD, r, self.col_slot(ci))
        rl = self.alloc()
        self.emit(K.X_LIT, rl, imm=(code if code is not None else -1) & ((1 << 64) - 1))
        self.emit(K.X_EQ_I if e.op == "==" else K.X_NE_I, r, r, rl)
        self.release(rl)
        return r, "b", t.valid[ci] is not None

    def _coalesce(self, e: _FuncExpr) -> Tuple[int, str, bool]:
        args = [a if isinstance(a, ColumnExpr) else _LiteralColumnExpr(a) for a in e.args]
        if len(args) == 0:
            raise ValueError("COALESCE needs arguments")
        # the result class: float if any argument is float, else int / bool
        probe = [self._static_cls(a) for a in args]
        if "s" in probe:
            raise NotImplementedError(f"COALESCE on strings: {e}")
        want = "f" if "f" in probe else ("b" if all(p in ("b", "n") for p in probe) and "b" in probe else "i")
        acc: Optional[int] = None
        nullable = True
        for a in args:
            r, cls, n = self.compile(a)
            if cls != "n" and want == "f":
                self._to_float(r, cls)
            if acc is None:
                acc, nullable = r, n
            else:
                self.emit(K.X_COALESCE, acc, acc, r)
                self.release(r)
                nullable = nullable and n
        assert acc is not None
        return acc, want, nullable

    def _static_cls(self, e: ColumnExpr) -> str:
        """Class an expression will evaluate to (without emitting code)."""
        if e.as_type is not None:
            return _cls_of(e.as_type)
        if isinstance(e, _NamedColumnExpr):
            return _cls_of(self.t.schema[e.name].type)
        if isinstance(e, _LiteralColumnExpr):
            v = e.value
            return "n" if v is None else "b" if isinstance(v, bool) else "i" if isinstance(v, int) else \
                "f" if isinstance(v, float) else "s"
        if isinstance(e, _UnaryOpExpr):
            if e.op in ("IS_NULL", "NOT_NULL", "~"):
                return "b"
            c = self._static_cls(e.col)
            return "i" if c == "b" else c
        if isinstance(e, _BinaryOpExpr):
            if e.op in ("+", "-", "*", "/"):
                cs = (self._static_cls(e.left), self._static_cls(e.right))
                return "f" if (e.op == "/" or "f" in cs) else "i"
            return "b"
        if isinstance(e, _FuncExpr) and e.func.upper() == "COALESCE":
            cs = [self._static_cls(a if isinstance(a, ColumnExpr) else _LiteralColumnExpr(a)) for a in e.args]
            return "f" if "f" in cs else ("b" if "b" in cs and all(c in ("b", "n") for c in cs) else "i")
        return "i"

    def run(self) -> Tuple[List[torch.Tensor], List[Optional[torch.Tensor]]]:
        t = self.t
        return K.eval_expr(t.num_rows, t.device, [t.columns[i] for i in self.cols],
                           [t.valid[i] for i in self.cols], self.ins, [o[0] for o in self.outs],
                           [o[1] for o in self.outs], [o[2] for o in self.outs])


def _default_type(cls: str, e: ColumnExpr, schema: Schema) -> pa.DataType:
    tp = e.infer_type(schema)
    if tp is not None:
        return tp
    return {"i": pa.int64(), "f": pa.float64(), "b": pa.bool_()}.get(cls, pa.int64())


def _format_values(vals: torch.Tensor, tp_from: str) -> List[str]:
    host = vals.cpu().tolist()
    if tp_from == "b":
        return ["true" if v else "false" for v in host]
    return [str(v) for v in host]


def _to_string_column(col: torch.Tensor, valid: Optional[torch.Tensor], cls: str
                      ) -> Tuple[torch.Tensor, pa.Array]:
    """``CAST(x AS str)``: dictionary = the distinct values, formatted on the host."""
    if col.numel() == 0:
        return torch.empty(0, dtype=torch.int32, device=col.device), pa.array([], type=pa.string())
    src = col if valid is None else torch.where(valid.bool(), col, torch.zeros_like(col))
    uniq, inv = torch.unique(src, return_inverse=True)
    return inv.to(torch.int32).contiguous(), pa.array(_format_values(uniq, cls), type=pa.string())


def project(t: B200Table, exprs: Sequence[ColumnExpr]) -> B200Table:
    """Evaluate a SELECT list (no aggregations, wildcards already expanded, every column named)."""
    n, dev = t.num_rows, t.device
    names = [e.output_name for e in exprs]
    out_cols: List[Any] = [None] * len(exprs)
    out_valid: List[Any] = [None] * len(exprs)
    out_types: List[Any] = [None] * len(exprs)
    dicts: Dict[str, pa.Array] = {}
    pending: List[Tuple[int, ColumnExpr]] = []
    for i, e in enumerate(exprs):
        if isinstance(e, _NamedColumnExpr):
            if e.name not in t.schema:
                raise KeyError(f"column {e.name} is not in {t.schema}")
            ci = t.schema.index_of_key(e.name)
            tp = t.schema.types[ci]
            if e.as_type is None or e.as_type == tp:  # pass through, zero copy
                out_cols[i], out_valid[i], out_types[i] = t.columns[ci], t.valid[ci], tp
                if e.name in t.dictionaries:
                    dicts[names[i]] = t.dictionaries[e.name]
                continue
            if _is_str(tp):
                raise NotImplementedError(f"cast of string column {e.name} to {e.as_type}")
        if isinstance(e, _LiteralColumnExpr) and (isinstance(e.value, str) or e.value is None):
            tp = e.as_type or (pa.string() if isinstance(e.value, str) else None)
            if tp is None:
                raise NotImplementedError(f"NULL literal {e} needs a cast to know its type")
            out_types[i] = tp
            out_cols[i] = torch.zeros(n, dtype=_storage_dtype(tp), device=dev)
            if e.value is None:
                out_valid[i] = torch.zeros(n, dtype=torch.uint8, device=dev)
                if _is_str(tp):
                    dicts[names[i]] = pa.array([], type=pa.string())
            else:
                if not _is_str(tp):
                    raise NotImplementedError(f"cast of string literal {e} to {tp}")
                dicts[names[i]] = pa.array([e.value], type=pa.string())
            continue
        pending.append((i, e))
    # ---- computed columns: as many as fit into one launch at a time
    k = 0
    while k < len(pending):
        prog = _Program(t)
        batch: List[Tuple[int, ColumnExpr, str, bool]] = []
        while k < len(pending) and len(batch) < K.EXPR_MAX_OUTS:
            i, e = pending[k]
            mark = (len(prog.ins), list(prog.cols), list(prog.free))
            try:
                r, cls, nullable = prog.compile(e, top=True)
            except _OutOfResources as ex:
                if not batch:
                    raise NotImplementedError(f"expression too large for one device program ({ex}): {e}")
                del prog.ins[mark[0]:]
                prog.cols, prog.free = mark[1], mark[2]
                break
            if cls == "n":  # a bare NULL-valued expression
                cls = _cls_of(e.as_type) if e.as_type is not None else "i"
            tp = e.as_type if e.as_type is not None else _default_type(cls, e, t.schema)
            store_tp = {"i": pa.int64(), "f": pa.float64(), "b": pa.bool_()}[cls] if _is_str(tp) else tp
            if _cls_of(store_tp) != cls and not _is_str(tp):
                # an inferred type of another class (can't happen for casts: _cast converted already)
                store_tp = {"i": pa.int64(), "f": pa.float64(), "b": pa.bool_()}[cls]
                tp = store_tp
            prog.outs.append((r, _storage_dtype(store_tp), nullable))
            out_types[i] = tp
            batch.append((i, e, cls, nullable))
            k += 1
        cols, valids = prog.run()
        for (i, e, cls, nullable), c, v in zip(batch, cols, valids):
            if _is_str(out_types[i]):
                c, dicts[names[i]] = _to_string_column(c, v, cls)
            out_cols[i], out_valid[i] = c, v
    schema = Schema([pa.field(nm, tp) for nm, tp in zip(names, out_types)])
    return B200Table(schema, out_cols, out_valid, dicts)


def predicate_mask(t: B200Table, condition: ColumnExpr) -> torch.Tensor:
    """uint8 mask: 1 where ``condition`` is TRUE (NULL counts as FALSE, like SQL WHERE)."""
    prog = _Program(t)
    try:
        r, cls, _ = prog.compile(condition.alias("") if condition.as_name else condition)
    except _OutOfResources as ex:
        raise NotImplementedError(f"condition too large for one device program ({ex}): {condition}")
    if cls == "s":
        raise ValueError(f"{condition} is not a boolean expression")
    if cls == "n":
        return torch.zeros(t.num_rows, dtype=torch.uint8, device=t.device)
    prog._to_bool(r, cls)
    prog.outs.append((r, torch.uint8, False))  # the store writes 0 for NULL
    cols, _ = prog.run()
    return cols[0]


def filter_table(t: B200Table, condition: ColumnExpr) -> B200Table:
    """``SELECT * FROM t WHERE condition`` (rows keep their order)."""
    if t.num_rows == 0:
        return t
    idx = K.compact_indices(predicate_mask(t, condition))
    if int(idx.shape[0]) == t.num_rows:
        return B200Table(t.schema, t.columns, t.valid, t.dictionaries)
    cols, valid = K.gather_rows(t.columns, t.valid, idx, False)
    return B200Table(t.schema, cols, valid, t.dictionaries)


def rewrite(e: Any, mapper: Any) -> Any:
    """Copy of the tree with ``mapper(node)`` applied top-down (a non-None result replaces the node,
    keeping the node's alias and cast)."""
    if not isinstance(e, ColumnExpr):
        return e
    rep = mapper(e)
    if rep is not None:
        if e.as_type is not None:
            rep = rep.cast(e.as_type)
        return rep.alias(e.as_name) if e.as_name != "" else rep
    if isinstance(e, _FuncExpr):
        args = [rewrite(a, mapper) for a in e.args]
        kwargs = {k: rewrite(v, mapper) for k, v in e.kwargs.items()}
        if isinstance(e, AggFuncExpr):
            res: ColumnExpr = type(e)(e.func, args[0], arg_distinct=e.is_distinct)
        elif isinstance(e, _UnaryOpExpr):
            res = type(e)(e.op, args[0])
        elif isinstance(e, _BinaryOpExpr):
            res = type(e)(e.op, args[0], args[1])
        else:
            res = _FuncExpr(e.func, *args, arg_distinct=e.is_distinct, **kwargs)
        if e.as_type is not None:
            res = res.cast(e.as_type)
        return res.alias(e.as_name) if e.as_name != "" else res
    return e


def find_aggs(e: Any, out: List[AggFuncExpr]) -> None:
    if isinstance(e, AggFuncExpr):
        out.append(e)
    elif isinstance(e, _FuncExpr):
        for a in e.args:
            find_aggs(a, out)
        for a in e.kwargs.values():
            find_aggs(a, out)
