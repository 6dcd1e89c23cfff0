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
from .column import ColumnExpr, Kind, lit as _lit
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
    """One ``fb_eval_expr`` launch.  Code generation for the accumulator machine: ``compile`` leaves the
    value of an expression in the accumulator; the right-hand side of an operator is used in place
    when it is a leaf (column / literal), otherwise the left value waits in a temporary."""

    _REVERSE = {"+": "+", "*": "*", "-": "r-", "/": "r/", "<": ">", "<=": ">=", ">": "<", ">=": "<=",
                "==": "==", "!=": "!=", "&": "&", "|": "|"}
    _OPS = {  # (int opcode, float opcode)
        "+": (K.X_ADD_I, K.X_ADD_F), "-": (K.X_SUB_I, K.X_SUB_F), "r-": (K.X_RSUB_I, K.X_RSUB_F),
        "*": (K.X_MUL_I, K.X_MUL_F), "/": (None, K.X_DIV_F), "r/": (None, K.X_RDIV_F),
        "<": (K.X_LT_I, K.X_LT_F), "<=": (K.X_LE_I, K.X_LE_F), ">": (K.X_GT_I, K.X_GT_F),
        ">=": (K.X_GE_I, K.X_GE_F), "==": (K.X_EQ_I, K.X_EQ_F), "!=": (K.X_NE_I, K.X_NE_F),
    }

    def __init__(self, table: B200Table):
        self.t = table
        self.ins: List[Tuple[int, int, int, int, int]] = []
        self.cols: List[int] = []          # table column indices, in load order
        self.free = list(range(K.EXPR_NREGS - 1, -1, -1))
        self.outs: List[Tuple[torch.dtype, bool]] = []  # (dtype, want_valid) per FB_X_OUT

    # -- resources
    def alloc(self) -> int:
        if not self.free:
            raise _OutOfResources("temporaries")
        return self.free.pop()

    def release(self, r: int) -> None:
        self.free.append(r)

    def emit(self, op: int, kind: int = K.XK_NONE, b: int = 0, flags: int = 0, imm: int = 0) -> None:
        if len(self.ins) >= K.EXPR_MAX_INS:
            raise _OutOfResources("instructions")
        self.ins.append((op, kind, b, flags, imm))

    def col_slot(self, ci: int) -> int:
        if ci in self.cols:
            return self.cols.index(ci)
        if len(self.cols) >= K.EXPR_MAX_COLS:
            raise _OutOfResources("columns")
        self.cols.append(ci)
        return len(self.cols) - 1

    def output(self, dtype: torch.dtype, want_valid: bool) -> None:
        if len(self.outs) >= K.EXPR_MAX_OUTS:
            raise _OutOfResources("outputs")
        self.emit(K.X_OUT, K.XK_NONE, len(self.outs))
        self.outs.append((dtype, want_valid))

    def mark(self) -> Any:
        return (len(self.ins), list(self.cols), list(self.free), len(self.outs))

    def rollback(self, m: Any) -> None:
        del self.ins[m[0]:]
        self.cols, self.free = m[1], m[2]
        del self.outs[m[3]:]

    # -- leaves: usable directly as operand B.  Returns (kind, b, imm, cls, nullable) or None
    def _leaf(self, e: Any) -> Optional[Tuple[int, int, int, str, bool]]:
        if e.as_type is not None:
            return None
        if e.kind == Kind.NAMED:
            t = self.t
            if e.name not in t.schema:
                raise KeyError(f"column {e.name} is not in {t.schema}")
            ci = t.schema.index_of_key(e.name)
            cls = _cls_of(t.schema.types[ci])
            return (K.XK_COL, ci, 0, cls, t.valid[ci] is not None)  # b = table column, slot assigned on use
        if e.kind == Kind.LITERAL:
            v = e.value
            if v is None:
                return (K.XK_NULL, 0, 0, "n", True)
            if isinstance(v, bool):
                return (K.XK_IMM, 0, int(v), "b", False)
            if isinstance(v, int):
                return (K.XK_IMM, 0, v & ((1 << 64) - 1), "i", False)
            if isinstance(v, float):
                return (K.XK_IMM, 0, _f64_bits(v), "f", False)
        return None

    def _emit_with(self, op: int, leaf: Tuple[int, int, int, str, bool], ctx: str) -> None:
        """``acc <- acc op leaf`` with the leaf converted to the context class ('i', 'f' or 'b')."""
        kind, b, imm, cls, _ = leaf
        flags = 0
        if kind == K.XK_COL:
            b = self.col_slot(b)
            if ctx == "f" and cls != "f":
                flags = K.XF_B_I2F
        elif kind == K.XK_IMM and ctx == "f" and cls != "f":
            v = imm - (1 << 64) if imm >= (1 << 63) else imm
            imm = _f64_bits(float(v))
        self.emit(op, kind, b, flags, imm)

    def _acc_to(self, cls: str, ctx: str) -> None:
        if cls == "n" or cls == ctx:
            return
        if ctx == "f":
            # peephole: the accumulator was just loaded from an integer leaf -> convert while loading
            if self.ins and self.ins[-1][0] == K.X_MOV and self.ins[-1][3] == 0:
                op, kind, b, _, imm = self.ins[-1]
                if kind == K.XK_COL:
                    self.ins[-1] = (op, kind, b, K.XF_B_I2F, imm)
                    return
                if kind == K.XK_IMM:
                    v = imm - (1 << 64) if imm >= (1 << 63) else imm
                    self.ins[-1] = (op, kind, b, 0, _f64_bits(float(v)))
                    return
            self.emit(K.X_I2F)
        elif ctx == "b":
            self.emit(K.X_TOBOOL_F if cls == "f" else K.X_TOBOOL_I)
        elif ctx == "i" and cls == "f":
            self.emit(K.X_F2I)

    # -- compilation: value ends up in the accumulator; returns (class, nullable)
    def compile(self, e: ColumnExpr, top: bool = False) -> Tuple[str, bool]:
        cls, nullable = self._node(e)
        if e.as_type is not None:
            want = _cls_of(e.as_type)
            if want == "s":
                if not top and cls != "s":
                    raise NotImplementedError(f"cast to str inside an expression (only on a whole column): {e}")
            elif cls == "s":
                raise NotImplementedError(f"cast of a string expression to {e.as_type}: {e}")
            elif cls != "n":
                self._acc_to(cls, want)
                cls = want
        return cls, nullable

    def _node(self, e: ColumnExpr) -> Tuple[str, bool]:  # noqa: C901
        if e.kind == Kind.WILDCARD:
            raise ValueError("'*' can't be evaluated as a value")
        if e.kind == Kind.AGG:
            raise ValueError(f"aggregation {e} in a row-wise expression")
        if e.kind in (Kind.NAMED, Kind.LITERAL):
            if e.kind == Kind.LITERAL and isinstance(e.value, str):
                raise NotImplementedError(f"string literal {e} outside a comparison with a string column")
            bare = e.cast(None) if e.as_type is not None else e
            leaf = self._leaf(bare)
            assert leaf is not None
            self._emit_with(K.X_MOV, leaf, leaf[3])
            return leaf[3], leaf[4]
        if e.kind == Kind.UNARY:
            cls, nullable = self.compile(e.col)
            if e.op in ("IS_NULL", "NOT_NULL"):  # string columns are fine here: only validity is read
                self.emit(K.X_IS_NULL if e.op == "IS_NULL" else K.X_NOT_NULL)
                return "b", False
            if cls == "s":
                raise NotImplementedError(f"{e.op} on a string expression: {e}")
            if e.op == "-":
                if cls == "n":
                    return cls, True
                self.emit(K.X_NEG_F if cls == "f" else K.X_NEG_I)
                return ("i" if cls == "b" else cls), nullable
            if e.op == "~":
                self._acc_to(cls, "b")
                self.emit(K.X_NOT)
                return "b", nullable
            raise NotImplementedError(f"unary operator {e.op}")
        if e.kind == Kind.BINARY:
            return self._binary(e)
        if e.kind == Kind.CALL:
            if e.func.upper() == "COALESCE":
                return self._coalesce(e)
            raise NotImplementedError(f"function {e.func} has no device implementation")
        raise NotImplementedError(f"can't evaluate {e!r}")

    def _binary(self, e: ColumnExpr) -> Tuple[str, bool]:
        op = e.op
        if op not in self._REVERSE:
            raise NotImplementedError(f"operator {op}")
        str_cmp = self._string_compare(e)
        if str_cmp is not None:
            return str_cmp
        cl, cr = self._static_cls(e.left), self._static_cls(e.right)
        if "s" in (cl, cr):
            raise NotImplementedError(f"operator {op} on string operands: {e}")
        logical = op in ("&", "|")
        ctx = "b" if logical else ("f" if (op == "/" or "f" in (cl, cr)) else "i")
        res = "b" if (logical or op in ("<", "<=", ">", ">=", "==", "!=")) else ctx

        def usable(leaf: Any) -> bool:  # a leaf whose class needs no instruction of its own in this context
            return leaf is not None and (not logical or leaf[3] in ("b", "n"))

        def code(o: str) -> int:
            if logical:
                return K.X_AND if o == "&" else K.X_OR
            oi, of = self._OPS[o]
            return of if ctx == "f" else oi

        lr, ll = self._leaf(e.right), self._leaf(e.left)
        if usable(lr):
            ca, na = self.compile(e.left)
            self._acc_to(ca, ctx)
            self._emit_with(code(op), lr, ctx)
            return res, na or lr[4]
        if usable(ll):
            cb, nb = self.compile(e.right)
            self._acc_to(cb, ctx)
            self._emit_with(code(self._REVERSE[op]), ll, ctx)
            return res, nb or ll[4]
        ca, na = self.compile(e.left)
        self._acc_to(ca, ctx)
        tmp = self.alloc()
        self.emit(K.X_ST, K.XK_NONE, tmp)
        cb, nb = self.compile(e.right)
        self._acc_to(cb, ctx)
        self.emit(code(self._REVERSE[op]), K.XK_REG, tmp)
        self.release(tmp)
        return res, na or nb

    def _string_compare(self, e: ColumnExpr) -> Optional[Tuple[str, bool]]:
        """``strcol == 'lit'`` / ``!=``: compare dictionary codes."""
        t = self.t
        sides = [e.left, e.right]
        named = [s.kind == Kind.NAMED and s.as_type is None and s.name in t.schema
                 and _is_str(t.schema[s.name].type) for s in sides]
        lits = [s.kind == Kind.LITERAL and isinstance(s.value, str) for s in sides]
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
        self.emit(K.X_MOV, K.XK_COL, self.col_slot(ci))
        self.emit(K.X_EQ_I if e.op == "==" else K.X_NE_I, K.XK_IMM, 0, 0,
                  (code if code is not None else -1) & ((1 << 64) - 1))
        return "b", t.valid[ci] is not None

    def _coalesce(self, e: ColumnExpr) -> Tuple[str, bool]:
        args = [a if isinstance(a, ColumnExpr) else _lit(a) for a in e.args]
        if len(args) == 0:
            raise ValueError("COALESCE needs arguments")
        probe = [self._static_cls(a) for a in args]
        if "s" in probe:
            raise NotImplementedError(f"COALESCE on strings: {e}")
        want = "f" if "f" in probe else ("b" if all(p in ("b", "n") for p in probe) and "b" in probe else "i")
        cls, nullable = self.compile(args[0])
        self._acc_to(cls, want)
        for a in args[1:]:
            leaf = self._leaf(a)
            if leaf is not None and (want != "b" or leaf[3] in ("b", "n")):
                self._emit_with(K.X_COALESCE, leaf, want)
                nullable = nullable and leaf[4]
                continue
            tmp = self.alloc()
            self.emit(K.X_ST, K.XK_NONE, tmp)
            cls, n = self.compile(a)
            self._acc_to(cls, want)
            self.emit(K.X_RCOALESCE, K.XK_REG, tmp)  # acc <- tmp if tmp is not NULL else acc
            self.release(tmp)
            nullable = nullable and n
        return want, nullable

    def _static_cls(self, e: ColumnExpr) -> str:
        """Class an expression will evaluate to (without emitting code)."""
        if e.as_type is not None:
            return _cls_of(e.as_type)
        if e.kind == Kind.NAMED:
            if e.name not in self.t.schema:
                raise KeyError(f"column {e.name} is not in {self.t.schema}")
            return _cls_of(self.t.schema[e.name].type)
        if e.kind == Kind.LITERAL:
            v = e.value
            return "n" if v is None else "b" if isinstance(v, bool) else "i" if isinstance(v, int) else \
                "f" if isinstance(v, float) else "s"
        if e.kind == Kind.UNARY:
            if e.op in ("IS_NULL", "NOT_NULL", "~"):
                return "b"
            c = self._static_cls(e.col)
            return "i" if c == "b" else c
        if e.kind == Kind.BINARY:
            if e.op in ("+", "-", "*", "/"):
                cs = (self._static_cls(e.left), self._static_cls(e.right))
                return "f" if (e.op == "/" or "f" in cs) else "i"
            return "b"
        if e.kind == Kind.CALL and e.func.upper() == "COALESCE":
            cs = [self._static_cls(a if isinstance(a, ColumnExpr) else _lit(a)) for a in e.args]
            return "f" if "f" in cs else ("b" if "b" in cs and all(c in ("b", "n") for c in cs) else "i")
        return "i"

    def run(self) -> Tuple[List[torch.Tensor], List[Optional[torch.Tensor]]]:
        t = self.t
        return K.eval_expr(t.num_rows, t.device, [t.columns[i] for i in self.cols],
                           [t.valid[i] for i in self.cols], self.ins, [o[0] for o in self.outs],
                           [o[1] for o in self.outs])


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
        if e.kind == Kind.NAMED:
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
        if e.kind == Kind.LITERAL and (isinstance(e.value, str) or e.value is None):
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
        batch: List[Tuple[int, ColumnExpr, str]] = []
        while k < len(pending):
            i, e = pending[k]
            mark = prog.mark()
            try:
                cls, nullable = prog.compile(e, top=True)
                if cls == "n":  # a bare NULL-valued expression
                    cls = _cls_of(e.as_type) if e.as_type is not None and not _is_str(e.as_type) else "i"
                tp = e.as_type if e.as_type is not None else _default_type(cls, e, t.schema)
                if _is_str(tp) or _cls_of(tp) != cls:
                    store_tp = {"i": pa.int64(), "f": pa.float64(), "b": pa.bool_()}[cls]
                    if not _is_str(tp):
                        tp = store_tp  # an inferred type of another class than the computed value
                else:
                    store_tp = tp
                prog.output(_storage_dtype(store_tp), nullable)
            except _OutOfResources as ex:
                if not batch:
                    raise NotImplementedError(f"expression too large for one device program ({ex}): {e}")
                prog.rollback(mark)
                break
            out_types[i] = tp
            batch.append((i, e, cls))
            k += 1
        cols, valids = prog.run()
        for (i, e, cls), c, v in zip(batch, cols, valids):
            if _is_str(out_types[i]):
                c, dicts[names[i]] = _to_string_column(c, v, cls)
            out_cols[i], out_valid[i] = c, v
    schema = Schema([pa.field(nm, tp) for nm, tp in zip(names, out_types)])
    return B200Table(schema, out_cols, out_valid, dicts)


def predicate_mask(t: B200Table, condition: ColumnExpr) -> torch.Tensor:
    """uint8 mask: 1 where ``condition`` is TRUE (NULL counts as FALSE, like SQL WHERE)."""
    prog = _Program(t)
    try:
        cls, _ = prog.compile(condition.alias("") if condition.as_name else condition)
        if cls == "s":
            raise ValueError(f"{condition} is not a boolean expression")
        if cls == "n":
            return torch.zeros(t.num_rows, dtype=torch.uint8, device=t.device)
        prog._acc_to(cls, "b")
        prog.output(torch.uint8, False)  # FB_X_OUT stores 0 for NULL
    except _OutOfResources as ex:
        raise NotImplementedError(f"condition too large for one device program ({ex}): {condition}")
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
    if e.has_args:
        args = [rewrite(a, mapper) for a in e.args]
        kwargs = {k: rewrite(v, mapper) for k, v in e.kwargs.items()}
        return ColumnExpr(e.kind, e.head, args, kwargs, e.is_distinct, e.as_name, e.as_type)
    return e


def find_aggs(e: Any, out: List[ColumnExpr]) -> None:
    if not isinstance(e, ColumnExpr):
        return
    if e.kind == Kind.AGG:
        out.append(e)
    elif e.has_args:
        for a in e.args:
            find_aggs(a, out)
        for a in e.kwargs.values():
            find_aggs(a, out)
