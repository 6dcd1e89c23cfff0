"""``B200SQLEngine``: recognises the two SQL shapes that reach the GPU kernels.

Reference: ``SQLEngine.select(dfs, statement)`` fugue/execution/execution_engine.py:209-238;
``StructuredRawSQL`` fugue/collections/sql.py:48-151 (pieces ``(is_table_ref, text)``);
FugueSQL forwards plain SELECT text to it (fugue/sql/_visitors.py:743-766,
fugue/workflow/workflow.py:2109-2166) and ``ExecutionEngine.aggregate`` reaches it through
``SQLExpressionGenerator`` (fugue/column/sql.py:275-334).  The native engine hands the text to qpd
(native_execution_engine.py:59-66); here:

    SELECT [DISTINCT] <expr [AS a], ...> FROM t [WHERE <expr>] [GROUP BY <expr, ...>] [HAVING <expr>]
             [ORDER BY c [ASC|DESC], ...] [LIMIT n]
        -> parsed into column expressions (fugue_b200.column) and run by ``engine.select``: row-wise
           parts in the device expression evaluator, SUM/COUNT/MIN/MAX/AVG in the hash group-by kernel
    SELECT * FROM a [INNER|LEFT|RIGHT|FULL [OUTER]|LEFT SEMI|LEFT ANTI|CROSS] JOIN b
             [USING (k, ...) | ON a.k = b.k [AND ...]]                        -> hash join kernels

Expressions: + - * /, comparisons (= == != <> < <= > >=), AND / OR / NOT, IS [NOT] NULL, [NOT] IN (...),
[NOT] BETWEEN, CAST(x AS type), COALESCE, literals, `quoted` and table-qualified names.
Anything else raises NotImplementedError (there is no host SQL fallback in this package).
"""
import re
from typing import Any, Dict, List, Tuple

from .column import ColumnExpr, Kind, SelectColumns, all_cols, col, function, functions, is_agg, lit, null
from .dataframe import DataFrame

_AGG = r"(SUM|COUNT|MIN|MAX|AVG|MEAN)\s*\(\s*(\*|[A-Za-z_][\w]*)\s*\)"
_IDENT = r"[A-Za-z_][\w]*"


_TEMP_TABLE_EXPR_PREFIX, _TEMP_TABLE_EXPR_SUFFIX = "<tmpdf:", ">"


class TempTableName:
    """A random table name that prints as ``<tmpdf:_XXXXX>`` - the placeholder ``from_expr`` recognises
    (fugue/collections/sql.py:14-21)."""

    def __init__(self) -> None:
        import uuid

        self.key = "_" + uuid.uuid4().hex[:5].upper()

    def __repr__(self) -> str:
        return _TEMP_TABLE_EXPR_PREFIX + self.key + _TEMP_TABLE_EXPR_SUFFIX


class StructuredRawSQL:
    """SQL text as ``(is_table_ref, text)`` pieces plus the dialect it is written in
    (fugue/collections/sql.py:48-151).  ``construct`` joins the pieces with single spaces, table references
    mapped through ``name_map`` (a function, or a dict - names it does not hold stay as they are).  A change
    of dialect needs sqlglot (``transpile_sql``, :24-45), which this image does not have: asking for one
    raises; the engine's own dialect is None, which never transpiles (:99-103)."""

    def __init__(self, statements: Any, dialect: Any = None):
        self._statements: List[Tuple[bool, str]] = [(bool(a), str(b)) for a, b in statements]
        self._dialect = dialect

    @property
    def dialect(self) -> Any:
        return self._dialect

    def __uuid__(self) -> str:
        import json
        import uuid

        return str(uuid.uuid5(uuid.NAMESPACE_OID, json.dumps([self._statements, self._dialect])))

    @staticmethod
    def from_expr(sql: str, prefix: str = _TEMP_TABLE_EXPR_PREFIX, suffix: str = _TEMP_TABLE_EXPR_SUFFIX,
                  dialect: Any = None) -> "StructuredRawSQL":
        pieces: List[Tuple[bool, str]] = []
        pos = 0
        while pos < len(sql):
            start = sql.find(prefix, pos)
            if start < 0:
                pieces.append((False, sql[pos:]))
                break
            if start > pos:
                pieces.append((False, sql[pos:start]))
            end = sql.find(suffix, start + len(prefix))
            if end < 0:
                raise SyntaxError(f"unterminated table reference in {sql!r}")
            pieces.append((True, sql[start + len(prefix):end]))
            pos = end + len(suffix)
        return StructuredRawSQL(pieces, dialect=dialect)

    def construct(self, name_map: Any = None, dialect: Any = None, log: Any = None) -> str:
        if name_map is None:
            rename = lambda x: x  # noqa: E731
        elif isinstance(name_map, dict):
            rename = lambda x: name_map.get(x, x)  # noqa: E731
        else:
            rename = name_map
        text = " ".join(rename(t) if is_ref else t for is_ref, t in self._statements)
        if self._dialect is not None and dialect is not None and self._dialect != dialect:
            raise NotImplementedError(
                f"SQL transpilation {self._dialect} -> {dialect} needs sqlglot, which is not installed")
        return text


class B200SQLEngine:
    """The SQL facet (``SQLEngine``, fugue/execution/execution_engine.py:183-274)."""

    def __init__(self, execution_engine: Any):
        import uuid

        self._engine = execution_engine
        self._uid = "_" + uuid.uuid4().hex[:5] + "_"

    @property
    def execution_engine(self) -> Any:
        return self._engine

    @property
    def log(self) -> Any:
        return self._engine.log

    @property
    def conf(self) -> Any:
        return self._engine.conf

    def to_df(self, df: Any, schema: Any = None) -> Any:
        return self._engine.to_df(df, schema)

    def encode_name(self, name: str) -> str:
        """A table name no other statement of this process uses (:197-198)."""
        return self._uid + name

    def encode(self, dfs: Dict[str, Any], statement: "StructuredRawSQL") -> Tuple[Dict[str, Any], str]:
        """Tables and statement text with the table references renamed consistently (:200-207)."""
        return ({self.encode_name(k): v for k, v in dfs.items()},
                statement.construct(self.encode_name, dialect=self.dialect))

    def table_exists(self, table: str) -> bool:
        raise NotImplementedError("the b200 SQL engine has no table catalogue")

    def load_table(self, table: str, **kwargs: Any) -> Any:
        raise NotImplementedError("the b200 SQL engine has no table catalogue")

    def save_table(self, df: Any, table: str, mode: str = "overwrite", partition_spec: Any = None,
                   **kwargs: Any) -> None:
        raise NotImplementedError("the b200 SQL engine has no table catalogue")

    @property
    def dialect(self) -> Any:
        return None  # no transpile (fugue/collections/sql.py:99-103)

    @property
    def is_distributed(self) -> bool:
        return self._engine.is_distributed

    def select(self, dfs: Dict[str, Any], statement: Any) -> DataFrame:
        sql = statement.construct() if isinstance(statement, StructuredRawSQL) else str(statement)
        sql = re.sub(r"\s+", " ", sql.strip().rstrip(";"))
        tables = {k: self._engine.to_df(v) for k, v in dfs.items()}
        m = re.match(r"(?is)^SELECT (.+?) FROM (.+)$", sql)
        if m is None:
            raise NotImplementedError(f"unsupported SQL: {sql}")
        items, rest = m.group(1).strip(), m.group(2).strip()
        if re.search(r"(?i)\bJOIN\b", rest):
            return self._join(items, rest, tables, sql)
        return self._single(items, rest, tables, sql)

    def _table(self, name: str, tables: Dict[str, DataFrame], sql: str) -> DataFrame:
        name = name.strip().strip("`")
        if name not in tables:
            raise KeyError(f"table {name} is not among {list(tables)} in: {sql}")
        return tables[name]

    def _single(self, items: str, rest: str, tables: Dict[str, DataFrame], sql: str) -> DataFrame:
        st = _parse_select(items, rest, sql)
        df = self._table(st.table, tables, sql)
        cols = list(st.columns)
        hidden: List[str] = []
        if st.group_by:
            # Fugue infers the GROUP BY keys from the select list (SelectColumns.group_keys); the SQL
            # text must agree with that, extra keys ride along as hidden columns
            probe = SelectColumns(*cols)
            if not probe.has_agg:
                raise NotImplementedError(f"GROUP BY without aggregates: {sql}")
            inferred = {k.fingerprint() for k in probe.group_keys}
            listed = set()
            for g in st.group_by:
                uid = g.alias("").cast(None).fingerprint()
                listed.add(uid)
                if uid not in inferred:
                    name = f"__fb_g{len(hidden)}"
                    hidden.append(name)
                    cols.append(g.alias(name))
            for k in probe.group_keys:
                if k.fingerprint() not in listed:
                    raise ValueError(f"{k} is neither aggregated nor in GROUP BY: {sql}")
        res = self._engine.select(df, SelectColumns(*cols, arg_distinct=st.distinct), where=st.where,
                                  having=st.having)
        if hidden:
            res = res[[n for n in res.columns if n not in hidden]]
        if st.order_by:
            from collections import OrderedDict

            from .dataframe import B200DataFrame
            from .sort import sort_table

            sorts = OrderedDict((n, asc) for n, asc in st.order_by)
            for n in sorts:
                if n not in res.schema:
                    raise ValueError(f"ORDER BY {n}: not an output column of: {sql}")
            res = B200DataFrame(sort_table(res.native, sorts, "last"))
        if st.limit is not None:
            from .dataframe import B200DataFrame

            res = B200DataFrame(res.native.slice(0, min(st.limit, res.native.num_rows)))
        return res

    def _join(self, items: str, rest: str, tables: Dict[str, DataFrame], sql: str) -> DataFrame:
        if items.strip() != "*":
            raise NotImplementedError(f"only SELECT * is supported for joins: {sql}")
        m = re.match(rf"(?is)^(`?{_IDENT}`?)(?:\s+AS)?(?:\s+({_IDENT}))?\s+"
                     r"((?:INNER|CROSS|LEFT SEMI|LEFT ANTI|SEMI|ANTI|LEFT OUTER|RIGHT OUTER|FULL OUTER|"
                     r"LEFT|RIGHT|FULL)\s+)?JOIN\s+"
                     rf"(`?{_IDENT}`?)(?:\s+AS)?(?:\s+({_IDENT}))?(?:\s+(USING|ON)\s+(.+))?$", rest)
        if m is None:
            raise NotImplementedError(f"unsupported join SQL: {sql}")
        t1, t2 = self._table(m.group(1), tables, sql), self._table(m.group(4), tables, sql)
        kind = (m.group(3) or "INNER").strip().upper()
        how = {"INNER": "inner", "CROSS": "cross", "LEFT SEMI": "semi", "SEMI": "semi", "LEFT ANTI": "anti",
               "ANTI": "anti", "LEFT OUTER": "left_outer", "LEFT": "left_outer", "RIGHT OUTER": "right_outer",
               "RIGHT": "right_outer", "FULL OUTER": "full_outer", "FULL": "full_outer"}[kind]
        on = None
        if m.group(6):
            cond = m.group(7).strip()
            if m.group(6).upper() == "USING":
                on = [c.strip().strip("`") for c in cond.strip("() ").split(",")]
            else:
                on = []
                for part in re.split(r"(?i)\s+AND\s+", cond):
                    mm = re.match(rf"^\(?\s*(?:{_IDENT}\.)?({_IDENT})\s*=\s*(?:{_IDENT}\.)?({_IDENT})\s*\)?$", part.strip())
                    if mm is None or mm.group(1) != mm.group(2):
                        raise NotImplementedError(f"only equi-joins on equally named columns: {sql}")
                    on.append(mm.group(1))
        return self._engine.join(t1, t2, how=how, on=on)


def _split_commas(text: str) -> List[str]:
    out, depth, cur = [], 0, []
    for ch in text:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out


# ---------------------------------------------------------------------------------------------
# SELECT statement parser (single table) -> column expressions
# ---------------------------------------------------------------------------------------------
class _Select:
    def __init__(self) -> None:
        self.distinct = False
        self.columns: List[ColumnExpr] = []
        self.table = ""
        self.where: Any = None
        self.group_by: List[ColumnExpr] = []
        self.having: Any = None
        self.order_by: List[Tuple[str, bool]] = []
        self.limit: Any = None


_TOKEN = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)
  | (?P<str>'(?:[^'\\]|\\.|'')*')
  | (?P<bq>`(?:[^`]|``)*`)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op><=|>=|<>|!=|==|=|<|>|\+|-|\*|/|\(|\)|,|\.)
)""", re.X)

_AGG_FUNCS = {"SUM": functions.sum, "COUNT": functions.count, "MIN": functions.min, "MAX": functions.max,
              "AVG": functions.avg, "MEAN": functions.avg, "FIRST": functions.first, "LAST": functions.last}
_CLAUSES = ("WHERE", "GROUP", "HAVING", "ORDER", "LIMIT")


def _tokenize(text: str, sql: str) -> List[Tuple[str, str]]:
    out: List[Tuple[str, str]] = []
    pos = 0
    text = text.rstrip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if m is None or m.end() == pos:
            raise NotImplementedError(f"can't tokenize {text[pos:pos + 20]!r} in: {sql}")
        kind = m.lastgroup
        out.append((kind, m.group(kind)))
        pos = m.end()
    return out


class _Parser:
    """Recursive descent over the token list; precedence OR < AND < NOT < comparison < + - < * / < unary."""

    def __init__(self, tokens: List[Tuple[str, str]], sql: str):
        self.t = tokens
        self.i = 0
        self.sql = sql

    def peek(self, k: int = 0) -> Tuple[str, str]:
        return self.t[self.i + k] if self.i + k < len(self.t) else ("end", "")

    def kw(self, *words: str) -> bool:
        """Consume the keyword sequence if it is next."""
        for k, w in enumerate(words):
            kind, val = self.peek(k)
            if kind != "id" or val.upper() != w:
                return False
        self.i += len(words)
        return True

    def at_kw(self, word: str) -> bool:
        kind, val = self.peek()
        return kind == "id" and val.upper() == word

    def op(self, sym: str) -> bool:
        if self.peek() == ("op", sym):
            self.i += 1
            return True
        return False

    def expect(self, sym: str) -> None:
        if not self.op(sym):
            raise NotImplementedError(f"expected {sym!r} near token {self.i} in: {self.sql}")

    def fail(self, what: str) -> Any:
        raise NotImplementedError(f"{what} in: {self.sql}")

    # ---- expressions
    def expr(self) -> ColumnExpr:
        e = self.and_expr()
        while self.kw("OR"):
            e = e | self.and_expr()
        return e

    def and_expr(self) -> ColumnExpr:
        e = self.not_expr()
        while self.kw("AND"):
            e = e & self.not_expr()
        return e

    def not_expr(self) -> ColumnExpr:
        if self.kw("NOT"):
            return ~self.not_expr()
        return self.comparison()

    def comparison(self) -> ColumnExpr:
        e = self.additive()
        while True:
            if self.kw("IS", "NOT", "NULL"):
                e = e.not_null()
            elif self.kw("IS", "NULL"):
                e = e.is_null()
            elif self.at_kw("NOT") and self.peek(1)[1].upper() in ("IN", "BETWEEN"):
                self.i += 1
                e = ~self._in_or_between(e)
            elif self.at_kw("IN") or self.at_kw("BETWEEN"):
                e = self._in_or_between(e)
            else:
                kind, val = self.peek()
                if kind == "op" and val in ("=", "==", "!=", "<>", "<", "<=", ">", ">="):
                    self.i += 1
                    r = self.additive()
                    e = {"=": e == r, "==": e == r, "!=": e != r, "<>": e != r, "<": e < r, "<=": e <= r,
                         ">": e > r, ">=": e >= r}[val]
                else:
                    return e

    def _in_or_between(self, e: ColumnExpr) -> ColumnExpr:
        if self.kw("IN"):
            self.expect("(")
            res: Any = None
            while True:
                c = e == self.additive()
                res = c if res is None else (res | c)
                if not self.op(","):
                    break
            self.expect(")")
            return res
        self.kw("BETWEEN")
        lo = self.additive()
        if not self.kw("AND"):
            self.fail("BETWEEN without AND")
        hi = self.additive()
        return (e >= lo) & (e <= hi)

    def additive(self) -> ColumnExpr:
        e = self.multiplicative()
        while True:
            if self.op("+"):
                e = e + self.multiplicative()
            elif self.op("-"):
                e = e - self.multiplicative()
            else:
                return e

    def multiplicative(self) -> ColumnExpr:
        e = self.unary()
        while True:
            if self.op("*"):
                e = e * self.unary()
            elif self.op("/"):
                e = e / self.unary()
            else:
                return e

    def unary(self) -> ColumnExpr:
        if self.op("-"):
            kind, val = self.peek()
            if kind == "num":  # a negative literal, not a negated expression
                self.i += 1
                return lit(-_number(val))
            return -self.unary()
        if self.op("+"):
            return self.unary()
        return self.primary()

    def primary(self) -> ColumnExpr:
        kind, val = self.peek()
        if kind == "num":
            self.i += 1
            return lit(_number(val))
        if kind == "str":
            self.i += 1
            body = val[1:-1].replace("''", "'")
            return lit(re.sub(r"\\(.)", r"\1", body))
        if kind == "op" and val == "(":
            self.i += 1
            e = self.expr()
            self.expect(")")
            return e
        if kind == "bq":
            self.i += 1
            return self._maybe_qualified(val[1:-1].replace("``", "`"))
        if kind == "id":
            up = val.upper()
            if up == "NULL":
                self.i += 1
                return null()
            if up in ("TRUE", "FALSE"):
                self.i += 1
                return lit(up == "TRUE")
            if up == "CAST" and self.peek(1) == ("op", "("):
                self.i += 2
                e = self.expr()
                if not self.kw("AS"):
                    self.fail("CAST without AS")
                tp = []
                while self.peek() != ("op", ")") and self.peek()[0] != "end":
                    tp.append(self.peek()[1])
                    self.i += 1
                self.expect(")")
                return e.cast("".join(tp).lower())
            if self.peek(1) == ("op", "("):
                return self._call(up)
            self.i += 1
            return self._maybe_qualified(val)
        return self.fail(f"unexpected token {val!r}")

    def _maybe_qualified(self, name: str) -> ColumnExpr:
        if self.peek() == ("op", ".") and self.peek(1)[0] in ("id", "bq"):  # table.column
            kind, val = self.peek(1)
            self.i += 2
            name = val[1:-1].replace("``", "`") if kind == "bq" else val
        return col(name)

    def _call(self, fn: str) -> ColumnExpr:
        self.i += 2  # name (
        if fn in _AGG_FUNCS:
            distinct = self.kw("DISTINCT")
            if self.op("*"):
                arg: ColumnExpr = all_cols()
            else:
                arg = self.expr()
            self.expect(")")
            if distinct:
                if fn != "COUNT":
                    self.fail(f"{fn}(DISTINCT ...)")
                return functions.count_distinct(arg)
            return _AGG_FUNCS[fn](arg)
        args: List[Any] = []
        if not self.op(")"):
            while True:
                args.append(self.expr())
                if not self.op(","):
                    break
            self.expect(")")
        if fn == "COALESCE":
            return functions.coalesce(*args)
        return function(fn, *args)

    # ---- select items / lists
    def item(self) -> ColumnExpr:
        if self.peek() == ("op", "*"):
            self.i += 1
            return all_cols()
        e = self.expr()
        if self.kw("AS"):
            kind, val = self.peek()
            if kind not in ("id", "bq"):
                self.fail("AS without a name")
            self.i += 1
            return e.alias(val[1:-1].replace("``", "`") if kind == "bq" else val)
        kind, val = self.peek()
        if kind == "bq" or (kind == "id" and val.upper() not in _CLAUSES + ("FROM",)):
            self.i += 1  # implicit alias
            return e.alias(val[1:-1].replace("``", "`") if kind == "bq" else val)
        return e


def _number(text: str) -> Any:
    return int(text) if re.fullmatch(r"\d+", text) else float(text)


def _default_alias(e: ColumnExpr) -> ColumnExpr:
    """Name an unnamed select item the way the SQL engines do for the common cases."""
    if e.kind == Kind.WILDCARD or e.output_name != "":
        return e
    if e.kind == Kind.AGG and e.arg.kind == Kind.WILDCARD:
        return e.alias(e.func.lower())          # COUNT(*) -> "count"
    named = e.infer_alias()
    return named


def _parse_select(items: str, rest: str, sql: str) -> _Select:
    st = _Select()
    p = _Parser(_tokenize(items, sql), sql)
    st.distinct = p.kw("DISTINCT")
    while True:
        st.columns.append(_default_alias(p.item()))
        if not p.op(","):
            break
    if p.peek()[0] != "end":
        p.fail(f"unexpected token {p.peek()[1]!r} in the select list")
    p = _Parser(_tokenize(rest, sql), sql)
    kind, val = p.peek()
    if kind == "op" and val == "(":
        p.fail("sub-queries are not on the GPU path")
    if kind not in ("id", "bq"):
        p.fail("FROM needs a table name")
    p.i += 1
    st.table = val.strip("`")
    kind, val = p.peek()  # optional alias
    if p.kw("AS"):
        p.i += 1
    elif kind == "id" and val.upper() not in _CLAUSES:
        p.i += 1
    if p.kw("WHERE"):
        st.where = p.expr()
    if p.kw("GROUP", "BY"):
        while True:
            st.group_by.append(p.expr())
            if not p.op(","):
                break
    if p.kw("HAVING"):
        st.having = p.expr()
    if p.kw("ORDER", "BY"):
        while True:
            kind, val = p.peek()
            if kind not in ("id", "bq"):
                p.fail("ORDER BY takes output column names")
            p.i += 1
            asc = True
            if p.kw("DESC"):
                asc = False
            else:
                p.kw("ASC")
            st.order_by.append((val.strip("`"), asc))
            if not p.op(","):
                break
    if p.kw("LIMIT"):
        kind, val = p.peek()
        if kind != "num" or not val.isdigit():
            p.fail("LIMIT takes an integer")
        p.i += 1
        st.limit = int(val)
    if p.peek()[0] != "end":
        p.fail(f"unsupported SQL near {p.peek()[1]!r}")
    if st.where is not None and is_agg(st.where):
        raise ValueError(f"aggregation in WHERE: {sql}")
    return st
