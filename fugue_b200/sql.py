"""``B200SQLEngine``: recognises the two SQL shapes that reach the GPU kernels.

Reference: ``SQLEngine.select(dfs, statement)`` fugue/execution/execution_engine.py:209-238;
``StructuredRawSQL`` fugue/collections/sql.py:48-151 (pieces ``(is_table_ref, text)``);
FugueSQL forwards plain SELECT text to it (fugue/sql/_visitors.py:743-766,
fugue/workflow/workflow.py:2109-2166) and ``ExecutionEngine.aggregate`` reaches it through
``SQLExpressionGenerator`` (fugue/column/sql.py:275-334).  The native engine hands the text to qpd
(native_execution_engine.py:59-66); here:

    SELECT k[, ...], AGG(x) [AS a], COUNT(*) [AS c] FROM t [GROUP BY k]      -> hash group-by kernel
    SELECT * FROM a [INNER|LEFT|RIGHT|FULL [OUTER]|LEFT SEMI|LEFT ANTI|CROSS] JOIN b
             [USING (k, ...) | ON a.k = b.k [AND ...]]                        -> hash join kernels
    SELECT c1, c2 FROM t  /  SELECT * FROM t                                  -> column projection

Anything else raises NotImplementedError (there is no host SQL fallback in this package).
"""
import re
from typing import Any, Dict, List, Tuple

from .column import AggFuncExpr, col
from .dataframe import DataFrame
from .partition import PartitionSpec

_AGG = r"(SUM|COUNT|MIN|MAX|AVG|MEAN)\s*\(\s*(\*|[A-Za-z_][\w]*)\s*\)"
_IDENT = r"[A-Za-z_][\w]*"


class StructuredRawSQL:
    """(is_table_ref, text) pieces - fugue/collections/sql.py:48-151."""

    def __init__(self, statements: Any):
        self._pieces: List[Tuple[bool, str]] = [(bool(a), str(b)) for a, b in statements]

    @staticmethod
    def from_expr(sql: str, prefix: str = "<tmpdf:", suffix: str = ">") -> "StructuredRawSQL":
        pieces: List[Tuple[bool, str]] = []
        pos = 0
        for m in re.finditer(re.escape(prefix) + r"([^>]+)" + re.escape(suffix), sql):
            if m.start() > pos:
                pieces.append((False, sql[pos:m.start()]))
            pieces.append((True, m.group(1)))
            pos = m.end()
        if pos < len(sql):
            pieces.append((False, sql[pos:]))
        return StructuredRawSQL(pieces)

    def construct(self, name_map: Any = None) -> str:
        f = (lambda x: x) if name_map is None else (name_map if callable(name_map) else (lambda x: name_map[x]))
        return " ".join(f(t) if is_t else t.strip() for is_t, t in self._pieces).strip()


class B200SQLEngine:
    def __init__(self, execution_engine: Any):
        self._engine = execution_engine

    @property
    def execution_engine(self) -> Any:
        return self._engine

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
        m = re.match(rf"(?is)^(`?{_IDENT}`?)(?:\s+GROUP BY (.+))?$", rest)
        if m is None:
            raise NotImplementedError(f"unsupported SQL (WHERE/HAVING/ORDER BY are not on the GPU path): {sql}")
        df = self._table(m.group(1), tables, sql)
        group = [g.strip().strip("`") for g in m.group(2).split(",")] if m.group(2) else []
        plain: List[Tuple[str, str]] = []
        aggs: List[AggFuncExpr] = []
        order: List[str] = []
        for raw in _split_commas(items):
            it = raw.strip()
            ma = re.match(rf"(?is)^{_AGG}(?:\s+AS\s+(`?{_IDENT}`?))?$", it)
            if ma:
                func = "AVG" if ma.group(1).upper() == "MEAN" else ma.group(1).upper()
                alias = (ma.group(3) or "").strip("`")
                if alias == "":
                    alias = ma.group(2) if ma.group(2) != "*" else func.lower()
                aggs.append(AggFuncExpr(func, col(ma.group(2)), alias))
                order.append(alias)
                continue
            mc = re.match(rf"(?is)^(\*|`?{_IDENT}`?)(?:\s+AS\s+(`?{_IDENT}`?))?$", it)
            if mc is None:
                raise NotImplementedError(f"unsupported select item {it!r} in: {sql}")
            name = mc.group(1).strip("`")
            plain.append((name, (mc.group(2) or name).strip("`")))
            order.append((mc.group(2) or name).strip("`"))
        if not aggs:
            if group:
                raise NotImplementedError(f"GROUP BY without aggregates: {sql}")
            if len(plain) == 1 and plain[0][0] == "*":
                return df
            res = df[[p[0] for p in plain]]
            ren = {a: b for a, b in plain if a != b}
            return res.rename(ren) if ren else res
        for name, _ in plain:
            if name not in group:
                raise ValueError(f"{name} is neither aggregated nor in GROUP BY: {sql}")
        res = self._engine.aggregate(df, PartitionSpec(by=group) if group else None, aggs)
        ren = {a: b for a, b in plain if a != b}
        if ren:
            res = res.rename(ren)
        want = [o for o in order if o in res.schema]
        return res[want] if want != res.columns else res

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
