"""``PartitionSpec`` / ``PartitionCursor`` with the reference's names and semantics.

Mirrors fugue/collections/partition.py (PartitionSpec :79-333, PartitionCursor
:404-469, parse_presort_exp :13-76) without the ``triad`` dependency, so that
the parity tests read like the reference's own
(tests/fugue/collections/test_partition.py, fugue_test/execution_suite.py:208-256).
"""
import json
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional

from .schema import Schema

KEYWORD_ROWCOUNT = "ROWCOUNT"        # fugue/constants.py:10
KEYWORD_PARALLELISM = "CONCURRENCY"  # fugue/constants.py:11
_ALGOS = ("", "default", "hash", "rand", "even", "coarse")


def _split_outside_quotes(text: str, sep: str) -> List[str]:
    """Split on ``sep`` except inside `backquoted names` (a doubled backquote is an escape)."""
    out: List[str] = []
    cur: List[str] = []
    quoted = False
    i = 0
    while i < len(text):
        ch = text[i]
        if ch == "`":
            if quoted and i + 1 < len(text) and text[i + 1] == "`":
                cur.append("``")
                i += 2
                continue
            quoted = not quoted
            cur.append(ch)
        elif ch == sep and not quoted:
            if cur or sep != " ":
                out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
        i += 1
    if quoted:
        raise SyntaxError(f"unbalanced quote in {text}")
    if cur or sep != " ":
        out.append("".join(cur))
    return [x for x in out if x != ""] if sep == " " else out


def _unquote(name: str) -> str:
    name = name.strip()
    if len(name) >= 2 and name[0] == "`" and name[-1] == "`":
        return name[1:-1].replace("``", "`")
    return name


def _to_size(v: Any) -> int:
    """"5k" -> 5120 (triad.utils.convert.to_size)."""
    if v is None or v == "":
        return 0
    if isinstance(v, (int, float)):
        return int(v)
    t = str(v).strip().lower()
    mult = {"k": 1 << 10, "m": 1 << 20, "g": 1 << 30, "t": 1 << 40}
    t = t[:-1] if t.endswith("b") and len(t) > 1 and not t[-2].isdigit() else t.rstrip("b") if t.endswith("b") else t
    if t and t[-1] in mult:
        return int(float(t[:-1]) * mult[t[-1]])
    return int(float(t))


def parse_presort_exp(presort: Any) -> "OrderedDict[str, bool]":
    """``"b desc, c"`` or ``[("b", False), "c"]`` -> ordered {column: ascending}."""
    if isinstance(presort, OrderedDict):
        return presort
    res: "OrderedDict[str, bool]" = OrderedDict()
    if presort is None:
        return res
    pairs: List[Any] = []
    if isinstance(presort, str):
        text = presort.strip()
        if text == "":
            return res
        for item in _split_outside_quotes(text, ","):
            tokens = _split_outside_quotes(item.strip(), " ")
            if len(tokens) == 1:
                pairs.append((_unquote(tokens[0]), True))
            elif len(tokens) == 2 and tokens[1].lower() in ("asc", "desc"):
                pairs.append((_unquote(tokens[0]), tokens[1].lower() == "asc"))
            else:
                raise SyntaxError(f"Invalid expression {presort}")
    elif isinstance(presort, (list, tuple)):
        for item in presort:
            if isinstance(item, str):
                pairs.append((item, True))
            elif (isinstance(item, tuple) and len(item) == 2 and isinstance(item[0], str)
                  and isinstance(item[1], bool)):
                pairs.append(item)
            else:
                raise SyntaxError(f"Invalid expression {presort}")
    else:
        raise SyntaxError(f"Invalid expression {presort}")
    for name, asc in pairs:
        if name in res:
            raise SyntaxError(f"Invalid expression {presort} duplicated key {name}")
        res[name] = asc
    return res


class PartitionSpec:
    """``algo`` (default/hash/rand/even/coarse), ``num`` (int or expression over
    ROWCOUNT / CONCURRENCY), ``by`` (partition keys), ``presort``."""

    def __init__(self, *args: Any, **kwargs: Any):
        p: Dict[str, Any] = {}
        handled = False
        if len(args) == 1 and len(kwargs) == 0:
            a = args[0]
            if isinstance(a, str):
                if a.lower() == "per_row":
                    p["algo"] = "even"
                    p["num_partitions"] = KEYWORD_ROWCOUNT
                    handled = True
                elif not a.startswith("{"):
                    p["partition_by"] = [a]
                    handled = True
            elif isinstance(a, bool):
                raise TypeError(f"{a} is not supported")
            elif isinstance(a, int):
                p["num_partitions"] = str(a)
                handled = True
            elif isinstance(a, (list, tuple)):
                p["partition_by"] = list(a)
                handled = True
        if not handled:
            for a in args:
                if a is None:
                    continue
                if isinstance(a, PartitionSpec):
                    self._merge(p, a.jsondict)
                elif isinstance(a, dict):
                    self._merge(p, a)
                elif isinstance(a, str):
                    self._merge(p, json.loads(a))
                else:
                    raise TypeError(f"{a} is not supported")
            self._merge(p, kwargs)
        self._num_partitions = str(p.get("num_partitions", "0"))
        self._algo = str(p.get("algo", "")).lower()
        by = p.get("partition_by", [])
        if isinstance(by, str):
            by = [by]
        elif isinstance(by, (list, tuple)):
            by = list(by)
        else:
            raise SyntaxError(by)
        if len(by) != len(set(by)):
            raise SyntaxError(f"{by} has duplicated keys")
        self._partition_by: List[str] = by
        self._presort = parse_presort_exp(p.get("presort", None))
        if any(k in self._presort for k in self._partition_by):
            raise SyntaxError(
                f"partition by overlap with presort: {self._partition_by}, {self._presort}")
        self._size_limit = _to_size(p.get("size_limit", 0))
        self._row_limit = int(p.get("row_limit", 0) or 0)

    @staticmethod
    def _merge(d: Dict[str, Any], u: Dict[str, Any]) -> None:
        for k, v in u.items():
            if k == "by":
                k = "partition_by"
            elif k == "num":
                k = "num_partitions"
            d[k] = v

    def __repr__(self) -> str:
        return (f"PartitionSpec(num='{self._num_partitions}', by={self._partition_by}, "
                f"presort='{self.presort_expr}')")

    def __eq__(self, other: Any) -> bool:
        if other is self:
            return True
        if not isinstance(other, PartitionSpec):
            other = PartitionSpec(other)
        return self.jsondict == other.jsondict

    @property
    def empty(self) -> bool:
        return (self._num_partitions == "0" and self._algo == "" and not self._partition_by
                and not self._presort and self._size_limit == 0 and self._row_limit == 0)

    @property
    def num_partitions(self) -> str:
        return self._num_partitions

    def get_num_partitions(self, **expr_map_funcs: Callable[[], Any]) -> int:
        expr = self._num_partitions
        for k, fn in expr_map_funcs.items():
            if k in expr:
                expr = expr.replace(k, str(fn()))
        # the reference evaluates the text with a bare eval (partition.py:203-207); here only arithmetic and
        # a few pure functions are visible ("min(ROWCOUNT,CONCURRENCY)" is one of its test cases)
        import math

        allowed = dict(min=min, max=max, abs=abs, round=round, int=int, float=float, pow=pow,
                       ceil=math.ceil, floor=math.floor, sqrt=math.sqrt, log2=math.log2)
        return int(eval(expr, {"__builtins__": {}}, allowed))

    @property
    def algo(self) -> str:
        return self._algo if self._algo != "" else "default"

    @property
    def partition_by(self) -> List[str]:
        return self._partition_by

    @property
    def presort(self) -> "OrderedDict[str, bool]":
        return self._presort

    @property
    def presort_expr(self) -> str:
        return ",".join(f"{k} {'ASC' if v else 'DESC'}" for k, v in self._presort.items())

    @property
    def jsondict(self) -> Dict[str, Any]:
        return dict(num_partitions=self._num_partitions, algo=self._algo,
                    partition_by=self._partition_by, presort=self.presort_expr,
                    size_limit=self._size_limit, row_limit=self._row_limit)

    def __uuid__(self) -> str:
        """Deterministic id of the spec (fugue/collections/partition.py:259-261: the id of ``jsondict``):
        equal for specs that mean the same (``num=0`` and no argument, ``num=2`` and ``num="2"``), different
        when the key ORDER differs.  A name-based UUID over the canonical JSON text of ``jsondict``."""
        import json
        import uuid

        return str(uuid.uuid5(uuid.NAMESPACE_OID, json.dumps(self.jsondict, sort_keys=True)))

    def get_sorts(self, schema: Schema, with_partition_keys: bool = True) -> "OrderedDict[str, bool]":
        d: "OrderedDict[str, bool]" = OrderedDict()
        if with_partition_keys:
            for k in self._partition_by:
                if k not in schema:
                    raise KeyError(f"{k} not in {schema}")
                d[k] = True
        for k, v in self._presort.items():
            if k not in schema:
                raise KeyError(f"{k} not in {schema}")
            d[k] = v
        return d

    def get_key_schema(self, schema: Schema) -> Schema:
        return schema.extract(self._partition_by)

    def get_cursor(self, schema: Schema, physical_partition_no: int) -> "PartitionCursor":
        return PartitionCursor(schema, self, physical_partition_no)


class PartitionCursor:
    """Points at the first row of the current logical partition
    (fugue/collections/partition.py:404-469)."""

    def __init__(self, schema: Schema, spec: PartitionSpec, physical_partition_no: int):
        self._orig_schema = schema
        self._key_index = [schema.index_of_key(k) for k in spec.partition_by]
        self._schema = schema.extract(spec.partition_by)
        self._physical_partition_no = physical_partition_no
        self._partition_no = 0
        self._slice_no = 0
        self._item: Any = None

    def set(self, row: Any, partition_no: int, slice_no: int) -> None:
        self._item = (lambda: list(row())) if callable(row) else list(row)
        self._partition_no = partition_no
        self._slice_no = slice_no

    @property
    def item(self) -> Any:
        if callable(self._item):
            self._item = self._item()
        return self._item

    @property
    def row(self) -> List[Any]:
        return self.item

    @property
    def partition_no(self) -> int:
        return self._partition_no

    @property
    def physical_partition_no(self) -> int:
        return self._physical_partition_no

    @property
    def slice_no(self) -> int:
        return self._slice_no

    @property
    def row_schema(self) -> Schema:
        return self._orig_schema

    @property
    def key_schema(self) -> Schema:
        return self._schema

    @property
    def key_value_dict(self) -> Dict[str, Any]:
        return {self._orig_schema.names[i]: self.row[i] for i in self._key_index}

    @property
    def key_value_array(self) -> List[Any]:
        return [self.row[i] for i in self._key_index]

    def __getitem__(self, key: str) -> Any:
        return self.row[self._orig_schema.index_of_key(key)]
