"""A small, dependency-free ``Schema`` with the expression syntax Fugue uses.

The reference takes ``Schema`` from the third-party ``triad`` package
(``triad.collections.schema.Schema``, used all over e.g.
fugue/dataframe/dataframe.py:42-52, fugue/extensions/transformer/convert.py:357-364),
which is not installed in this image.  This class restates the subset the hot path
needs: ``"a:int,b:str"`` expressions, ``"*"`` transforms, extract/exclude/union,
equality, and the pyarrow view.
"""
from typing import Any, Dict, Iterable, List, Optional, Tuple, Union

import pyarrow as pa

_TYPE_ALIASES: Dict[str, pa.DataType] = {
    "bool": pa.bool_(), "boolean": pa.bool_(),
    "int8": pa.int8(), "byte": pa.int8(),
    "int16": pa.int16(), "short": pa.int16(),
    "int32": pa.int32(), "int": pa.int32(),
    "int64": pa.int64(), "long": pa.int64(),
    "uint8": pa.uint8(), "ubyte": pa.uint8(),
    "uint16": pa.uint16(), "ushort": pa.uint16(),
    "uint32": pa.uint32(), "uint": pa.uint32(),
    "uint64": pa.uint64(), "ulong": pa.uint64(),
    "float16": pa.float16(), "float32": pa.float32(), "float": pa.float32(),
    "float64": pa.float64(), "double": pa.float64(),
    "str": pa.string(), "string": pa.string(),
    "bytes": pa.binary(), "binary": pa.binary(),
    "date": pa.date32(), "datetime": pa.timestamp("us"),
}
# canonical names, triad style (int -> "int", int64 -> "long", float64 -> "double" ...)
_CANONICAL: List[Tuple[pa.DataType, str]] = [
    (pa.bool_(), "bool"), (pa.int8(), "byte"), (pa.int16(), "short"), (pa.int32(), "int"),
    (pa.int64(), "long"), (pa.uint8(), "ubyte"), (pa.uint16(), "ushort"), (pa.uint32(), "uint"),
    (pa.uint64(), "ulong"), (pa.float16(), "float16"), (pa.float32(), "float"),
    (pa.float64(), "double"), (pa.string(), "str"), (pa.binary(), "bytes"), (pa.date32(), "date"),
]


class SchemaError(Exception):
    pass


def parse_type(expr: str) -> pa.DataType:
    """One type expression -> pyarrow type.  Flat names (``int``, ``long``, ``str``, ``datetime`` ...),
    ``timestamp(unit[,tz])``, ``decimal(p[,s])``, ``null`` and the nested forms ``[T]`` (list),
    ``{name:T,...}`` (struct) and ``<K,V>`` (map), recursively.  Nested columns live in host frames only
    (``B200Table.from_arrow`` refuses them)."""
    raw = expr.strip()
    e = raw.lower()
    if e in _TYPE_ALIASES:
        return _TYPE_ALIASES[e]
    if e == "null":
        return pa.null()
    if e.startswith("timestamp(") and e.endswith(")"):
        args = [x.strip() for x in raw[10:-1].split(",")]
        return pa.timestamp(args[0].lower(), args[1] if len(args) > 1 else None)
    if e.startswith("decimal(") and e.endswith(")"):
        args = [int(x) for x in e[8:-1].split(",")]
        return pa.decimal128(args[0], args[1] if len(args) > 1 else 0)
    if len(raw) >= 2:
        inner = raw[1:-1]
        if raw[0] == "[" and raw[-1] == "]":
            return pa.list_(parse_type(inner))
        if raw[0] == "{" and raw[-1] == "}":
            return pa.struct(Schema(inner).fields)
        if raw[0] == "<" and raw[-1] == ">":
            kv = _split_top_level(inner)
            if len(kv) != 2:
                raise SchemaError(f"a map type needs a key and a value type: {expr!r}")
            return pa.map_(parse_type(kv[0]), parse_type(kv[1]))
    raise SchemaError(f"unsupported type expression {expr!r}")


def type_to_expr(tp: pa.DataType) -> str:
    for t, name in _CANONICAL:
        if tp == t:
            return name
    if pa.types.is_timestamp(tp):
        if tp.unit == "us" and tp.tz is None:
            return "datetime"
        return f"timestamp({tp.unit},{tp.tz})" if tp.tz else f"timestamp({tp.unit})"
    if pa.types.is_large_string(tp):
        return "str"
    if pa.types.is_null(tp):
        return "null"
    if pa.types.is_decimal(tp):
        return f"decimal({tp.precision},{tp.scale})"
    if pa.types.is_map(tp):
        return f"<{type_to_expr(tp.key_type)},{type_to_expr(tp.item_type)}>"
    if pa.types.is_list(tp) or pa.types.is_large_list(tp):
        return f"[{type_to_expr(tp.value_type)}]"
    if pa.types.is_struct(tp):
        return "{" + ",".join(f"{tp.field(i).name}:{type_to_expr(tp.field(i).type)}"
                              for i in range(tp.num_fields)) + "}"
    raise SchemaError(f"unsupported arrow type {tp}")


class Schema:
    """Ordered ``name -> pyarrow type`` mapping with Fugue's schema expressions."""

    def __init__(self, *args: Any):
        self._fields: List[pa.Field] = []
        for a in args:
            self._append(a)
        names = [f.name for f in self._fields]
        if len(names) != len(set(names)):
            raise SchemaError(f"duplicated column names in {names}")

    # ---- construction -----------------------------------------------------------------
    def _append(self, obj: Any) -> None:
        if obj is None:
            return
        if isinstance(obj, Schema):
            self._fields.extend(obj._fields)
        elif isinstance(obj, pa.Schema):
            self._fields.extend(list(obj))
        elif isinstance(obj, pa.Field):
            self._fields.append(obj)
        elif isinstance(obj, str):
            for part in _split_top_level(obj):
                part = part.strip()
                if part == "":
                    continue
                if ":" not in part:
                    raise SchemaError(f"invalid schema expression {obj!r}")
                name, tp = part.split(":", 1)
                name = name.strip().strip("`")
                if name == "":
                    raise SchemaError(f"invalid schema expression {obj!r}")
                self._fields.append(pa.field(name, parse_type(tp)))
        elif isinstance(obj, dict):
            for k, v in obj.items():
                self._fields.append(pa.field(k, v if isinstance(v, pa.DataType) else parse_type(str(v))))
        elif isinstance(obj, (list, tuple)):
            for x in obj:
                if isinstance(x, tuple) and len(x) == 2:
                    self._fields.append(
                        pa.field(x[0], x[1] if isinstance(x[1], pa.DataType) else parse_type(str(x[1]))))
                else:
                    self._append(x)
        else:
            raise SchemaError(f"can't build a schema from {type(obj)}")

    # ---- views ------------------------------------------------------------------------
    @property
    def names(self) -> List[str]:
        return [f.name for f in self._fields]

    @property
    def fields(self) -> List[pa.Field]:
        return list(self._fields)

    @property
    def types(self) -> List[pa.DataType]:
        return [f.type for f in self._fields]

    @property
    def pa_schema(self) -> pa.Schema:
        return pa.schema(self._fields)

    def __len__(self) -> int:
        return len(self._fields)

    def __iter__(self):
        return iter(self.names)

    def __contains__(self, key: Any) -> bool:
        if isinstance(key, str):
            if ":" in key:
                return all(self._has_field(f) for f in Schema(key)._fields)
            return key in self.names
        if isinstance(key, (list, tuple, set)):
            return all(k in self for k in key)
        if isinstance(key, Schema):
            return all(self._has_field(f) for f in key._fields)
        return False

    def _has_field(self, f: pa.Field) -> bool:
        return any(x.name == f.name and x.type == f.type for x in self._fields)

    def __getitem__(self, name: Union[str, int]) -> pa.Field:
        if isinstance(name, int):
            return self._fields[name]
        return self._fields[self.index_of_key(name)]

    def index_of_key(self, name: str) -> int:
        try:
            return self.names.index(name)
        except ValueError:
            raise KeyError(f"{name} not in {self}") from None

    def __eq__(self, other: Any) -> bool:
        if other is None:
            return False
        if not isinstance(other, Schema):
            try:
                other = Schema(other)
            except Exception:
                return False
        return len(self) == len(other) and all(
            a.name == b.name and a.type == b.type for a, b in zip(self._fields, other._fields))

    def __ne__(self, other: Any) -> bool:
        return not self.__eq__(other)

    def __hash__(self) -> int:
        return hash(str(self))

    def __repr__(self) -> str:
        return ",".join(f"{f.name}:{type_to_expr(f.type)}" for f in self._fields)

    __str__ = __repr__

    def is_like(self, other: Any, equal_groups: Any = None) -> bool:
        return self == other

    # ---- algebra ----------------------------------------------------------------------
    def __add__(self, other: Any) -> "Schema":
        return Schema(self, other)

    def union(self, other: Any) -> "Schema":
        o = other if isinstance(other, Schema) else Schema(other)
        res = Schema(self)
        for f in o._fields:
            if f.name in res.names:
                if res[f.name].type != f.type:
                    raise SchemaError(f"{f} conflicts with {res[f.name]}")
            else:
                res._fields.append(f)
        return res

    def extract(self, names: Any) -> "Schema":
        if isinstance(names, str):
            names = [names]
        return Schema([self[n] for n in names])

    def exclude(self, names: Any) -> "Schema":
        if isinstance(names, str):
            names = [names]
        if isinstance(names, Schema):
            names = names.names
        ex = set(names)
        return Schema([f for f in self._fields if f.name not in ex])

    def intersect(self, names: Any) -> "Schema":
        if isinstance(names, Schema):
            names = names.names
        keep = set(names)
        return Schema([f for f in self._fields if f.name in keep])

    def rename(self, columns: Dict[str, str]) -> "Schema":
        for k in columns:
            if k not in self.names:
                raise SchemaError(f"{k} not in {self}")
        return Schema([pa.field(columns.get(f.name, f.name), f.type) for f in self._fields])

    def alter(self, sub: Any) -> "Schema":
        sub = sub if isinstance(sub, Schema) else Schema(sub)
        for f in sub._fields:
            if f.name not in self.names:
                raise SchemaError(f"{f.name} not in {self}")
        m = {f.name: f for f in sub._fields}
        return Schema([m.get(f.name, f) for f in self._fields])

    def transform(self, *exprs: Any, **kwargs: Any) -> "Schema":
        """``"*"`` keeps everything, ``"*,c:int"`` appends, ``"*-a,b"``/``"*~a"`` drop columns,
        an explicit expression replaces (triad ``Schema.transform`` semantics used by
        fugue/extensions/transformer/convert.py:357-364)."""
        res = Schema()
        for e in exprs:
            if e is None:
                continue
            if callable(e):
                res = res + e(self)
                continue
            if isinstance(e, Schema):
                res = res + e
                continue
            if not isinstance(e, str):
                res = res + Schema(e)
                continue
            for piece in _split_top_level(e):
                piece = piece.strip()
                if piece == "":
                    continue
                if piece.startswith("*"):
                    cur = Schema(self)
                    rest = piece[1:]
                    while rest:
                        op = rest[0]
                        if op not in "-~":
                            raise SchemaError(f"invalid schema transform {e!r}")
                        j = 1
                        while j < len(rest) and rest[j] not in "-~":
                            j += 1
                        name = rest[1:j].strip()
                        if op == "-" and name not in cur.names:
                            raise SchemaError(f"{name} not in {cur}")
                        cur = cur.exclude([name])
                        rest = rest[j:]
                    res = res + cur
                else:
                    res = res + Schema(piece)
        if kwargs:
            res = res + Schema({k: v for k, v in kwargs.items()})
        return res


def _split_top_level(expr: str) -> List[str]:
    out, depth, cur = [], 0, []
    for ch in expr:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out
