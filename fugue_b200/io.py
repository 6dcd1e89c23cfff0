"""Host file IO either side of the device path: parquet / csv / json <-> a local Arrow frame.

Restates the behaviour of the reference's ``fugue/_utils/io.py`` (``FileParser`` :17-105, ``load_df`` :107-122,
``save_df`` :125-143, the per-format loaders :152-301) on pyarrow readers instead of pandas: the engine's
``load_df`` copies the Arrow columns to the device right after (fugue_b200/relational.py), so a pandas round
trip would only add a host copy.  What is kept from the reference:

* the format comes from ``format_hint`` or from the file suffix (``.csv``, ``.csv.gz``, ``.parquet``, ``.json``,
  ``.json.gz``); anything else is a NotImplementedError;
* a path may be one file, a list of files, a directory of part files, or a glob pattern;
* ``columns`` is None (everything), a list of names (projection) or a schema expression (projection + cast);
* csv: ``header`` decides whether the first line holds the names; without a header ``columns`` must be given;
  without ``infer_schema`` every field is read as text (then cast if ``columns`` is a schema);
  ``infer_schema`` together with a schema is a ValueError;
* ``save_df``: modes ``overwrite`` (replaces a file OR a directory) and ``error`` (FileExistsError), parent
  directories are created, csv is written without a header unless ``header=True``, json as one record per line.

Local file systems only (the reference goes through fsspec; remote stores are outside the hot path).
"""
import glob as _glob
import os
import shutil
from typing import Any, Dict, Iterable, List, Optional

import pyarrow as pa

from .dataframe import ArrowDataFrame, DataFrame, LocalDataFrame, _cast_table
from .schema import Schema

_SUFFIX_FORMATS: Dict[str, str] = {".csv": "csv", ".csv.gz": "csv", ".parquet": "parquet", ".json": "json",
                                   ".json.gz": "json"}
_PART_PATTERNS: Dict[str, List[str]] = {"csv": ["*.csv", "*.csv.gz"], "parquet": ["*.parquet"],
                                        "json": ["*.json", "*.json.gz"]}


class FilePath:
    """One ``uri`` of ``load_df`` / ``save_df``: its format, whether it is a pattern or a directory, and the
    files it stands for (fugue/_utils/io.py:17-105)."""

    def __init__(self, uri: str, format_hint: Optional[str] = None):
        self.raw_path = uri
        path = uri[len("file://"):] if uri.startswith("file://") else uri
        if "://" in path:
            raise NotImplementedError(f"only local paths are supported: {uri}")
        self.has_glob = "*" in path or "?" in path
        self.path = os.path.abspath(path)
        self.is_dir = not self.has_glob and os.path.isdir(self.path)
        name = os.path.basename(path.rstrip("/")).lower()
        self.suffix = name[name.index("."):] if "." in name else ""
        if format_hint is None or format_hint == "":
            found = [fmt for sfx, fmt in _SUFFIX_FORMATS.items() if self.suffix.endswith(sfx)]
            if not found:
                raise NotImplementedError(f"{self.suffix!r} is not a supported file suffix ({uri})")
            self.file_format = found[0]
        else:
            fmt = str(format_hint).lower().lstrip(".")
            if fmt not in _SUFFIX_FORMATS.values():
                raise NotImplementedError(f"{format_hint} is not supported")
            self.file_format = fmt

    def files(self) -> List[str]:
        """The data files behind this path, in name order.  A directory stands for its part files of the
        format (marker files like ``_SUCCESS`` are not data)."""
        if self.has_glob:
            return sorted(p for p in _glob.glob(self.path) if os.path.isfile(p))
        if not self.is_dir:
            return [self.path]
        found: List[str] = []
        for pat in _PART_PATTERNS[self.file_format]:
            found.extend(_glob.glob(os.path.join(self.path, pat)))
        if not found:  # part files without a suffix: everything that is not a marker / hidden file
            found = [os.path.join(self.path, f) for f in os.listdir(self.path)
                     if not f.startswith(("_", ".")) and os.path.isfile(os.path.join(self.path, f))]
        return sorted(set(found))


def _names_and_schema(columns: Any) -> Any:
    if columns is None:
        return None, None
    if isinstance(columns, (list, tuple)):
        return [str(c) for c in columns], None
    schema = columns if isinstance(columns, Schema) else Schema(columns)
    return schema.names, schema


def _project(table: pa.Table, names: Optional[List[str]], schema: Optional[Schema]) -> pa.Table:
    if names is None:
        return table
    missing = [n for n in names if n not in table.schema.names]
    if missing:
        raise KeyError(f"{missing} not in {table.schema.names}")
    table = table.select(names)
    return table if schema is None else _cast_table(table, schema)


def _read_parquet(path: str, names: Optional[List[str]], schema: Optional[Schema], kwargs: Dict[str, Any]) -> pa.Table:
    import pyarrow.parquet as pq

    table = pq.read_table(path, columns=names, **kwargs)
    return _project(table, names, schema)


def _read_csv(path: str, names: Optional[List[str]], schema: Optional[Schema], kwargs: Dict[str, Any]) -> pa.Table:
    import pyarrow.csv as pcsv

    kw = dict(kwargs)
    infer = bool(kw.pop("infer_schema", False))
    header = kw.pop("header", False)
    if infer and schema is not None:
        raise ValueError("can't set columns as a schema when infer schema is true")
    with_header = str(header) in ("True", "0")
    if not with_header and not (header is None or str(header) == "False"):
        raise NotImplementedError(f"header={header} is not supported")
    if not with_header and names is None:
        raise ValueError("columns must be set if without header")
    parse = pcsv.ParseOptions(delimiter=kw.pop("sep", kw.pop("delimiter", ",")))
    if kw:
        raise NotImplementedError(f"csv options {sorted(kw)} are not supported")
    if with_header:
        with pcsv.open_csv(path, parse_options=parse) as reader:
            file_names = list(reader.schema.names)
        read = pcsv.ReadOptions()
    else:
        file_names = list(names or [])
        read = pcsv.ReadOptions(column_names=file_names)
    convert = pcsv.ConvertOptions(strings_can_be_null=True) if infer else pcsv.ConvertOptions(
        column_types={n: pa.string() for n in file_names}, strings_can_be_null=True, quoted_strings_can_be_null=False)
    table = pcsv.read_csv(path, read_options=read, parse_options=parse, convert_options=convert)
    return _project(table, names, schema) if with_header else (table if schema is None else _cast_table(table, schema))


def _read_json(path: str, names: Optional[List[str]], schema: Optional[Schema], kwargs: Dict[str, Any]) -> pa.Table:
    import pyarrow.json as pjson

    if kwargs:
        raise NotImplementedError(f"json options {sorted(kwargs)} are not supported")
    return _project(pjson.read_json(path), names, schema)


_READERS = {"parquet": _read_parquet, "csv": _read_csv, "json": _read_json}


def load_df(uri: Any, format_hint: Optional[str] = None, columns: Any = None, **kwargs: Any) -> LocalDataFrame:
    """File(s) -> one local frame (fugue/_utils/io.py:107-122)."""
    uris: Iterable[str] = [uri] if isinstance(uri, str) else list(uri)
    names, schema = _names_and_schema(columns)
    tables: List[pa.Table] = []
    for u in uris:
        fp = FilePath(u, format_hint)
        files = fp.files()
        if not files:
            raise FileNotFoundError(u)
        for f in files:
            tables.append(_READERS[fp.file_format](f, names, schema, kwargs))
    if len(tables) == 0:
        raise FileNotFoundError(str(uri))
    table = tables[0] if len(tables) == 1 else pa.concat_tables(tables, promote_options="default")
    return ArrowDataFrame(table, schema)


def save_df(df: DataFrame, uri: str, format_hint: Optional[str] = None, mode: str = "overwrite",
            **kwargs: Any) -> None:
    """A local frame -> one file (fugue/_utils/io.py:125-143)."""
    if mode not in ("overwrite", "error"):
        raise NotImplementedError(f"{mode} is not supported")
    fp = FilePath(uri, format_hint)
    if fp.has_glob:
        raise AssertionError(f"{uri} has glob pattern")
    if os.path.lexists(fp.path):
        if mode == "error":
            raise FileExistsError(uri)
        if os.path.isdir(fp.path) and not os.path.islink(fp.path):
            shutil.rmtree(fp.path)
        else:
            os.remove(fp.path)
    parent = os.path.dirname(fp.path)
    if parent:
        os.makedirs(parent, exist_ok=True)
    table = df.as_arrow()
    if fp.file_format == "parquet":
        import pyarrow.parquet as pq

        pq.write_table(table, fp.path, **kwargs)
    elif fp.file_format == "csv":
        import pyarrow.csv as pcsv

        kw = dict(kwargs)
        header = bool(kw.pop("header", False))
        if kw:
            raise NotImplementedError(f"csv options {sorted(kw)} are not supported")
        textual = any(pa.types.is_string(t) or pa.types.is_large_string(t) or pa.types.is_binary(t) or pa.types.is_nested(t)
                      for t in table.schema.types)
        if not textual:   # numbers, booleans, dates: Arrow's writer, no quoting question
            opts = pcsv.WriteOptions(include_header=header, quoting_style="none")
            if fp.suffix.endswith(".gz"):
                with pa.CompressedOutputStream(fp.path, "gzip") as out:
                    pcsv.write_csv(table, out, opts)
            else:
                pcsv.write_csv(table, fp.path, opts)
        else:             # text columns: quote only what needs it (what pandas.to_csv, the reference's writer, does)
            import csv
            import gzip

            opener = gzip.open if fp.suffix.endswith(".gz") else open
            with opener(fp.path, "wt", newline="") as out:  # type: ignore
                w = csv.writer(out, quoting=csv.QUOTE_MINIMAL, lineterminator="\n")
                if header:
                    w.writerow(table.schema.names)
                for batch in table.to_batches(max_chunksize=1 << 16):
                    cols = [c.to_pylist() for c in batch.columns]
                    w.writerows(["" if v is None else v for v in r] for r in zip(*cols))
    else:
        import gzip
        import json

        if kwargs:
            raise NotImplementedError(f"json options {sorted(kwargs)} are not supported")
        rows = ArrowDataFrame(table).as_dicts()
        opener = gzip.open if fp.suffix.endswith(".gz") else open
        with opener(fp.path, "wt") as out:  # type: ignore
            for r in rows:
                out.write(json.dumps(r, default=str) + "\n")
