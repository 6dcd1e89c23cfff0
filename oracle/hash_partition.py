"""Oracle: key tuple -> physical partition id, and the stable partition it induces.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference definition (the only in-tree hash -> partition function):

    fugue_dask/_utils.py:146-169  (_add_hash_index)
        pd.util.hash_pandas_object(df[cols], index=False).mod(num).astype(int)
    fugue_ray/_utils/dataframe.py:115-118   (same expression)
    fugue_dask/_utils.py:124-130  (_postprocess: for ct == num every hash id is
                                   its own physical partition)

``hash_pandas_object`` lives in the third-party dependency pandas (setup.py:33
pins ``pandas>=2.0.2``; this image has 3.0.2).  Its published algorithm for
fixed-width columns, restated here:

    pandas/core/util/hashing.py::_hash_ndarray
        bits  = column viewed as unsigned of its own width, zero-extended to u64
                (bool -> 0/1, datetime/timedelta -> their int64 ticks)
        h     = bits; h ^= h >> 30; h *= 0xBF58476D1CE4E5B9
                h ^= h >> 27; h *= 0x94D049BB133111EB; h ^= h >> 31
    pandas/core/util/hashing.py::combine_hash_arrays   (tuple combine, n columns)
        out = 0x345678; mult = 1000003
        for i, h_i: out ^= h_i; out *= mult; mult += 82520 + 2 * (n - i)
        out += 97531                                        (all mod 2**64)

NULL keys (an extension of ours, the reference's numpy path has no validity
bitmap): a null cell contributes the bit pattern of the canonical float64 NaN
(0x7FF8000000000000) - i.e. exactly what pandas hashes for a missing double,
which is how the reference's own tests spell a NULL key
(fugue_test/execution_suite.py:218-234 uses ``a:double`` with ``None``).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

NULL_KEY_BITS = np.uint64(0x7FF8000000000000)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)


def key_bits(col: np.ndarray) -> np.ndarray:
    """Column -> u64 bit patterns exactly as pandas' ``_hash_ndarray`` does."""
    col = np.ascontiguousarray(col)
    if col.dtype == np.bool_:
        return col.astype("u8")
    if col.dtype.kind in "mM":
        return col.view("i8").astype("u8")
    if col.dtype.kind in "iuf" and col.dtype.itemsize <= 8:
        return col.view(f"u{col.dtype.itemsize}").astype("u8")
    raise NotImplementedError(f"oracle hashes fixed-width columns only, got {col.dtype}")


def fmix64(v: np.ndarray) -> np.ndarray:
    v = v.astype("u8", copy=True)
    with np.errstate(over="ignore"):
        v ^= v >> np.uint64(30)
        v *= _C1
        v ^= v >> np.uint64(27)
        v *= _C2
        v ^= v >> np.uint64(31)
    return v


def row_hash(
    cols: Sequence[np.ndarray], valid: Optional[Sequence[Optional[np.ndarray]]] = None
) -> np.ndarray:
    """u64 hash of each row's key tuple (``hash_pandas_object(df[cols], index=False)``)."""
    n = len(cols)
    assert n >= 1
    out = np.full(len(cols[0]), 0x345678, dtype="u8")
    mult = 1000003
    with np.errstate(over="ignore"):
        for i, c in enumerate(cols):
            b = key_bits(c)
            if valid is not None and valid[i] is not None:
                b = np.where(np.asarray(valid[i]).astype(bool), b, NULL_KEY_BITS)
            out ^= fmix64(b)
            out *= np.uint64(mult)
            mult = (mult + 82520 + 2 * (n - i)) & 0xFFFFFFFFFFFFFFFF
        out += np.uint64(97531)
    return out


def partition_ids(
    cols: Sequence[np.ndarray],
    num: int,
    valid: Optional[Sequence[Optional[np.ndarray]]] = None,
) -> np.ndarray:
    """``hash % num`` as int64 in [0, num)."""
    assert num >= 1
    return (row_hash(cols, valid) % np.uint64(num)).astype("i8")


def stable_partition(
    pids: np.ndarray, num: int
) -> Tuple[np.ndarray, np.ndarray]:
    """Return (order, offsets): ``order`` is the stable permutation that makes rows
    with equal partition id contiguous in ascending id order while keeping the
    input order inside each partition; ``offsets`` has ``num + 1`` entries."""
    order = np.argsort(pids, kind="stable")
    counts = np.bincount(pids, minlength=num).astype("i8")
    offsets = np.zeros(num + 1, dtype="i8")
    np.cumsum(counts, out=offsets[1:])
    return order, offsets


def partition_table(
    cols: List[np.ndarray],
    key_idx: Sequence[int],
    num: int,
    valid: Optional[Sequence[Optional[np.ndarray]]] = None,
) -> Tuple[List[np.ndarray], np.ndarray]:
    """Hash-partition a columnar table (list of equal-length arrays)."""
    kv = None if valid is None else [valid[i] for i in key_idx]
    pids = partition_ids([cols[i] for i in key_idx], num, kv)
    order, offsets = stable_partition(pids, num)
    return [np.ascontiguousarray(c[order]) for c in cols], offsets
