"""Which host inputs take the pipelined host -> device -> host route (``streaming._eligible``): fixed-width,
NULL-free columns with a keyed hash / default spec without presort; everything else goes step by step through
the engine (same results, no overlap).  Host logic, CPU only."""
import pyarrow as pa
import pytest

from fugue_b200.partition import PartitionSpec
from fugue_b200.schema import Schema
from fugue_b200.streaming import _eligible

GOOD = pa.table({"key": pa.array([1, 2, 3], pa.int64()), "v": pa.array([1.0, 2.0, 3.0], pa.float64()),
                 "d": pa.array([1, 2, 3], pa.int32()), "t": pa.array([1, 2, 3], pa.timestamp("us"))})
HASH = PartitionSpec(by=["key"], algo="hash", num=4)


def _ok(table, spec=HASH):
    return _eligible(table, Schema(table.schema), spec)


def test_plain_fixed_width_tables_qualify():
    assert _ok(GOOD) and _ok(GOOD, PartitionSpec(by=["key"])) and _ok(GOOD, PartitionSpec(by=["key", "d"], num=7))
    chunked = pa.concat_tables([GOOD, GOOD])          # several chunks per column are fine (copied slice by slice)
    assert chunked.column("key").num_chunks == 2 and _ok(chunked)


@pytest.mark.parametrize("why,table,spec", [
    ("no keys", GOOD, PartitionSpec(num=4)),
    ("coarse partitioning", GOOD, PartitionSpec(by=["key"], algo="coarse")),
    ("even spreads distinct keys, not hashes", GOOD, PartitionSpec(by=["key"], algo="even", num=2)),
    ("rand likewise", GOOD, PartitionSpec(by=["key"], algo="rand", num=2)),
    ("presort", GOOD, PartitionSpec(by=["key"], presort="v desc")),
    ("empty input", GOOD.slice(0, 0), HASH),
    ("a NULL", pa.table({"key": pa.array([1, None], pa.int64())}), HASH),
    ("a string column", pa.table({"key": pa.array([1, 2], pa.int64()), "s": pa.array(["a", "b"])}), HASH),
    ("a boolean column", pa.table({"key": pa.array([1, 2], pa.int64()), "b": pa.array([True, False])}), HASH),
    ("a nested column", pa.table({"key": pa.array([1, 2], pa.int64()), "l": pa.array([[1], [2]])}), HASH),
    ("a decimal column", pa.table({"key": pa.array([1, 2], pa.int64()),
                                   "m": pa.array([1, 2], pa.decimal128(5, 0))}), HASH),
])
def test_everything_else_takes_the_plain_route(why, table, spec):
    assert not _ok(table, spec), why


def test_declared_schema_must_match_the_stored_types():
    # a frame whose schema says long but whose Arrow column is int32 needs a cast first: not this route
    assert not _eligible(GOOD, Schema("key:long,v:double,d:long,t:datetime"), HASH)
