"""Pins the oracle's hash -> partition function against golden vectors
(pandas.util.hash_pandas_object outputs) and the reference's known-answer test."""
import ctypes as C
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import hash_partition as hp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "hash_vectors.npz"), allow_pickle=False)
LIT = json.load(open(os.path.join(HERE, "golden", "reference_literals.json")))


@pytest.mark.parametrize("i", range(len(GOLD["combos"])))
def test_row_hash_matches_golden(i):
    names = str(GOLD["combos"][i]).split(",")
    got = hp.row_hash([GOLD[n] for n in names])
    assert np.array_equal(got, GOLD[f"hash_{i}"])


def test_row_hash_matches_live_pandas():
    # pandas is present on every box of this image: also check against it live
    rng = np.random.default_rng(7)
    df = pd.DataFrame({"a": rng.integers(-2**62, 2**62, 1000), "b": rng.integers(0, 9, 1000).astype("int32")})
    for cols in (["a"], ["b"], ["a", "b"]):
        ref = pd.util.hash_pandas_object(df[cols], index=False).to_numpy()
        assert np.array_equal(hp.row_hash([df[c].to_numpy() for c in cols]), ref)
        for num in (1, 3, 256, 1000):
            ref_pid = pd.util.hash_pandas_object(df[cols], index=False).mod(num).astype(int).to_numpy()
            assert np.array_equal(hp.partition_ids([df[c].to_numpy() for c in cols], num), ref_pid)


def test_reference_known_answer_hash_repartition():
    # tests/fugue_dask/test_utils.py:106-108: hash_repartition(df, 3, ["aa"])
    lit = LIT["test_hash_repartition"]
    aa = np.array(lit["aa"], dtype="int64")
    pids = hp.partition_ids([aa], lit["num"])
    buckets = sorted(sorted(aa[pids == p].tolist()) for p in np.unique(pids))
    assert buckets == lit["buckets_sorted"]
    # :110-112 num=1 -> one bucket with everything
    pids = hp.partition_ids([aa], 1)
    assert sorted(aa[pids == 0].tolist()) == lit["num1_bucket"][0]


def test_survey_known_answers():
    lit = LIT["survey_known_answers"]
    keys = np.array(lit["keys"], dtype="int64")
    assert hp.partition_ids([keys], lit["num"]).tolist() == lit["pids"]
    aa = np.array(LIT["test_hash_repartition"]["aa"], dtype="int64")
    assert hp.row_hash([aa]).tolist() == lit["raw_hash_aa"]


def test_null_key_hashes_as_nan():
    a = np.array([1.0, np.nan, 3.0])
    v = np.array([1, 0, 1], dtype="uint8")
    b = np.array([1.0, 123.0, 3.0])  # payload under the null is ignored
    assert np.array_equal(hp.row_hash([a]), hp.row_hash([b], [v]))


def test_stable_partition_properties():
    rng = np.random.default_rng(3)
    n, num = 5000, 17
    key = rng.integers(0, 100, n)
    val = np.arange(n)
    (k2, v2), off = hp.partition_table([key, val], [0], num)
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
    pid2 = hp.partition_ids([k2], num)
    for p in range(num):
        seg = slice(off[p], off[p + 1])
        assert np.all(pid2[seg] == p)
        assert np.all(np.diff(v2[seg]) > 0)  # stable: original order kept
    assert sorted(v2.tolist()) == list(range(n))


def _c_oracle():
    path = os.path.join(os.path.dirname(HERE), "oracle", "_build", "libfb_oracle.so")
    if not os.path.exists(path):
        pytest.skip("C oracle not built (run __graft_entry__.build())")
    return C.CDLL(path)


def test_c_oracle_matches_numpy_oracle():
    lib = _c_oracle()
    rng = np.random.default_rng(5)
    n, num = 20000, 256
    cols = [rng.integers(0, 1 << 16, n).astype("int64"), rng.integers(-5, 5, n).astype("int32"),
            rng.standard_normal(n), rng.integers(0, 255, n).astype("uint8")]
    widths = (C.c_int32 * 4)(*[c.dtype.itemsize for c in cols])
    outs = [np.empty_like(c) for c in cols]
    offs = np.zeros(num + 1, dtype="int64")
    src = (C.c_void_p * 4)(*[c.ctypes.data for c in cols])
    dst = (C.c_void_p * 4)(*[o.ctypes.data for o in outs])
    kidx = (C.c_int32 * 2)(0, 1)
    rc = lib.fbo_partition_cols(C.c_int64(n), 4, src, widths, 2, kidx, None, C.c_uint32(num), dst,
                                offs.ctypes.data_as(C.c_void_p))
    assert rc == 0
    exp, exp_off = hp.partition_table(cols, [0, 1], num)
    assert np.array_equal(offs, exp_off)
    for a, b in zip(outs, exp):
        assert np.array_equal(a.view("u1"), b.view("u1"))
