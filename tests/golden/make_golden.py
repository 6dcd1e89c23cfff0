"""Generates tests/golden/hash_vectors.npz and reference_literals.json.

Run in the build container:  python tests/golden/make_golden.py

* hash_vectors.npz : inputs + outputs of the function the reference calls for
  hash partitioning, ``pandas.util.hash_pandas_object(df[cols], index=False)``
  (fugue_dask/_utils.py:155-161), evaluated with the pandas of this image.
  pandas is a third-party dependency of the reference (setup.py:33); the
  reference itself cannot be imported here (triad/adagio missing).
* reference_literals.json : literal expectations copied from the reference's own
  tests (file:line recorded per entry).
"""
import json
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    rng = np.random.default_rng(20260921)
    n = 4096
    edge = np.array([0, 1, 2, 3, 4, -5, 2**62, -(2**63), 2**63 - 1, -1, 65535, 65536], dtype="int64")
    cols = {
        "k_i64": np.concatenate([edge, rng.integers(-(2**62), 2**62, n - len(edge))]).astype("int64"),
        "k_i32": rng.integers(-(2**31), 2**31 - 1, n).astype("int32"),
        "k_i16": rng.integers(-(2**15), 2**15 - 1, n).astype("int16"),
        "k_u8": rng.integers(0, 255, n).astype("uint8"),
        "k_bool": rng.integers(0, 2, n).astype(bool),
        "k_f64": np.concatenate([[0.0, -0.0, np.nan, np.inf, -np.inf], rng.standard_normal(n - 5)]),
        "k_f32": rng.standard_normal(n).astype("float32"),
        "k_small": rng.integers(0, 37, n).astype("int64"),
    }
    df = pd.DataFrame(cols)
    out = dict(cols)
    combos = [["k_i64"], ["k_i32"], ["k_i16"], ["k_u8"], ["k_bool"], ["k_f64"], ["k_f32"], ["k_small"],
              ["k_i64", "k_i32"], ["k_small", "k_u8", "k_bool"],
              ["k_i64", "k_i32", "k_i16", "k_u8", "k_bool", "k_f64", "k_f32", "k_small"]]
    names = []
    for i, c in enumerate(combos):
        h = pd.util.hash_pandas_object(df[c], index=False).to_numpy().astype("uint64")
        out[f"hash_{i}"] = h
        names.append(",".join(c))
    out["combos"] = np.array(names)
    out["pandas_version"] = np.array(pd.__version__)
    np.savez_compressed(os.path.join(HERE, "hash_vectors.npz"), **out)

    literals = {
        "test_hash_repartition": {
            "source": "tests/fugue_dask/test_utils.py:84-108",
            "aa": [0, 1, 1, 2, 3, 4],
            "num": 3,
            "by": ["aa"],
            "buckets_sorted": [[0, 2], [1, 1, 3, 4]],
            "num1_bucket": [[0, 1, 1, 2, 3, 4]],
        },
        "survey_known_answers": {
            "source": "SURVEY.md section 8 A3 (pandas 3.0.2 evaluated in the build container)",
            "keys": [0, 1, 2, 3, 4, -5, 2**62, -(2**63)],
            "num": 256,
            "pids": [99, 18, 81, 147, 63, 29, 114, 81],
            "raw_hash_aa": [3430018387555, 11358988112447789330, 11358988112447789330,
                            826468140851422801, 18319371940472138387, 1644348678988017215],
        },
    }
    with open(os.path.join(HERE, "reference_literals.json"), "w") as f:
        json.dump(literals, f, indent=1)
    print("wrote golden vectors with pandas", pd.__version__)


if __name__ == "__main__":
    main()
