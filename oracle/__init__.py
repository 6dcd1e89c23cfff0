"""CPU oracle for the fugue-b200 hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of what the reference (fugue-project/fugue
@ 91648b2, v0.9.4) computes on the ``fa.transform() -> MapEngine.map_dataframe``
path and its two neighbours (``ExecutionEngine.join`` / ``aggregate``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``fugue_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to this code.

Parity pinning (see DESIGN.md section "Oracle"):

* ``hash_partition`` is pinned against (1) the reference's own known-answer
  test ``tests/fugue_dask/test_utils.py:106-108`` (literal vectors in
  ``tests/golden/reference_literals.json``), and (2) outputs of the third-party
  function the reference calls, ``pandas.util.hash_pandas_object`` (pandas
  3.0.2 in this image), committed as ``tests/golden/hash_vectors.npz`` by
  ``tests/golden/make_golden.py``.
* ``native_engine`` is pinned against the literal tables of the reference's
  conformance suite (``fugue_test/execution_suite.py:208-314, 366-543, 177-206``).

The reference package itself cannot be imported here (``triad``/``adagio``
are absent, no network), so it is not executed; the literals above are the
reference's own expected outputs.
"""
