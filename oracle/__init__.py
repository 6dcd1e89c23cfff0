"""CPU oracle for the fugue-b200 hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of what the reference (fugue-project/fugue
@ 91648b2, v0.9.4) computes on the ``fa.transform() -> MapEngine.map_dataframe``
path and its neighbours (``ExecutionEngine.join`` / ``aggregate`` / ``select`` /
``filter`` / ``assign``).

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
* ``expressions`` (select / filter / assign / expression aggregates, restating the
  SQL that ``fugue/column/sql.py:275-347`` generates for qpd) is pinned against the
  literal tables of ``fugue_test/execution_suite.py:85-206`` (test_filter,
  test_select, test_assign, test_aggregate) in ``tests/test_oracle_expressions.py``;
  the column DSL it walks is pinned against the expected strings / types of the
  reference's ``tests/fugue/column/*.py`` in ``tests/test_column_dsl.py`` and against
  golden vectors produced by running the reference's own ``fugue/column`` modules in the
  build container (``tests/golden/make_column_golden.py`` ->
  ``tests/golden/column_dsl_vectors.json``, replayed by ``tests/test_column_golden.py``).

The reference package itself cannot be imported here (``triad``/``adagio``
are absent, no network), so it is not executed; the literals above are the
reference's own expected outputs.
"""
