"""CPU oracle for the pathpyG lift -> DBGNN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pathpyg_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or as the
CPU arm that is timed next to the GPU path.

What is restated here (reference paths are relative to ``/root/reference``):

* ``oracle/pyg.py``      torch_geometric 2.7.0 utilities the path calls
  (``degree``, ``cumsum``, ``coalesce``, ``scatter``, ``add_remaining_self_loops``,
  ``gcn_norm``) -- third-party, pinned in ``uv.lock:5910-5912``, absent here.
* ``oracle/lift.py``     ``src/pathpyG/algorithms/lift_order.py`` and
  ``src/pathpyG/algorithms/temporal.py:17-54``.
* ``oracle/mom.py``      the four ``MultiOrderModel`` builders
  (``src/pathpyG/core/multi_order_model.py:83-241,511-554``),
  ``PathData.append_walks`` (``core/path_data.py:126-159``) and
  ``generate_bipartite_edge_index`` (``utils/dbgnn.py:10-46``).
* ``oracle/dbgnn.py``    ``src/pathpyG/nn/dbgnn.py`` + PyG ``GCNConv``.
* ``oracle/selection.py`` ``Graph.degrees`` / ``transition_probabilities`` (``core/graph.py:486-533``) and the
  model-selection statistics of ``MultiOrderModel`` (``core/multi_order_model.py:243-509``).
* ``oracle/ingest.py``   ``df_to_temporal_graph`` / ``read_csv_path_data`` (``io/pandas.py``).
* ``oracle/paths.py``    ``temporal_shortest_paths`` (``algorithms/temporal.py:57-107``),
  ``temporal_closeness_centrality`` and ``temporal_betweenness_centrality`` (``algorithms/centrality.py:164-324``).

* ``oracle/containers.py`` the edge-merging members of ``Graph`` / ``TemporalGraph``; ``oracle/ref_loader.py`` also runs the
  reference's own algorithm / io modules ON this package's containers (``reference_module_on``), and ``oracle/ref_suite.py``
  runs the reference's own TEST FILES against this package under an import alias.

Pinning status
--------------
* lift / indexing (a1-a9): PINNED.  ``tests/golden/make_golden.py`` executes the
  reference's own ``lift_order.py`` / ``lift_order_temporal`` source (with the
  four PyG utilities replaced by ``oracle/pyg.py``) in this container and
  commits the input/output vectors under ``tests/golden/``; the oracle is checked
  against those and against every known-answer vector in the reference's tests
  (SURVEY.md section 8c).
* statistics, ingest, shortest paths: PINNED the same way -- ``oracle/ref_loader.py`` compiles the
  reference's own method bodies / modules from ``/root/reference`` and ``tests/golden/make_selection_golden.py`` /
  ``make_ingest_golden.py`` commit their outputs; known answers of the reference's tests are test cases.
* DBGNN numerics (a10/a11): PARITY UNPINNED.  The reference holds no activation
  values (``tests/nn/test_dbgnn.py:33-43`` asserts ``out is not None``) and
  ``GCNConv`` lives in un-vendored PyG, so ``oracle/dbgnn.py`` is a restatement of
  PyG 2.7.0's published algorithm, not a checked copy.
"""
