"""CPU restatement of the edge-merging members of the reference's containers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned by ``tests/golden/container_golden.npz``, which
``tests/golden/make_container_golden.py`` produced by executing the reference's OWN method bodies
(``oracle/ref_loader.container_methods``).

* ``graph_to_undirected``      ``src/pathpyG/core/graph.py:211-251``
* ``graph_to_weighted``        ``src/pathpyG/core/graph.py:253-270``
* ``temporal_to_static``       ``src/pathpyG/core/temporal_graph.py:191-220``
"""
from __future__ import annotations

import torch

from . import pyg


def _row_sorted(edge_index, *attrs):
    """``Graph.__init__`` (graph.py:103-105): stable sort by row, carried over to the edge attributes."""
    perm = torch.sort(edge_index[0], stable=True).indices
    return (edge_index[:, perm],) + tuple(a[perm] if a is not None else None for a in attrs)


def graph_to_undirected(edge_index: torch.Tensor, num_nodes: int, edge_attr: torch.Tensor | None = None):
    """-> (edge_index, edge_attr of the merged edges): the attribute of a merged edge is the one of its
    lowest-numbered original edge (``reduce="min"`` over the edge numbers, graph.py:227-233,249)."""
    number = torch.arange(edge_index.size(1))
    ei, attr_idx = pyg.to_undirected(edge_index, number, num_nodes, reduce="min")
    ei, attr_idx = _row_sorted(ei, attr_idx)
    return ei, (edge_attr[attr_idx] if edge_attr is not None else None), attr_idx


def graph_to_weighted(edge_index: torch.Tensor, num_nodes: int):
    """-> (edge_index, edge_weight = multiplicity), graph.py:266-270."""
    ei, w = pyg.coalesce(edge_index, torch.ones(edge_index.size(1)), num_nodes)
    return _row_sorted(ei, w)


def temporal_to_static(edge_index: torch.Tensor, time: torch.Tensor, weighted: bool, time_window=None):
    """-> (edge_index, edge_weight or None, n); ``n`` is the largest index present + 1 (temporal_graph.py:207)."""
    if time_window is not None:
        idx = ((time >= time_window[0]) & (time < time_window[1])).nonzero().ravel()
        edge_index = edge_index[:, idx]
    n = int(edge_index.max()) + 1
    if weighted:
        ei, w = pyg.coalesce(edge_index, torch.ones(edge_index.size(1)), n)
        ei, w = _row_sorted(ei, w)
        return ei, w, n
    return _row_sorted(edge_index)[0], None, n
