"""Hop distances in a static graph (reference ``src/pathpyG/algorithms/shortest_paths.py``): scipy's Dijkstra with unit
edge lengths on the host-side adjacency matrix -- utilities around the containers, not part of the GPU path."""
from __future__ import annotations

import numpy as np
from scipy.sparse.csgraph import dijkstra

from ..core.graph import Graph


def _hops(graph: Graph, predecessors: bool):
    return dijkstra(graph.sparse_adj_matrix(), directed=graph.is_directed(), return_predecessors=predecessors, unweighted=True)


def shortest_paths_dijkstra(graph: Graph):
    """``(dist [n, n], pred [n, n])`` over all node pairs; unreachable pairs read ``inf`` / ``-9999``."""
    dist, pred = _hops(graph, True)
    return dist, pred


def diameter(graph: Graph) -> float:
    return np.max(_hops(graph, False))


def avg_path_length(graph: Graph) -> float:
    return np.sum(_hops(graph, False)) / (graph.n * (graph.n - 1))
