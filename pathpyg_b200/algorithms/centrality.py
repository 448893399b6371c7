"""``temporal_closeness_centrality`` (reference ``src/pathpyG/algorithms/centrality.py:303-324``): a consumer of
the shortest time-respecting path distances, and ``temporal_betweenness_centrality`` (``:164-300``): Brandes'
dependency accumulation over the event DAG."""
from __future__ import annotations

import torch

from .. import _staging, ops


def path_node_traversals(paths) -> dict:
    """How often the walks of a ``PathData`` object pass each node (centrality.py:50-57): ``{node id: visits}``."""
    nodes, visits = torch.unique(paths.data.node_sequence, return_counts=True)
    return {paths.mapping.to_id(v): c for v, c in zip(nodes.tolist(), visits.tolist())}


def path_visitation_probabilities(paths) -> dict:
    """Share of all node visits that falls on each node (centrality.py:134-161)."""
    visits = path_node_traversals(paths)
    total = float(sum(visits.values()))
    return {v: c / total for v, c in visits.items()}


def map_to_nodes(graph, centralities: dict) -> dict:
    """Re-key ``{node index: value}`` by node id (centrality.py:60-76)."""
    return {graph.mapping.to_id(i): centralities[i] for i in centralities}


def temporal_closeness_centrality(graph, delta: int) -> dict:
    """closeness(v) = sum over the other nodes x of (n - 1) / dist[x, v] with dist from ``temporal_shortest_paths``;
    unreachable pairs contribute (n - 1) / inf = 0.  The column sums are formed on the GPU in float64, in the
    reference's order of addition."""
    edge_index, time = graph.data.edge_index, graph.data.time
    dev, _ = _staging.compute_device(edge_index, time)
    ei = _staging.up(edge_index, dev)
    n = int(graph.data.num_nodes)
    event_graph = ops.lift_order_temporal(ei, _staging.up(time, dev), delta, n)
    dist, _ = ops.temporal_paths(ei, event_graph, n)
    closeness = ops.temporal_closeness(dist).cpu().tolist()
    return {x: float(closeness[graph.mapping.to_idx(x)]) for x in graph.nodes}


def temporal_betweenness_centrality(graph, delta: int = 1) -> dict:
    """Temporal betweenness of the nodes over shortest time-respecting paths (path length = number of traversed
    edges, waiting time at most ``delta``), centrality.py:164-300.  The reference walks the event DAG in Python once
    per source node; here every source is one CTA sweeping the time groups of the event list forward (distances and
    path counts) and backward (dependencies).  Returns ``{node id: value}``; nodes that get no contribution read 0.0
    (the reference returns a ``defaultdict`` with the same values)."""
    from collections import defaultdict

    edge_index, time = graph.data.edge_index, graph.data.time
    dev, _ = _staging.compute_device(edge_index, time)
    ei, t = _staging.up(edge_index, dev), _staging.up(time, dev)
    n = int(graph.data.num_nodes)
    event_graph = ops.lift_order_temporal(ei, t, delta, n)
    bw = ops.temporal_betweenness(ei, t, event_graph, n).cpu().tolist()
    out = defaultdict(lambda: 0.0)
    for idx, value in enumerate(bw):
        out[graph.mapping.to_id(idx)] = float(value)
    return out
