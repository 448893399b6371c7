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


def betweenness_centrality(graph, sources: list | None = None) -> dict:
    """Unnormalised betweenness of a static graph over shortest (hop) paths, Brandes' accumulation as the reference
    states it (centrality.py:79-131; parallel edges count as separate paths, and a node's dependency is weighted by
    its number of shortest-path predecessor edges).
    Level-synchronous on the host: every BFS level and every dependency step is one sparse matrix-vector product.
    Nodes that no source reaches are absent from the returned ``defaultdict`` (they read 0.0)."""
    from collections import defaultdict

    import numpy as np

    n = graph.n
    adj = graph.sparse_adj_matrix().tocsr()          # duplicates summed: entry = number of parallel edges
    adj_t = adj.T.tocsr()
    total = np.zeros(n)
    touched = np.zeros(n, dtype=bool)
    for s in (range(n) if sources is None else [graph.mapping.to_idx(v) for v in sources]):
        level = np.full(n, -1)
        sigma = np.zeros(n)
        level[s], sigma[s] = 0, 1.0
        frontier, depth = np.zeros(n), 0
        frontier[s] = 1.0
        while True:
            reach = adj_t @ frontier                   # path counts handed to the next level
            new = (reach > 0) & (level < 0)
            if not new.any():
                break
            depth += 1
            level[new] = depth
            sigma[new] = reach[new]
            frontier = np.where(new, sigma, 0.0)
        delta = np.zeros(n)
        for d in range(depth - 1, -1, -1):
            carry = np.where(level == d + 1, (1.0 + delta) / np.where(sigma > 0, sigma, 1.0), 0.0)
            here = level == d
            delta[here] = sigma[here] * (adj @ carry)[here]
        reached = level > 0
        # the reference adds a node's dependency once per predecessor EDGE on a shortest path (the update sits inside
        # its loop over the predecessor list, centrality.py:126-129), not once per source as in Brandes' paper
        pred_edges = np.zeros(n)
        for d in range(1, depth + 1):
            here = level == d
            pred_edges[here] = (adj_t @ (level == d - 1).astype(float))[here]
        total[reached] += delta[reached] * pred_edges[reached]
        touched |= reached
    out = defaultdict(lambda: 0.0)
    for i in np.flatnonzero(touched):
        out[graph.mapping.to_id(int(i))] = float(total[i])
    return out


def __getattr__(name: str):
    """Any other ``*centrality*`` function is networkx's, applied to the graph's edges and re-keyed by node id
    (centrality.py:327-356)."""
    if name.startswith("__"):
        raise AttributeError(name)

    def delegate(*args, **kwargs):
        import networkx as nx

        from ..core.graph import Graph
        from ..core.temporal_graph import TemporalGraph

        if len(args) == 0:
            raise RuntimeError(f"Did not find method {name} with no arguments")
        if isinstance(args[0], TemporalGraph):
            raise NotImplementedError(f"Missing implementation of {name} for temporal graphs")
        if not isinstance(args[0], Graph):
            raise RuntimeError(f"Did not find method {name} that accepts first argument of type {type(args[0])}")
        g = nx.DiGraph()
        g.add_nodes_from(range(args[0].n))
        g.add_edges_from(args[0].data.edge_index.as_tensor().t().cpu().tolist())
        result = getattr(nx.algorithms.centrality, name)(g, *args[1:], **kwargs)
        return map_to_nodes(args[0], result) if "centrality" in name and isinstance(result, dict) else result

    return delegate


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
