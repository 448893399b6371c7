"""CPU restatement of the shortest time-respecting paths (consumers of the event graph a1 produces).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference (relative to ``/root/reference``):
* ``src/pathpyG/algorithms/temporal.py:57-107``      temporal_shortest_paths
* ``src/pathpyG/algorithms/centrality.py:303-324``   temporal_closeness_centrality
* ``src/pathpyG/algorithms/centrality.py:164-300``   temporal_betweenness_centrality

``scipy.sparse.csgraph.dijkstra`` (scipy is a dependency of the reference and present here) is called exactly as
the reference calls it.  Distances are unique; the predecessor matrix is NOT: where several shortest paths reach a
node, scipy reports whichever predecessor its heap settles first, which is an implementation detail of scipy.
``is_valid_pred`` states the property every correct predecessor matrix has.
Pinned by the known answers of ``tests/algorithms/test_temporal.py:20-93`` and
``tests/algorithms/test_centrality.py:45-70`` (betweenness, closeness).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
from scipy.sparse.csgraph import dijkstra

from . import lift


def temporal_shortest_paths(edge_index: torch.Tensor, time: torch.Tensor, num_nodes: int, delta):
    """temporal.py:69-107 on the raw tensors of a TemporalGraph (time-sorted)."""
    m, n = edge_index.size(1), num_nodes
    event_graph = lift.lift_order_temporal(edge_index, time, delta)                                # :70
    src_edges = torch.stack([edge_index[0] + m, torch.arange(m)])                                  # :74-75,82
    dst_edges = torch.stack([torch.arange(m), edge_index[1] + m + n])                              # :77-78,83
    aug = torch.cat([event_graph, src_edges, dst_edges], dim=1).numpy()                            # :84
    adj = sp.coo_matrix((np.ones(aug.shape[1]), (aug[0], aug[1])), shape=(m + 2 * n, m + 2 * n))   # :87-88
    dist, pred = dijkstra(adj, directed=True, indices=np.arange(m, m + n), return_predecessors=True, unweighted=True)  # :93-95
    dist_fo = dist[:, m + n:] - 1                                                                  # :98
    np.fill_diagonal(dist_fo, 0)
    pred_fo = pred[:, n + m:]                                                                      # :102
    pred_fo[pred_fo == -9999] = -1
    idx_map = np.concatenate([edge_index[0].numpy(), [-1]])                                        # :104
    pred_fo = idx_map[pred_fo]
    np.fill_diagonal(pred_fo, np.arange(n))                                                        # :106
    return dist_fo, pred_fo


def temporal_closeness_centrality(dist: np.ndarray) -> np.ndarray:
    """centrality.py:320-322 as an array over node indices: sum over the other nodes x of (n - 1) / dist[x, v]."""
    n = dist.shape[0]
    return np.array([float(sum((n - 1) / dist[np.arange(n) != v, v])) for v in range(n)])


def is_valid_pred(edge_index: torch.Tensor, time: torch.Tensor, delta, dist: np.ndarray, pred: np.ndarray) -> bool:
    """Every entry names a node p with dist[s, p] + 1 == dist[s, v] such that an event (p -> v) ends a shortest
    time-respecting path from s (checked through the event-level hop counts), -1 exactly where v is unreachable
    and s itself on the diagonal."""
    n = dist.shape[0]
    src, dst = edge_index[0].numpy(), edge_index[1].numpy()
    try:
        eg = lift.lift_order_temporal(edge_index, time, delta).numpy()
    except RuntimeError:
        eg = np.empty((2, 0), dtype=np.int64)
    m = src.shape[0]
    for s in range(n):
        hops = np.full(m, np.inf)
        hops[src == s] = 1
        for e, f in eg.T:              # columns ascend in (e, f) and e < f in time order: one sweep is a relaxation in
            if hops[e] + 1 < hops[f]:  # topological order
                hops[f] = hops[e] + 1
        for v in range(n):
            if v == s:
                if pred[s, v] != s:
                    return False
            elif np.isinf(dist[s, v]):
                if pred[s, v] != -1:
                    return False
            else:
                ok = (dst == v) & (src == pred[s, v]) & (hops == dist[s, v])
                if not ok.any():
                    return False
    return True


def temporal_betweenness_centrality(edge_index: torch.Tensor, time: torch.Tensor, num_nodes: int, delta) -> np.ndarray:
    """centrality.py:164-300 (Brandes on the event DAG, Buss et al.) as an array over node indices; pure-Python
    loops in the reference's own order of operations -- small cases only."""
    from collections import defaultdict, deque
    from math import isnan

    m, n = edge_index.size(1), num_nodes
    event_graph = lift.lift_order_temporal(edge_index, time, delta)                                # :196
    src_edges = torch.stack([edge_index[0] + m, torch.arange(m)])                                  # :200-204
    aug = torch.cat([event_graph, src_edges], dim=1)                                               # :205
    order = torch.sort(aug[0], stable=True).indices                                                # Graph.__init__ row sort
    aug = aug[:, order]
    succ = defaultdict(list)
    for v, w in aug.t().tolist():
        succ[v].append(w)
    src_indices = torch.unique(src_edges[0]).tolist()                                              # :206
    e_i = edge_index.numpy()
    fo = lambda v: int(e_i[1, v]) if v < m else v - m                                              # noqa: E731  (:212-217)
    bw = defaultdict(float)
    for s in src_indices:                                                                          # :222
        delta_ = defaultdict(float)
        sigma = defaultdict(float)
        sigma[s] = 1.0
        sigma_fo = defaultdict(float)
        sigma_fo[fo(s)] = 1.0
        dist = defaultdict(lambda: -1)
        dist[s] = 0
        dist_fo = defaultdict(lambda: -1)
        dist_fo[fo(s)] = 0
        P = defaultdict(set)
        Q, S = deque([s]), []
        while Q:                                                                                   # :253-271
            v = Q.popleft()
            for w in succ[v]:
                if dist[w] == -1:
                    dist[w] = dist[v] + 1
                    if dist_fo[fo(w)] == -1:
                        dist_fo[fo(w)] = dist[v] + 1
                    S.append(w)
                    Q.append(w)
                if dist[w] == dist[v] + 1:
                    sigma[w] += sigma[v]
                    P[w].add(v)
                    if dist[w] == dist_fo[fo(w)]:
                        sigma_fo[fo(w)] += sigma[v]
        c = sum(1.0 for i in list(dist_fo) if dist_fo[i] >= 0)                                     # :273-276
        bw[fo(s)] = bw[fo(s)] - c + 1.0                                                            # :277
        while S:                                                                                   # :279-293
            w = S.pop()
            if dist[w] == dist_fo[fo(w)]:
                x = sigma[w] / sigma_fo[fo(w)]
                delta_[w] += 0.0 if isnan(x) else x
            for v in P[w]:
                x = sigma[v] / sigma[w]
                x = 0.0 if isnan(x) else x
                delta_[v] += x * delta_[w]
                bw[fo(v)] += delta_[w] * x
    return np.array([float(bw[i]) if i in bw else 0.0 for i in range(n)])
