"""Drop-ins for ``pathpyG.algorithms.temporal``: ``lift_order_temporal`` (reference
``src/pathpyG/algorithms/temporal.py:17-54``) and ``temporal_shortest_paths`` (``:57-107``)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, _staging, ops


def lift_order_temporal(g, delta: float | int = 1) -> torch.Tensor:
    """Second-order event graph of a temporal graph: edge (e -> f) between positions of the
    time-sorted edge list iff ``dst(e) == src(f)`` and ``t_e < t_f <= t_e + delta``; columns ascend
    in (e, f).  Raises ``RuntimeError`` when no such pair exists, like the reference's empty
    ``torch.cat`` (temporal.py:53)."""
    edge_index, time = g.data.edge_index, g.data.time
    dev, to_host = _staging.compute_device(edge_index, time)
    num_nodes = g.data.num_nodes
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    known = getattr(g, "time_is_known_sorted", lambda: False)()
    out = ops.lift_order_temporal(_staging.up(edge_index, dev), _staging.up(time, dev), delta, int(num_nodes), assume_sorted=known)
    return _staging.down(out, to_host)


def temporal_shortest_paths(g, delta: int):
    """Shortest time-respecting paths between all first-order nodes (temporal.py:57-107): ``(dist, pred)`` as
    numpy arrays like the reference.  ``dist[s, v]`` = fewest events on a time-respecting path from s to v (``inf``
    if none, 0 on the diagonal); ``pred[s, v]`` = the node before v on such a path (-1 if none).  The reference
    hands the event DAG to scipy's dijkstra once per source node; here all sources advance together in a
    bit-parallel breadth-first search on the GPU (``csrc/paths.cu``)."""
    edge_index, time = g.data.edge_index, g.data.time
    dev, _ = _staging.compute_device(edge_index, time)
    ei = _staging.up(edge_index, dev)
    n = int(g.data.num_nodes)
    try:
        event_graph = ops.lift_order_temporal(ei, _staging.up(time, dev), delta, n)
    except _lib.EmptyLiftError:
        if ei.size(1) == 0:
            raise
        raise  # the reference fails in lift_order_temporal as well when no pair exists (temporal.py:70)
    dist, pred = ops.temporal_paths(ei, event_graph, n)
    return dist.cpu().numpy(), pred.cpu().numpy()
