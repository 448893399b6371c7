"""``temporal_closeness_centrality`` (reference ``src/pathpyG/algorithms/centrality.py:303-324``): a consumer of
the shortest time-respecting path distances.  ``temporal_betweenness_centrality`` (``:164-300``, Brandes over the
event DAG with per-source path counts) is not built."""
from __future__ import annotations

from .. import _staging, ops


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
