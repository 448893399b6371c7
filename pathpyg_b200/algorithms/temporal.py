"""Drop-in for ``pathpyG.algorithms.temporal.lift_order_temporal``
(reference ``src/pathpyG/algorithms/temporal.py:17-54``)."""
from __future__ import annotations

import torch

from .. import _staging, ops


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
    out = ops.lift_order_temporal(_staging.up(edge_index, dev), _staging.up(time, dev), delta, int(num_nodes))
    return _staging.down(out, to_host)
