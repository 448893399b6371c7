"""Drop-in for ``pathpyG.algorithms.lift_order`` (reference ``src/pathpyG/algorithms/lift_order.py``).

Same names, argument meaning, return fields, ordering and exceptions; the arithmetic runs in the
sm_100a kernels behind ``pathpyg_b200.ops``.
"""
from __future__ import annotations

import torch

from .. import _staging, ops
from ..core.data import Data, EdgeIndex
from ..core.graph import Graph

_PAIR_RULES = ("src", "dst", "max", "mul", "add")


def aggregate_node_attributes(edge_index: torch.Tensor, node_attribute: torch.Tensor, aggr: str = "src") -> torch.Tensor:
    """lift_order.py:10-45.  One value per edge from the attribute of its source / destination."""
    if aggr not in _PAIR_RULES:
        raise ValueError(f"Unknown aggregation method {aggr}")
    dev, to_host = _staging.compute_device(edge_index, node_attribute)
    out = ops.pair_attributes(_staging.up(edge_index, dev), _staging.up(node_attribute, dev), aggr)
    return _staging.down(out, to_host)


def lift_order_edge_index(edge_index: torch.Tensor, num_nodes: int | None = None) -> torch.Tensor:
    """lift_order.py:48-79.  Line graph of a row-sorted edge index."""
    dev, to_host = _staging.compute_device(edge_index)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    return _staging.down(ops.lift_order_edge_index(_staging.up(edge_index, dev), int(num_nodes)), to_host)


def lift_order_edge_index_weighted(edge_index: torch.Tensor, edge_weight: torch.Tensor, num_nodes: int | None = None,
                                   aggr: str = "src") -> tuple[torch.Tensor, torch.Tensor]:
    """lift_order.py:82-106."""
    if aggr not in _PAIR_RULES:
        raise ValueError(f"Unknown aggregation method {aggr}")
    dev, to_host = _staging.compute_device(edge_index, edge_weight)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    ho_index = ops.lift_order_edge_index(_staging.up(edge_index, dev), int(num_nodes))
    ho_weight = ops.pair_attributes(ho_index, _staging.up(edge_weight, dev), aggr, index_bound=int(edge_index.size(1)))
    return _staging.down(ho_index, to_host), _staging.down(ho_weight, to_host)


def aggregate_edge_index(edge_index: torch.Tensor, node_sequence: torch.Tensor, edge_weight: torch.Tensor | None = None,
                         aggr: str = "sum") -> Graph:
    """lift_order.py:109-152.  De Bruijn layer: distinct node sequences become the nodes (ids =
    lexicographic rank), edges are mapped through the inverse index and duplicates reduced."""
    dev, to_host = _staging.compute_device(edge_index, node_sequence, edge_weight)
    ei, ns, w = _staging.up(edge_index, dev), _staging.up(node_sequence, dev), _staging.up(edge_weight, dev)
    unique_nodes, inverse_idx = ops.unique_rows(ns)
    n = int(unique_nodes.size(0))
    # first order: the values of node_sequence are the ids themselves (lift_order.py:135-136)
    remap = ns.as_subclass(torch.Tensor).reshape(-1) if ns.size(1) == 1 else inverse_idx
    agg_index, agg_weight = ops.coalesce(ei, remap, n, w, aggr)
    data = Data(
        edge_index=EdgeIndex(_staging.down(agg_index, to_host), sparse_size=(n, n), sort_order="row"),
        num_nodes=n,
        node_sequence=_staging.down(unique_nodes, to_host),
        edge_weight=_staging.down(agg_weight, to_host),
        inverse_idx=_staging.down(inverse_idx, to_host),
    )
    return Graph._from_sorted(data)
