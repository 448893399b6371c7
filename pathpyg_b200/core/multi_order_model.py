"""``MultiOrderModel`` builders (reference ``src/pathpyG/core/multi_order_model.py``):
``iterate_lift_order`` (:83-122), ``from_temporal_graph`` (:124-192), ``from_path_data`` (:194-241),
``to_dbgnn_data`` (:511-554).  The degrees-of-freedom / likelihood part (:243-509) consumes the
layers built here and is outside the hot path.

All tensors of one build stay on the GPU from the first kernel to the last; host inputs are staged
once at entry and the finished layers are moved back once at exit.
"""
from __future__ import annotations

import logging

import torch

from .. import _staging, ops
from ..utils.dbgnn import generate_bipartite_edge_index
from .data import Data, EdgeIndex
from .graph import Graph
from .index_map import HigherOrderIndexMap, IndexMap
from .path_data import PathData
from .temporal_graph import TemporalGraph

logger = logging.getLogger("root")


def _plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t


def _aggregate(edge_index, node_sequence, edge_weight, aggr="sum") -> Graph:
    """Device-side aggregate_edge_index (lift_order.py:109-152)."""
    unique_nodes, inverse_idx = ops.unique_rows(node_sequence)
    n = int(unique_nodes.size(0))
    remap = node_sequence.reshape(-1) if node_sequence.size(1) == 1 else inverse_idx
    agg_index, agg_weight = ops.coalesce(edge_index, remap, n, edge_weight, aggr)
    data = Data(edge_index=EdgeIndex(agg_index, sparse_size=(n, n), sort_order="row"), num_nodes=n,
                node_sequence=unique_nodes, edge_weight=agg_weight, inverse_idx=inverse_idx)
    return Graph._from_sorted(data)


def _extend_node_sequence(node_sequence, edge_index):
    """k-gram of every line-graph edge: the (k-1)-gram of its source + the last node of its target
    (multi_order_model.py:114,165)."""
    return torch.cat([node_sequence[edge_index[0]], node_sequence[edge_index[1]][:, -1:]], dim=1)


def _lift_step(edge_index, node_sequence, edge_weight, aggr, save):
    n_prev = int(node_sequence.size(0))
    ho_index = ops.lift_order_edge_index(edge_index, n_prev)
    if edge_weight is not None:
        edge_weight = ops.pair_attributes(ho_index, edge_weight, aggr)
    node_sequence = _extend_node_sequence(node_sequence, edge_index)
    gk = _aggregate(ho_index, node_sequence, edge_weight) if save else None
    return ho_index, node_sequence, edge_weight, gk


class MultiOrderModel:
    """Higher-order De Bruijn graph layers keyed by order (``layers: dict[int, Graph]``)."""

    def __init__(self) -> None:
        self.layers: dict[int, Graph] = {}

    def __str__(self) -> str:
        max_order = max(list(self.layers.keys())) if self.layers else 0
        return f"MultiOrderModel with max. order {max_order}"

    def to(self, device) -> "MultiOrderModel":
        for g in self.layers.values():
            g.to(device)
        return self

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def iterate_lift_order(edge_index, node_sequence, mapping: IndexMap, edge_weight=None, aggr: str = "src",
                           save: bool = True):
        """multi_order_model.py:83-122: lift by one order, extend the node sequences, aggregate."""
        if aggr not in ("src", "dst", "max", "mul", "add"):
            raise ValueError(f"Unknown aggregation method {aggr}")
        dev, to_host = _staging.compute_device(edge_index, node_sequence, edge_weight)
        ei, ns, w = (_plain(_staging.up(t, dev)) for t in (edge_index, node_sequence, edge_weight))
        ho_index, ns, w, gk = _lift_step(ei.long(), ns.long(), w, aggr, save)
        if gk is not None:
            if to_host:
                gk.to("cpu")
            gk.mapping = HigherOrderIndexMap(mapping, gk.data.node_sequence)
        return _staging.down(ho_index, to_host), _staging.down(ns, to_host), _staging.down(w, to_host), gk

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def from_temporal_graph(g: TemporalGraph, delta: float | int = 1, max_order: int = 1, weight: str = "edge_weight",
                            cached: bool = True, event_graph: torch.Tensor | None = None) -> "MultiOrderModel":
        """multi_order_model.py:124-192."""
        m = MultiOrderModel()
        data = g.data if g.data.is_sorted_by_time() else g.data.sort_by_time()
        dev, to_host = _staging.compute_device(data.edge_index, data.time)
        edge_index = _plain(_staging.up(data.edge_index, dev)).long()
        time = _staging.up(data.time, dev)
        n = int(data.num_nodes)
        node_sequence = torch.arange(n, device=dev).unsqueeze(1)
        edge_weight = _staging.up(data[weight], dev) if weight in data else None  # None == ones(m), :154-157

        if cached or max_order == 1:
            m.layers[1] = _aggregate(edge_index, node_sequence, edge_weight)
            m.layers[1].mapping = g.mapping

        if max_order > 1:
            node_sequence = _extend_node_sequence(node_sequence, edge_index)
            if event_graph is None:
                # the reference passes `g`, not the locally re-sorted data (:167); identical unless the
                # caller shuffled time stamps after construction, in which case `g.data` is what counts
                src_ei = edge_index if data is g.data else _plain(_staging.up(g.data.edge_index, dev)).long()
                src_t = time if data is g.data else _staging.up(g.data.time, dev)
                edge_index = ops.lift_order_temporal(src_ei, src_t, delta, n)
            else:
                edge_index = _plain(_staging.up(event_graph, dev)).long()
            if edge_weight is not None:
                edge_weight = ops.pair_attributes(edge_index, edge_weight, "src")
            if cached or max_order == 2:
                m.layers[2] = _aggregate(edge_index, node_sequence, edge_weight)
            for k in range(3, max_order + 1):
                save = cached or k == max_order
                edge_index, node_sequence, edge_weight, gk = _lift_step(edge_index, node_sequence, edge_weight, "src", save)
                if save:
                    m.layers[k] = gk

        for k, layer in m.layers.items():
            if to_host:
                layer.to("cpu")
            if k > 1:
                layer.mapping = HigherOrderIndexMap(g.mapping, layer.data.node_sequence)
        return m

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def from_path_data(path_data: PathData, max_order: int = 1, mode: str = "propagation",
                       cached: bool = True) -> "MultiOrderModel":
        """multi_order_model.py:194-241."""
        m = MultiOrderModel()
        pg = path_data.data
        dev, to_host = _staging.compute_device(pg.edge_index, pg.node_sequence)
        edge_index = _plain(_staging.up(pg.edge_index, dev)).long()
        node_sequence = _plain(_staging.up(pg.node_sequence, dev)).long()
        edge_weight = _staging.up(pg.dag_weight, dev).repeat_interleave(_staging.up(pg.dag_num_edges, dev))
        if mode == "diffusion":
            outdeg = torch.bincount(edge_index[0], minlength=node_sequence.size(0))
            edge_weight = edge_weight / outdeg[edge_index[0]]
            aggr = "mul"
        elif mode == "propagation":
            aggr = "src"
        else:
            raise ValueError(f"Unknown mode {mode}")  # the reference dies with a NameError here (:218-224)

        m.layers[1] = _aggregate(edge_index, node_sequence, edge_weight)
        m.layers[1].mapping = path_data.mapping
        for k in range(2, max_order + 1):
            save = cached or k == max_order
            edge_index, node_sequence, edge_weight, gk = _lift_step(edge_index, node_sequence, edge_weight, aggr, save)
            if save:
                m.layers[k] = gk
        for k, layer in m.layers.items():
            if to_host:
                layer.to("cpu")
            if k > 1:
                layer.mapping = HigherOrderIndexMap(path_data.mapping, layer.data.node_sequence)
        return m

    # ------------------------------------------------------------------------------------------
    def to_dbgnn_data(self, max_order: int = 2, mapping: str = "last", x_h: torch.Tensor | None = None) -> Data:
        """multi_order_model.py:511-554.  ``x`` is ``layers[1].data.x`` or a dense one-hot matrix;
        ``x_h`` is the reference's dense ``eye(n_ho)`` unless given explicitly (an extension: the
        one-hot matrix needs n_ho^2 floats and cannot be built for large layers)."""
        if max_order not in self.layers:
            logger.error("Higher-order graph of specified order not found.")
            raise ValueError(f"Higher-order graph of order {max_order} not found.")
        g, gk = self.layers[1], self.layers[max_order]
        dev = g.data.edge_index.device
        x = g.data.x if g.data.x is not None else torch.eye(g.data.num_nodes, g.data.num_nodes, device=dev)
        if x_h is None:
            x_h = torch.eye(gk.data.num_nodes, gk.data.num_nodes, device=gk.data.edge_index.device)
        return Data(
            num_nodes=g.data.num_nodes,
            num_ho_nodes=gk.data.num_nodes,
            x=x,
            x_h=x_h,
            edge_index=g.data.edge_index,
            edge_index_higher_order=gk.data.edge_index,
            edge_weights=g.data.edge_weight.float(),
            edge_weights_higher_order=gk.data.edge_weight.float(),
            bipartite_edge_index=generate_bipartite_edge_index(g, gk, mapping=mapping, device=dev),
            y=g.data.y,
        )
