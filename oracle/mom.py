"""CPU restatement of the MultiOrderModel builders on raw tensors.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference (relative to ``/root/reference``):
* ``src/pathpyG/core/multi_order_model.py:83-122``  iterate_lift_order
* ``src/pathpyG/core/multi_order_model.py:124-192`` from_temporal_graph
* ``src/pathpyG/core/multi_order_model.py:194-241`` from_path_data
* ``src/pathpyG/core/multi_order_model.py:511-554`` to_dbgnn_data
* ``src/pathpyG/core/path_data.py:126-159``         PathData.append_walks
* ``src/pathpyG/utils/dbgnn.py:10-46``              generate_bipartite_edge_index

Label bookkeeping (``IndexMap``) is not restated: it carries no arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import lift, pyg


def iterate_lift_order(edge_index, node_sequence, edge_weight=None, aggr="src", save=True):
    """multi_order_model.py:108-122 without the IndexMap line (:119)."""
    if edge_weight is None:
        ho_index = lift.lift_order_edge_index(edge_index, num_nodes=node_sequence.size(0))       # :109
    else:
        ho_index, edge_weight = lift.lift_order_edge_index_weighted(
            edge_index, edge_weight=edge_weight, num_nodes=node_sequence.size(0), aggr=aggr)      # :111-113
    node_sequence = torch.cat([node_sequence[edge_index[0]], node_sequence[edge_index[1]][:, -1:]], dim=1)  # :114
    gk = lift.aggregate_edge_index(ho_index, node_sequence, edge_weight) if save else None         # :117-121
    return ho_index, node_sequence, edge_weight, gk


def from_temporal_graph(edge_index, time, num_nodes, delta=1, max_order=1, edge_weight=None,
                        cached=True, event_graph=None, temporal_fn=None) -> dict:
    """multi_order_model.py:124-192.  ``edge_index``/``time`` are already time-sorted
    (TemporalGraph.__init__, temporal_graph.py:58-63).  ``temporal_fn`` lets a test swap
    the O(T*m) loop for the closed form at large sizes; default is the reference's loop."""
    layers: dict[int, lift.Layer] = {}
    node_sequence = torch.arange(num_nodes).unsqueeze(1)                                            # :153
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1))                                                # :157
    if cached or max_order == 1:
        layers[1] = lift.aggregate_edge_index(edge_index, node_sequence, edge_weight)               # :159-161
    if max_order > 1:
        node_sequence = torch.cat([node_sequence[edge_index[0]], node_sequence[edge_index[1]][:, -1:]], dim=1)  # :165
        if event_graph is None:
            fn = temporal_fn or lift.lift_order_temporal
            edge_index = fn(edge_index, time, delta)                                                # :167
        else:
            edge_index = event_graph                                                                # :169
        edge_weight = lift.aggregate_node_attributes(edge_index, edge_weight, "src")                # :170
        if cached or max_order == 2:
            layers[2] = lift.aggregate_edge_index(edge_index, node_sequence, edge_weight)           # :174-176
        for k in range(3, max_order + 1):                                                           # :181-191
            edge_index, node_sequence, edge_weight, gk = iterate_lift_order(
                edge_index, node_sequence, edge_weight, aggr="src", save=cached or k == max_order)
            if cached or k == max_order:
                layers[k] = gk
    return layers


@dataclass
class Walks:
    """PathData.data fields (path_data.py:57-64)."""

    edge_index: torch.Tensor     # [2, sum(L) - P] over global positions
    node_sequence: torch.Tensor  # [sum(L), 1]
    dag_weight: torch.Tensor     # [P] float32
    dag_num_edges: torch.Tensor  # [P]
    dag_num_nodes: torch.Tensor  # [P]


def append_walks(node_seqs, weights) -> Walks:
    """path_data.py:126-159 for index sequences (IndexMap lookups left out)."""
    lengths = torch.tensor([len(s) for s in node_seqs])
    flat = torch.tensor([v for s in node_seqs for v in s], dtype=torch.long).unsqueeze(1)
    pos = torch.arange(int(lengths.sum()))
    chain = torch.stack([pos[:-1], pos[1:]])                                                        # :144
    keep = torch.ones(chain.size(1), dtype=torch.bool)
    bounds = pyg.cumsum(lengths)
    keep[bounds[1:-1] - 1] = False                                                                  # :147-150
    return Walks(chain[:, keep], flat, torch.tensor(weights, dtype=torch.float), lengths - 1, lengths)


def from_path_data(walks: Walks, max_order=1, mode="propagation", cached=True) -> dict:
    """multi_order_model.py:194-241."""
    layers: dict[int, lift.Layer] = {}
    edge_index, node_sequence = walks.edge_index, walks.node_sequence
    edge_weight = walks.dag_weight.repeat_interleave(walks.dag_num_edges)                           # :217
    if mode == "diffusion":
        outdeg = pyg.degree(edge_index[0], node_sequence.size(0), dtype=torch.long)
        edge_weight = edge_weight / outdeg[edge_index[0]]                                           # :219-221
        aggr = "mul"
    elif mode == "propagation":
        aggr = "src"
    layers[1] = lift.aggregate_edge_index(edge_index, node_sequence, edge_weight)                   # :226
    for k in range(2, max_order + 1):                                                               # :229-239
        edge_index, node_sequence, edge_weight, gk = iterate_lift_order(
            edge_index, node_sequence, edge_weight, aggr=aggr, save=cached or k == max_order)
        if cached or k == max_order:
            layers[k] = gk
    return layers


def generate_bipartite_edge_index(ho_node_sequence: torch.Tensor, mapping: str = "last") -> torch.Tensor:
    """utils/dbgnn.py:33-44 -- note that "last" reads column 1, not column -1 (:34)."""
    n = ho_node_sequence.size(0)
    ids = torch.arange(n)
    if mapping == "last":
        return torch.stack([ids, ho_node_sequence[:, 1]])
    if mapping == "first":
        return torch.stack([ids, ho_node_sequence[:, 0]])
    return torch.stack([torch.cat([ids, ids]), torch.cat([ho_node_sequence[:, 0], ho_node_sequence[:, 1]])])


def to_dbgnn_data(layers: dict, max_order=2, mapping="last", x=None, x_h=None) -> dict:
    """multi_order_model.py:511-554.  ``x`` / ``x_h`` default to the reference's dense one-hot
    matrices (:529-533); large configurations pass explicit feature matrices instead."""
    if max_order not in layers:
        raise ValueError(f"Higher-order graph of order {max_order} not found.")                     # :521-523
    g, gk = layers[1], layers[max_order]
    return dict(
        num_nodes=g.num_nodes,
        num_ho_nodes=gk.num_nodes,
        x=torch.eye(g.num_nodes) if x is None else x,
        x_h=torch.eye(gk.num_nodes) if x_h is None else x_h,
        edge_index=g.edge_index,
        edge_index_higher_order=gk.edge_index,
        edge_weights=g.edge_weight.float(),
        edge_weights_higher_order=gk.edge_weight.float(),
        bipartite_edge_index=generate_bipartite_edge_index(gk.node_sequence, mapping),
    )
