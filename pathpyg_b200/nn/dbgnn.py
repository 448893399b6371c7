"""Drop-in for ``pathpyG.nn.dbgnn`` (reference ``src/pathpyG/nn/dbgnn.py``): ``DBGNN`` and
``BipartiteGraphOperator`` as ``nn.Module``s with the reference's parameter names
(``first_order_layers.{i}.lin.weight`` / ``.bias``, ``higher_order_layers.{i}.…``,
``bipartite_layer.lin{1,2}.{weight,bias}``, ``lin.{weight,bias}``), so state dicts move between
the two implementations unchanged.

Message passing does not go through torch_geometric / torch_scatter: each graph is regrouped by
target node once per forward (``ops.gcn_prepare`` / ``ops.csc_build``) and every layer is a
segment-reduce SpMM + dense transform in the sm_100a kernels of ``csrc/dbgnn.cu``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib, _staging, ops


class GCNConv(nn.Module):
    """PyG ``GCNConv(in, out)`` with its defaults (normalize, add_self_loops, bias, not cached):
    ``out = D^-1/2 (A + I) D^-1/2 X W^T + b`` aggregated at the edge target."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))  # glorot, like PyG's Linear in GCNConv
        nn.init.uniform_(self.lin.weight, -a, a)
        nn.init.zeros_(self.bias)

    def forward_prepared(self, x: torch.Tensor, graph: ops.TargetGroupedEdges, act: int = _lib.ACT_NONE) -> torch.Tensor:
        w, b = self.lin.weight, self.bias
        if ops.fused_supported(self.in_channels, self.out_channels):
            return ops.gcn_layer_fused(graph, x, w, b, act)  # aggregate + transform + bias + act in one kernel
        if self.in_channels <= self.out_channels:      # aggregate the narrower side first
            return ops.linear(ops.spmm_csc(graph, x), w, b, act)
        return ops.spmm_csc(graph, ops.linear(x, w), b, act)

    def forward(self, x, edge_index, edge_weight=None):
        graph = ops.gcn_prepare(edge_index, edge_weight, x.size(0))
        return self.forward_prepared(x, graph)


class BipartiteGraphOperator(nn.Module):
    """nn/dbgnn.py:32-69: ``out[v] = sum_{u -> v} (lin1(x_h)[u] + lin2(x)[v])`` from higher-order
    nodes u to first-order nodes v.  Evaluated as ``W1 (sum_u x_h[u]) + indeg(v) (W2 x[v] + b1 + b2)``."""

    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.lin1 = nn.Linear(in_ch, out_ch)
        self.lin2 = nn.Linear(in_ch, out_ch)

    def forward(self, x, bipartite_index, n_ho: int, n_fo: int, act: int = _lib.ACT_NONE):
        x_h, x_fo = x
        grouped = ops.csc_build(bipartite_index, n_ho, n_fo)
        if ops.fused_supported(self.lin1.in_features, self.lin1.out_features):
            return ops.bipartite_fused(grouped, x_h, x_fo, self.lin1.weight, self.lin2.weight,
                                       self.lin1.bias + self.lin2.bias, act)
        summed = ops.spmm_csc(grouped, x_h)
        indeg = ops.colptr_counts(grouped)
        return ops.linear(summed, self.lin1.weight, self.lin1.bias + self.lin2.bias, act,
                          a2=x_fo, w2=self.lin2.weight, rowscale=indeg)


class DBGNN(nn.Module):
    """nn/dbgnn.py:72-151."""

    def __init__(self, num_classes: int, num_features, hidden_dims: list[int], p_dropout: float = 0.0):
        super().__init__()
        self.num_features = num_features
        self.num_classes = num_classes
        self.hidden_dims = hidden_dims
        self.p_dropout = p_dropout

        self.higher_order_layers = nn.ModuleList([GCNConv(num_features[1], hidden_dims[0])])
        self.first_order_layers = nn.ModuleList([GCNConv(num_features[0], hidden_dims[0])])
        for dim in range(1, len(hidden_dims) - 1):
            self.higher_order_layers.append(GCNConv(hidden_dims[dim - 1], hidden_dims[dim]))
            self.first_order_layers.append(GCNConv(hidden_dims[dim - 1], hidden_dims[dim]))
        self.bipartite_layer = BipartiteGraphOperator(hidden_dims[-2], hidden_dims[-1])
        self.lin = nn.Linear(hidden_dims[-1], num_classes)

    def forward(self, data) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("pathpyg_b200.nn.DBGNN: backward pass not built yet -- call under torch.no_grad()")
        dev, to_host = _staging.compute_device(data.x, data.x_h, data.edge_index)
        x, x_h = _staging.up(data.x, dev).float(), _staging.up(data.x_h, dev).float()
        ei, ei_h = _staging.up(data.edge_index, dev), _staging.up(data.edge_index_higher_order, dev)
        w, w_h = _staging.up(data.edge_weights, dev), _staging.up(data.edge_weights_higher_order, dev)
        bip = _staging.up(data.bipartite_edge_index, dev)
        n_fo, n_ho = int(data.num_nodes), int(data.num_ho_nodes)
        drop = self.training and self.p_dropout > 0.0

        fo_graph = ops.gcn_prepare(ei, w, n_fo)
        for layer in self.first_order_layers:
            if drop:
                x = F.dropout(x, p=self.p_dropout, training=True)
            x = layer.forward_prepared(x, fo_graph, _lib.ACT_ELU)
        ho_graph = ops.gcn_prepare(ei_h, w_h, n_ho)
        for layer in self.higher_order_layers:
            if drop:
                x_h = F.dropout(x_h, p=self.p_dropout, training=True)
            x_h = layer.forward_prepared(x_h, ho_graph, _lib.ACT_ELU)
        if drop:
            x = F.dropout(x, p=self.p_dropout, training=True)
            x_h = F.dropout(x_h, p=self.p_dropout, training=True)
        x = self.bipartite_layer((x_h, x), bip, n_ho, n_fo, _lib.ACT_ELU)
        if drop:
            x = F.dropout(x, p=self.p_dropout, training=True)
        out = ops.linear(x, self.lin.weight, self.lin.bias)
        return _staging.down(out, to_host)
