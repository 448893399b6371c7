"""CPU restatement of DBGNN.forward on plain torch (PARITY UNPINNED, see oracle/__init__.py).

TEST INFRASTRUCTURE ONLY.

Reference (relative to ``/root/reference``): ``src/pathpyG/nn/dbgnn.py:32-69``
(BipartiteGraphOperator), ``:86-151`` (DBGNN).  ``GCNConv`` and
``MessagePassing.propagate`` come from torch_geometric 2.7.0 (absent; restated from
its published algorithm): ``gcn_norm`` -> ``x @ W^T`` -> scatter-add at the target
(``edge_index[1]``) -> ``+ bias``.

Parameters are addressed by the reference module's ``state_dict`` keys so that
weights move between the two implementations unchanged:
``first_order_layers.{i}.lin.weight`` / ``.bias``, ``higher_order_layers.{i}.…``,
``bipartite_layer.lin{1,2}.{weight,bias}``, ``lin.{weight,bias}``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import pyg


def gcn_conv(x, edge_index, edge_weight, weight, bias):
    """PyG GCNConv.forward with defaults (normalize, add_self_loops, bias; not cached)."""
    n = x.size(0)
    ei, norm = pyg.gcn_norm(edge_index, edge_weight, n, dtype=x.dtype)
    h = x @ weight.t()
    out = torch.zeros(n, weight.size(0), dtype=h.dtype).index_add_(0, ei[1], norm.unsqueeze(1) * h[ei[0]])
    return out + bias


def bipartite_operator(x_h, x, bipartite_index, n_fo, w1, b1, w2, b2):
    """nn/dbgnn.py:47-69: message x_i + x_j with j = HO source, i = FO target, summed at the target."""
    h_ho = F.linear(x_h, w1, b1)
    h_fo = F.linear(x, w2, b2)
    msg = h_fo[bipartite_index[1]] + h_ho[bipartite_index[0]]
    return torch.zeros(n_fo, w1.size(0), dtype=msg.dtype).index_add_(0, bipartite_index[1], msg)


def num_gcn_layers(params: dict, prefix: str) -> int:
    i = 0
    while f"{prefix}.{i}.lin.weight" in params:
        i += 1
    return i


def dbgnn_forward(params: dict, data: dict) -> torch.Tensor:
    """nn/dbgnn.py:121-151 in eval mode (dropout is the identity)."""
    x, x_h = data["x"], data["x_h"]
    for i in range(num_gcn_layers(params, "first_order_layers")):                                   # :131-133
        p = f"first_order_layers.{i}"
        x = F.elu(gcn_conv(x, data["edge_index"], data["edge_weights"], params[p + ".lin.weight"], params[p + ".bias"]))
    for i in range(num_gcn_layers(params, "higher_order_layers")):                                  # :137-139
        p = f"higher_order_layers.{i}"
        x_h = F.elu(gcn_conv(x_h, data["edge_index_higher_order"], data["edge_weights_higher_order"],
                             params[p + ".lin.weight"], params[p + ".bias"]))
    x = F.elu(bipartite_operator(x_h, x, data["bipartite_edge_index"], data["num_nodes"],           # :143-145
                                 params["bipartite_layer.lin1.weight"], params["bipartite_layer.lin1.bias"],
                                 params["bipartite_layer.lin2.weight"], params["bipartite_layer.lin2.bias"]))
    return F.linear(x, params["lin.weight"], params["lin.bias"])                                    # :149


def init_params(num_classes: int, num_features, hidden_dims, seed: int = 0, dtype=torch.float32) -> dict:
    """Random parameters with the reference's shapes (nn/dbgnn.py:102-119): glorot for GCN
    weights / zero GCN bias as PyG does, uniform(-1/sqrt(fan_in), ..) for the Linear layers."""
    g = torch.Generator().manual_seed(seed)
    params = {}

    def glorot(out_f, in_f):
        a = (6.0 / (in_f + out_f)) ** 0.5
        return ((torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)

    def linear(out_f, in_f):
        a = 1.0 / in_f ** 0.5
        w = ((torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)
        b = ((torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)
        return w, b

    dims_fo = [num_features[0]] + list(hidden_dims[:-1])
    dims_ho = [num_features[1]] + list(hidden_dims[:-1])
    for i in range(len(hidden_dims) - 1):
        params[f"first_order_layers.{i}.lin.weight"] = glorot(dims_fo[i + 1], dims_fo[i])
        params[f"first_order_layers.{i}.bias"] = torch.zeros(dims_fo[i + 1], dtype=dtype)
        params[f"higher_order_layers.{i}.lin.weight"] = glorot(dims_ho[i + 1], dims_ho[i])
        params[f"higher_order_layers.{i}.bias"] = torch.zeros(dims_ho[i + 1], dtype=dtype)
    w, b = linear(hidden_dims[-1], hidden_dims[-2])
    params["bipartite_layer.lin1.weight"], params["bipartite_layer.lin1.bias"] = w, b
    w, b = linear(hidden_dims[-1], hidden_dims[-2])
    params["bipartite_layer.lin2.weight"], params["bipartite_layer.lin2.bias"] = w, b
    w, b = linear(num_classes, hidden_dims[-1])
    params["lin.weight"], params["lin.bias"] = w, b
    return params
