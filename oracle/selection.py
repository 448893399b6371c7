"""CPU restatement of the model-selection statistics that consume the built layers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference (relative to ``/root/reference``):
* ``src/pathpyG/core/graph.py:486-533``              Graph.degrees / transition_probabilities
* ``src/pathpyG/core/multi_order_model.py:243-312``  get_mon_dof
* ``src/pathpyG/core/multi_order_model.py:314-409``  zeroth / intermediate / multi-order log-likelihood
* ``src/pathpyG/core/multi_order_model.py:411-459``  likelihood_ratio_test
* ``src/pathpyG/core/multi_order_model.py:461-509``  estimate_order (without the IndexMap set check, :492-496)

``layers`` is the dict of ``oracle.lift.Layer`` that ``oracle.mom.from_path_data`` / ``from_temporal_graph``
return; ``walks`` is ``oracle.mom.Walks`` (the fields of ``PathData.data``).  Third-party arithmetic restated
from torch_geometric 2.7.0: ``EdgeIndex.matmul`` (sparse adjacency product, unit values) is done with a scipy CSR
product here.

Pinning: every known-answer vector of ``tests/core/test_multi_order_model.py:45-162,193-224`` is checked in
``tests/test_oracle.py``; ``oracle/ref_loader.selection_methods`` additionally executes the reference's own method
bodies on the same layers (``tests/test_oracle_vs_reference.py`` and the committed ``tests/golden/selection_golden.npz``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
from scipy.stats import chi2

from . import lift, pyg


def degrees(layer, mode: str = "in", edge_attr: bool = False) -> torch.Tensor:
    """graph.py:486-516 with ``return_tensor=True``."""
    ids = layer.edge_index[1] if mode == "in" else layer.edge_index[0]
    if not edge_attr:
        return pyg.degree(ids, layer.num_nodes, dtype=torch.int)                                   # :500,506
    return pyg.scatter(layer.edge_weight, ids, dim_size=layer.num_nodes, reduce="sum")             # :503,509


def transition_probabilities(layer, edge_attr: bool = False) -> torch.Tensor:
    """graph.py:518-533."""
    out_degree = degrees(layer, "out", edge_attr)                                                  # :528
    weight = layer.edge_weight if edge_attr else torch.ones(layer.edge_index.size(1))              # :530-532
    return weight / out_degree[layer.edge_index[0]]                                                # :533


def get_mon_dof(layers: dict, max_order=None, assumption: str = "paths") -> int:
    """multi_order_model.py:270-312."""
    if max_order is None:
        max_order = max(layers)
    if max_order > max(layers):
        raise ValueError("max_order cannot be larger than maximum order of multi-order network")
    n = layers[1].num_nodes
    dof = n - 1                                                                                    # :277
    if assumption == "paths":
        edge_index = layers[1].edge_index
        for k in range(1, max_order + 1):                                                          # :285-291
            if k > 1:
                num_nodes = 0 if edge_index.numel() == 0 else int(edge_index.max()) + 1
                edge_index = lift.lift_order_edge_index(edge_index, num_nodes)
            dof += edge_index.shape[1]
        e1 = layers[1].edge_index.numpy()
        adj = sp.csr_matrix((np.ones(e1.shape[1]), (e1[0], e1[1])), shape=(n, n))                  # :298
        power = None
        for k in range(1, max_order + 1):                                                          # :294-303
            power = adj if k == 1 else power @ adj
            dof -= int(np.unique(power.tocoo().row).shape[0])
    elif assumption == "ngrams":
        for order in range(1, max_order + 1):
            dof += (n ** order) * (n - 1)                                                          # :306-307
    else:
        raise ValueError(f"Unknown assumption {assumption}. Only 'path' and 'ngram' are accepted.")
    return int(dof)


def get_zeroth_order_log_likelihood(walks) -> float:
    """multi_order_model.py:314-339."""
    frequencies = walks.dag_weight
    mask = torch.ones(walks.node_sequence.size(0), dtype=torch.bool)
    mask[walks.edge_index[1]] = False                                                              # :329-330
    start_ixs = walks.node_sequence.squeeze()[mask]                                                # :331
    _, counts = torch.unique(walks.node_sequence, return_counts=True)                              # :335
    probs = counts / counts.sum()                                                                  # :338
    return torch.mul(frequencies, torch.log(probs[start_ixs])).sum().item()                        # :339


def get_intermediate_order_log_likelihood(layers: dict, walks, order: int) -> float:
    """multi_order_model.py:341-369."""
    frequencies = walks.dag_weight
    shrunk = walks.dag_num_nodes - order                                                           # :356
    lengths = shrunk[shrunk > 0]
    frequencies = frequencies[shrunk > 0]
    starts = pyg.cumsum(lengths)[:-1]                                                              # :361
    probs = transition_probabilities(layers[order])[layers[order + 1].inverse_idx[starts]]         # :363-365
    return torch.mul(frequencies, torch.log(probs)).sum().item()                                   # :367-369


def get_mon_log_likelihood(layers: dict, walks, max_order: int = 1) -> float:
    """multi_order_model.py:371-409."""
    llh = 0.0
    llh += get_zeroth_order_log_likelihood(walks)                                                  # :386
    for order in range(1, max_order):
        llh += get_intermediate_order_log_likelihood(layers, walks, order)                         # :389-390
    if max_order > 0:
        probs = transition_probabilities(layers[max_order], edge_attr=True)                        # :394
        llh += (torch.log(probs) * layers[max_order].edge_weight).sum().item()                     # :395-397
    else:
        frequencies = walks.dag_weight
        counts = torch.bincount(walks.node_sequence.squeeze(), frequencies.repeat_interleave(walks.dag_num_nodes))
        probs = counts / counts.sum()
        llh = torch.mul(torch.log(probs), counts).sum().item()                                     # :402-407
    return llh


def likelihood_ratio_test(layers: dict, walks, max_order_null: int = 0, max_order: int = 1, assumption: str = "paths",
                          significance_threshold: float = 0.01) -> tuple:
    """multi_order_model.py:436-459."""
    if max_order_null >= max_order:
        raise ValueError("order of null hypothesis must be smaller than order of alternative hypothesis")
    if max_order > max(layers):
        raise ValueError("order of hypotheses must be smaller than max. order of MultiOrderModel")
    x = -2 * (get_mon_log_likelihood(layers, walks, max_order_null) - get_mon_log_likelihood(layers, walks, max_order))
    dof_diff = get_mon_dof(layers, max_order, assumption) - get_mon_dof(layers, max_order_null, assumption)
    p = 1 - chi2.cdf(x, dof_diff)
    return (p < significance_threshold), p


def estimate_order(layers: dict, walks, max_order=None, significance_threshold: float = 0.01) -> int:
    """multi_order_model.py:484-509."""
    if max_order is None:
        max_order = max(layers)
    if max_order > max(layers):
        raise ValueError("max_order cannot be larger than maximum order of multi-order network")
    if max_order <= 1:
        raise ValueError("max_order must be larger than one")
    accepted = 1
    for k in range(2, max_order + 1):
        if likelihood_ratio_test(layers, walks, k - 1, k, significance_threshold=significance_threshold)[0]:
            accepted = k
    return accepted
