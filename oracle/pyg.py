"""torch_geometric 2.7.0 utilities used by the hot path, restated on plain torch.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

torch_geometric is a PyPI dependency of the reference (``pyproject.toml:34``,
pinned ``torch-geometric 2.7.0`` in ``uv.lock:5910-5912``); it is neither vendored
in ``/root/reference`` nor installed in this image, so the published algorithm of
each utility is restated here.  Call sites in the reference:

* ``degree``  -- ``algorithms/lift_order.py:65``, ``core/multi_order_model.py:220``
* ``cumsum``  -- ``algorithms/lift_order.py:74,77``, ``core/path_data.py:149``
* ``coalesce``-- ``algorithms/lift_order.py:139-144``
* ``GCNConv`` -- ``nn/dbgnn.py:104-114`` (``gcn_norm`` + ``add_remaining_self_loops``)
"""
from __future__ import annotations

import torch


def degree(index: torch.Tensor, num_nodes: int, dtype=torch.long) -> torch.Tensor:
    """PyG ``utils.degree``: ``zeros(N).scatter_add_(0, index, ones)``."""
    out = torch.zeros(num_nodes, dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=dtype, device=index.device))


def cumsum(x: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """PyG ``utils.cumsum``: inclusive cumsum with a leading zero (length + 1)."""
    assert dim == 0 and x.dim() == 1
    out = x.new_zeros(x.numel() + 1)
    torch.cumsum(x, 0, out=out[1:])
    return out


def scatter(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str = "sum") -> torch.Tensor:
    """PyG ``utils.scatter`` along dim 0 for the reductions ``coalesce`` forwards."""
    if reduce in ("sum", "add"):
        return src.new_zeros((dim_size,) + src.shape[1:]).index_add_(0, index, src)
    if reduce == "mean":
        count = src.new_zeros(dim_size).index_add_(0, index, src.new_ones(src.size(0)))
        out = src.new_zeros((dim_size,) + src.shape[1:]).index_add_(0, index, src)
        count = count.clamp(min=1)
        if out.is_floating_point():
            return out / count.view((-1,) + (1,) * (src.dim() - 1))
        return out.div(count.view((-1,) + (1,) * (src.dim() - 1)), rounding_mode="floor")
    if reduce in ("min", "max"):
        idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
        out = src.new_zeros((dim_size,) + src.shape[1:])
        return out.scatter_reduce_(0, idx, src, reduce="a" + reduce, include_self=False)
    raise ValueError(f"unknown reduce {reduce}")


def coalesce(edge_index: torch.Tensor, edge_attr: torch.Tensor | None, num_nodes: int, reduce: str = "sum"):
    """PyG ``utils.coalesce`` (``sort_by_row=True``, ``is_sorted=False``).

    key = row * num_nodes + col; sort keys carrying the permutation; keep the first
    column of every run of equal keys; reduce the (permuted) attributes per run.
    PyG sorts with ``Tensor.sort(stable=False)`` when pyg_lib is absent; a stable
    sort is used here so that the fp32 summation order is defined (identical
    results whenever the per-run sums are exact, e.g. integer-valued weights).
    """
    num_edges = edge_index.size(1)
    key = edge_index.new_empty(num_edges + 1)
    key[0] = -1
    key[1:] = edge_index[0] * num_nodes + edge_index[1]
    sorted_key, perm = key[1:].sort(stable=True)
    key[1:] = sorted_key
    edge_index = edge_index[:, perm]
    if edge_attr is not None:
        edge_attr = edge_attr[perm]
    mask = key[1:] > key[:-1]
    if bool(mask.all()):
        return edge_index, edge_attr
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    run_id = torch.arange(num_edges, device=key.device) - (~mask).cumsum(0)
    return edge_index, scatter(edge_attr, run_id, edge_index.size(1), reduce)


def add_remaining_self_loops(edge_index: torch.Tensor, edge_weight: torch.Tensor, fill_value: float, num_nodes: int):
    """PyG ``utils.add_remaining_self_loops``.

    Non-loop edges keep their order; then one loop per node is appended whose
    weight is the node's existing self-loop weight if it had one (the last one in
    edge order wins, as ``loop_attr[idx] = attr`` does), else ``fill_value``.
    """
    mask = edge_index[0] != edge_index[1]
    loop_index = torch.arange(num_nodes, device=edge_index.device)
    loop_attr = edge_weight.new_full((num_nodes,), fill_value)
    inv = ~mask
    # same statement as PyG; on CPU index_put_ runs in order, so the LAST duplicate
    # self-loop of a node wins (coalesced De Bruijn layers have at most one anyway)
    loop_attr[edge_index[0][inv]] = edge_weight[inv]
    ei = torch.cat([edge_index[:, mask], loop_index.unsqueeze(0).repeat(2, 1)], dim=1)
    ew = torch.cat([edge_weight[mask], loop_attr], dim=0)
    return ei, ew


def gcn_norm(edge_index: torch.Tensor, edge_weight: torch.Tensor | None, num_nodes: int, dtype=torch.float32):
    """PyG ``nn.conv.gcn_conv.gcn_norm`` (``improved=False``, ``add_self_loops=True``,
    ``flow='source_to_target'``): D^-1/2 (A + I) D^-1/2 with in-degree at ``col``."""
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, 1.0, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = edge_weight.new_zeros(num_nodes).index_add_(0, col, edge_weight)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


def to_undirected(edge_index: torch.Tensor, edge_attr: torch.Tensor | None, num_nodes: int, reduce: str = "add"):
    """PyG ``utils.to_undirected`` (call site ``core/graph.py:228-233``): every edge in both directions, the
    attribute repeated for the reversed copy, then ``coalesce`` with ``reduce``."""
    row, col = edge_index[0], edge_index[1]
    both = torch.stack([torch.cat([row, col]), torch.cat([col, row])])
    if edge_attr is not None:
        edge_attr = torch.cat([edge_attr, edge_attr])
    return coalesce(both, edge_attr, num_nodes, reduce)
