"""CPU restatement of the reference's lift / aggregation functions.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference (relative to ``/root/reference``):
* ``src/pathpyG/algorithms/lift_order.py:10-45``   aggregate_node_attributes
* ``src/pathpyG/algorithms/lift_order.py:48-79``   lift_order_edge_index
* ``src/pathpyG/algorithms/lift_order.py:82-106``  lift_order_edge_index_weighted
* ``src/pathpyG/algorithms/lift_order.py:109-152`` aggregate_edge_index
  (+ the row sort of ``Graph.__init__``, ``src/pathpyG/core/graph.py:103-105``)
* ``src/pathpyG/algorithms/temporal.py:17-54``     lift_order_temporal

Each function exists once in the reference's own operation order on torch CPU
tensors (the arm that ``bench.py`` times as the CPU baseline, kind "port") and,
where the reference's form is too slow for parity tests at bench sizes, once
more as a closed-form numpy statement (``*_closed_form``) that is checked against
the first on small inputs in ``tests/test_oracle.py``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import pyg

_PAIR_RULES = {
    "src": lambda a, b: a,
    "dst": lambda a, b: b,
    "max": torch.maximum,
    "mul": torch.mul,
    "add": torch.add,
}


def aggregate_node_attributes(edge_index: torch.Tensor, node_attribute: torch.Tensor, aggr: str = "src") -> torch.Tensor:
    """lift_order.py:10-45 -- one attribute per edge from its two end points."""
    if aggr not in _PAIR_RULES:
        raise ValueError(f"Unknown aggregation method {aggr}")  # lift_order.py:44
    at_src = node_attribute[edge_index[0]]
    if aggr == "src":
        return at_src
    at_dst = node_attribute[edge_index[1]]
    return _PAIR_RULES[aggr](at_src, at_dst)


def lift_order_edge_index(edge_index: torch.Tensor, num_nodes: int | None = None) -> torch.Tensor:
    """lift_order.py:48-79 -- line graph of a row-sorted edge_index.

    Column e=(u,v) yields outdeg(v) columns (e, ptr[v]+j), j = 0..outdeg(v)-1, in
    column order; ptr = exclusive prefix sum of the out-degrees.
    """
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1  # lift_order.py:62-63
    row, col = edge_index[0], edge_index[1]
    outdeg = pyg.degree(row, num_nodes, dtype=torch.long)            # :65
    fan = outdeg[col]                                                # :68
    ho_src = torch.repeat_interleave(fan)                            # :70
    first_out = pyg.cumsum(outdeg)[:-1]                              # :74
    ho_dst = torch.repeat_interleave(first_out[col], fan)            # :75
    within = torch.arange(ho_src.numel(), dtype=torch.long) - pyg.cumsum(fan)[ho_src]  # :76-77
    return torch.stack([ho_src, ho_dst + within], dim=0)             # :78-79


def lift_order_edge_index_weighted(edge_index, edge_weight, num_nodes=None, aggr="src"):
    """lift_order.py:82-106."""
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    ho_index = lift_order_edge_index(edge_index, num_nodes)
    return ho_index, aggregate_node_attributes(ho_index, edge_weight, aggr)


@dataclass
class Layer:
    """The fields ``aggregate_edge_index`` stores in ``Graph.data`` (lift_order.py:145-151)."""

    edge_index: torch.Tensor      # [2, E^] (row, col)-sorted
    num_nodes: int
    node_sequence: torch.Tensor   # [n, k] sorted distinct rows
    edge_weight: torch.Tensor     # [E^]
    inverse_idx: torch.Tensor     # [rows of the input node_sequence]


def aggregate_edge_index(edge_index, node_sequence, edge_weight=None, aggr: str = "sum") -> Layer:
    """lift_order.py:109-152 followed by the row sort of Graph.__init__ (graph.py:103-105)."""
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1))                                    # :130-131
    unique_nodes, inverse_idx = torch.unique(node_sequence, dim=0, return_inverse=True)  # :133
    if node_sequence.size(1) == 1:
        mapped = node_sequence.squeeze()[edge_index]                                    # :135-136
    else:
        mapped = inverse_idx[edge_index]                                                # :138
    agg_index, agg_weight = pyg.coalesce(mapped, edge_weight, unique_nodes.size(0), aggr)  # :139-144
    # Graph.__init__ -> EdgeIndex.sort_by("row") is stable; after coalesce it is the identity,
    # but restate it so that the oracle does not depend on that observation.
    order = torch.sort(agg_index[0], stable=True).indices
    return Layer(agg_index[:, order], int(unique_nodes.size(0)), unique_nodes, agg_weight[order], inverse_idx)


def lift_order_temporal(edge_index: torch.Tensor, timestamps: torch.Tensor, delta=1, *, max_source_stamps: int | None = None,
                        budget_s: float | None = None, stats: dict | None = None) -> torch.Tensor:
    """temporal.py:17-54 in the reference's own order of operations.

    (e -> f) iff dst(e) == src(f) and t_e < t_f <= t_e + delta.  The loop over the
    distinct timestamps and the full-length masks are the reference's
    (temporal.py:37-51); torch's type promotion of ``t + delta`` is therefore
    inherited, including float32 for int64 times with a float delta.
    Raises RuntimeError when no pair exists (``torch.cat`` of an empty list, :53).

    ``max_source_stamps`` / ``budget_s`` (bench.py's bounded CPU sample only): stop the loop after that many
    distinct source time stamps / seconds -- every iteration still runs against the FULL stream, so the
    per-time-stamp cost is the full run's; ``stats`` receives ``stamps`` (done) and ``of_stamps`` (all).
    """
    import time as _time

    delta_t = torch.tensor(delta)                                    # :30
    positions = torch.arange(0, edge_index.size(1))                  # :31
    pieces = []
    stamps = torch.unique(timestamps, sorted=True)                   # :33
    done, t0 = 0, _time.perf_counter()
    for t in stamps:                                                 # :37
        heads = positions[timestamps == t]                           # :39-40
        tails = positions[(timestamps > t) & (timestamps <= t + delta_t)]  # :43-44
        if heads.numel() and tails.numel():                          # :46
            pairs = torch.cartesian_prod(heads, tails)               # :49
            keep = edge_index[1, pairs[:, 0]] == edge_index[0, pairs[:, 1]]  # :50
            pieces.append(pairs[keep])
        done += 1
        if (max_source_stamps is not None and done >= max_source_stamps) or \
                (budget_s is not None and _time.perf_counter() - t0 > budget_s):
            break
    if stats is not None:
        stats["stamps"], stats["of_stamps"] = done, int(stamps.numel())
    return torch.cat(pieces, dim=0).t().contiguous()                 # :53


# ----------------------------------------------------------------------------------------------
# closed forms (numpy) for parity tests at sizes where the reference's loops do not finish
# ----------------------------------------------------------------------------------------------

def lift_order_temporal_closed_form(edge_index: np.ndarray, timestamps: np.ndarray, delta) -> np.ndarray:
    """Same relation as ``lift_order_temporal`` via grouping by source node.

    Valid for time-sorted input with int64 timestamps and integer delta, or float64
    timestamps with float delta (the promotions under which the reference's compare
    is exact).  Output columns ascend in (e, f) like the reference's.
    """
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    t = np.asarray(timestamps)
    m = src.shape[0]
    if t.dtype.kind == "f":
        hi_time = t + np.float64(np.float32(delta)) if isinstance(delta, float) else t + delta
    else:
        hi_time = t + np.int64(delta)
    by_src = np.argsort(src, kind="stable")          # groups keep time order
    # composite key (source node, time rank) ascends along by_src, so each bound of the
    # window (t_e, t_e + delta] inside the group of dst(e) is one global searchsorted
    distinct_t = np.unique(t)
    span = np.int64(distinct_t.shape[0] + 1)
    t_rank = np.searchsorted(distinct_t, t).astype(np.int64)
    grouped = src[by_src].astype(np.int64) * span + t_rank[by_src]
    hi_rank = np.searchsorted(distinct_t, hi_time, side="right").astype(np.int64) - 1
    lo = np.searchsorted(grouped, dst.astype(np.int64) * span + t_rank, side="right")
    hi = np.searchsorted(grouped, dst.astype(np.int64) * span + hi_rank, side="right")
    cnt = np.maximum(hi - lo, 0)
    total = int(cnt.sum())
    if total == 0:
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
    e_col = np.repeat(np.arange(m, dtype=np.int64), cnt)
    first = np.cumsum(cnt) - cnt
    within = np.arange(total, dtype=np.int64) - np.repeat(first, cnt)
    f_col = by_src[np.repeat(lo, cnt) + within]
    return np.stack([e_col, f_col])


def unique_rows_closed_form(node_sequence: np.ndarray):
    """Lexicographic distinct rows + inverse (what ``torch.unique(dim=0)`` returns, lift_order.py:133)."""
    ns = np.asarray(node_sequence)
    order = np.lexsort(ns.T[::-1])
    s = ns[order]
    head = np.ones(s.shape[0], dtype=bool)
    if s.shape[0] > 1:
        head[1:] = (s[1:] != s[:-1]).any(axis=1)
    rank = np.cumsum(head) - 1
    inverse = np.empty(ns.shape[0], dtype=np.int64)
    inverse[order] = rank
    return s[head], inverse
