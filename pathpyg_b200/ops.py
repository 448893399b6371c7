"""Tensor-level entry points: torch owns device memory and the stream, the C ABI does the work.

Every function here takes CUDA tensors and returns freshly allocated CUDA tensors.  Host
tensors are handled one level up (``pathpyg_b200.algorithms``), which stages them to the device
and copies results back, so that the device follows the input as it does in the reference.
There is no CPU implementation: without a CUDA device or without the built library these raise.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

_DTYPE_CODES = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.int64: _lib.I64, torch.int32: _lib.I32}


def _ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pathpyg_b200.ops works on CUDA tensors only (no CPU fallback); got a CPU tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    return dev


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _edge_index_arg(edge_index: torch.Tensor) -> torch.Tensor:
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must have shape [2, E], got {tuple(edge_index.shape)}")
    if edge_index.dtype != torch.int64:
        edge_index = edge_index.long()
    # .as_subclass drops the EdgeIndex wrapper without copying
    return edge_index.as_subclass(torch.Tensor).contiguous()


# --------------------------------------------------------------------------------------- deferred counts
def _read_result(ws: torch.Tensor, dev: torch.device, what: str) -> int:
    """Collect {count, status} of a pending count -> fill operation (synchronises the stream)."""
    total, status = ctypes.c_int64(0), ctypes.c_int(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().ppg_result_read(_ptr(ws), ctypes.byref(total), ctypes.byref(status), _stream(dev)))
    if status.value & 1:
        raise ValueError(f"{what}: node id outside [0, num_nodes)")
    return total.value


def _lift_limit(ws: torch.Tensor, temporal: int, num_sources: int, num_nodes: int, limit: int, dev: torch.device) -> None:
    with torch.cuda.device(dev):
        _lib.check(_lib.load().ppg_lift_limit(_ptr(ws), temporal, num_sources, num_nodes, max(int(limit), 0), _stream(dev)))


# --------------------------------------------------------------------------------------- a2
class PendingLift:
    """lift_order_edge_index whose count pass is enqueued; ``finish()`` reads the size, allocates and fills.
    ``limit_sources``: only the columns whose source edge is < limit_sources (a prefix of the full lift).
    ``result_words`` (device int64 [2]: column count, status bits) lets a caller collect the count together with other
    pending results in ONE synchronisation and hand it to ``finish(total=..., status=...)``."""

    def __init__(self, ei, num_nodes, limit_sources=None):
        self.ei, self.num_nodes, self.dev, self.E = ei, num_nodes, ei.device, ei.size(1)
        self.ws = None
        if self.E:
            lib = _lib.load()
            with torch.cuda.device(self.dev):
                self.ws = _workspace(lib.ppg_lift_order_workspace_bytes(self.E, num_nodes), self.dev)
                _lib.check(lib.ppg_lift_order_count(_ptr(ei), self.E, num_nodes, _ptr(self.ws), self.ws.numel(), None, _stream(self.dev)))
            if limit_sources is not None and limit_sources < self.E:
                _lift_limit(self.ws, 0, self.E, self.num_nodes, limit_sources, self.dev)
            self.result_words = self.ws[:16].view(torch.int64)
        else:
            self.result_words = torch.zeros(2, dtype=torch.int64, device=self.dev)

    def finish(self, total: int | None = None, status: int = 0, allow_empty: bool = True) -> torch.Tensor:
        if total is None:
            total = _read_result(self.ws, self.dev, "lift_order_edge_index") if self.E else 0
        elif status & 1:
            raise ValueError("lift_order_edge_index: node id outside [0, num_nodes)")
        out = torch.empty((2, total), dtype=torch.int64, device=self.dev)
        if total:
            with torch.cuda.device(self.dev):
                _lib.check(_lib.load().ppg_lift_order_fill(_ptr(self.ws), self.E, self.num_nodes, total, _ptr(out), _stream(self.dev)))
        return out


def lift_order_edge_index_begin(edge_index: torch.Tensor, num_nodes: int, limit_sources: int | None = None) -> PendingLift:
    ei = _edge_index_arg(edge_index)
    _require_cuda(ei)
    return PendingLift(ei, int(num_nodes), limit_sources)


def lift_order_edge_index(edge_index: torch.Tensor, num_nodes: int, limit_sources: int | None = None) -> torch.Tensor:
    return lift_order_edge_index_begin(edge_index, num_nodes, limit_sources).finish()


# --------------------------------------------------------------------------------------- a3
def pair_attributes(edge_index: torch.Tensor, node_attribute: torch.Tensor, aggr: str,
                    index_bound: int | None = None) -> torch.Tensor:
    """``index_bound``: exclusive upper bound of the ids the caller vouches for (ids produced by this library's own
    lift are below the number of lifted-from edges); it is compared with ``len(node_attribute)`` on the host.  Without
    it the kernel reports ids outside the attribute tensor through a status word that is read back (one
    synchronisation) and raised as IndexError, like the reference's indexing (lift_order.py:31-45)."""
    if aggr not in _lib.PAIR_RULES:
        raise ValueError(f"Unknown aggregation method {aggr}")
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    dev = _require_cuda(ei, node_attribute)
    attr = node_attribute.contiguous()
    if attr.dim() != 1 or attr.dtype not in _DTYPE_CODES:
        raise TypeError(f"node_attribute must be a 1-D float32/float64/int64/int32 tensor, got {attr.dtype} {tuple(attr.shape)}")
    E = ei.size(1)
    out = torch.empty(E, dtype=attr.dtype, device=dev)
    if E == 0:
        return out
    if index_bound is not None and index_bound > attr.numel():
        raise IndexError(f"index {index_bound - 1} is out of bounds for dimension 0 with size {attr.numel()}")
    if attr.numel() == 0:
        raise IndexError("index is out of bounds for dimension 0 with size 0")
    status = None if index_bound is not None else torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_pair_attributes(_ptr(ei), E, _ptr(attr), attr.numel(), _DTYPE_CODES[attr.dtype],
                                           _lib.PAIR_RULES[aggr], _ptr(out), _ptr(status), _stream(dev)))
    if status is not None and int(status.item()) & 1:
        raise IndexError(f"index out of range in self: edge_index holds an id outside [0, {attr.numel()})")
    return out


# --------------------------------------------------------------------------------------- a1
def _time_mode(time: torch.Tensor, delta):
    """Mirror torch's promotion of ``timestamps <= t + torch.tensor(delta)`` (temporal.py:30,43)."""
    if isinstance(delta, torch.Tensor):
        delta = delta.item()
    if time.dtype == torch.int64:
        if isinstance(delta, int):
            return time, _lib.TIME_I64, int(delta), 0.0
        # 0-dim float32 delta: both sides of the comparison are evaluated in float32
        return time, _lib.TIME_I64_F32DELTA, 0, float(torch.tensor(float(delta), dtype=torch.float32))
    if time.dtype == torch.float64:
        d = float(delta) if isinstance(delta, int) else float(torch.tensor(float(delta), dtype=torch.float32))
        return time, _lib.TIME_F64, 0, d
    if time.dtype in (torch.int32, torch.int16, torch.int8, torch.uint8):
        return _time_mode(time.long(), delta)
    raise TypeError(f"time must be int64 or float64 (got {time.dtype}); float32 time stamps are not supported")


class PendingTemporalLift:
    """``limit_sources``: only the pairs whose source event is < limit_sources (a prefix of the full output);
    ``result_words``: see ``PendingLift``."""

    def __init__(self, ei, time, mode, delta_i, delta_f, num_nodes, grouped_ws=None, limit_sources=None):
        self.dev, self.m, self.num_nodes = ei.device, ei.size(1), num_nodes
        self.keep = (ei, time)  # inputs stay alive until the kernels have run
        lib = _lib.load()
        with torch.cuda.device(self.dev):
            if grouped_ws is None:
                self.ws = _workspace(lib.ppg_lift_temporal_workspace_bytes(self.m, num_nodes), self.dev)
            else:
                self.ws, mode = grouped_ws, mode | _lib.TIME_GROUPED
            _lib.check(lib.ppg_lift_temporal_count(_ptr(ei), _ptr(time), self.m, num_nodes, mode, delta_i, delta_f, _ptr(self.ws),
                                                   self.ws.numel(), None, _stream(self.dev)))
        if limit_sources is not None and limit_sources < self.m:
            _lift_limit(self.ws, 1, self.m, self.num_nodes, limit_sources, self.dev)
        self.result_words = self.ws[:16].view(torch.int64)

    def finish(self, total: int | None = None, status: int = 0, allow_empty: bool = False) -> torch.Tensor:
        """``allow_empty``: return a [2, 0] tensor instead of raising when no pair exists (one rank's slice of a stream)."""
        if total is None:
            total = _read_result(self.ws, self.dev, "lift_order_temporal")
        elif status & 1:
            raise ValueError("lift_order_temporal: node id outside [0, num_nodes)")
        if total == 0 and allow_empty:
            return torch.empty((2, 0), dtype=torch.int64, device=self.dev)
        if total == 0:
            raise _lib.EmptyLiftError("torch.cat(): expected a non-empty list of Tensors "
                                      "(lift_order_temporal: no time-respecting pair for this delta)")
        out = torch.empty((2, total), dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().ppg_lift_temporal_fill(_ptr(self.ws), self.m, self.num_nodes, total, _ptr(out), _stream(self.dev)))
        return out


def lift_order_temporal_group(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """First half of the temporal lift (grouping the events by source node): needs ``edge_index[0]`` only, so it can
    be enqueued while ``edge_index[1]`` and the time stamps are still being uploaded.  Returns the workspace to hand
    to ``lift_order_temporal_begin(..., grouped_ws=...)``."""
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    dev = _require_cuda(ei)
    m = ei.size(1)
    if m == 0 or num_nodes <= 0:
        raise _lib.EmptyLiftError("torch.cat(): expected a non-empty list of Tensors (lift_order_temporal: empty input)")
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_lift_temporal_workspace_bytes(m, int(num_nodes)), dev)
        _lib.check(lib.ppg_lift_temporal_group(_ptr(ei), m, int(num_nodes), _ptr(ws), ws.numel(), _stream(dev)))
    return ws


class PendingUnsortedTemporalLift:
    """Temporal lift of a stream whose time stamps are NOT in ascending order (``TemporalGraph.shuffle_time()``
    followed by a lift, as the reference's DBGNN tutorial does).  The reference's loop (temporal.py:33-53) does not
    care about the order of the events: it emits the pairs by ascending source time stamp, then ascending source
    POSITION, then ascending target POSITION.  The kernels need a time-ordered stream (binary searches over the
    time-ordered out-edges of a node), so the stream is put in order with a stable sort, lifted, and the pairs are
    mapped back to the caller's positions: the source order (time, position) is already the kernel's output order,
    the targets of one source are re-ordered by position with one radix sort over (source rank, target position)."""

    def __init__(self, ei, time, mode, delta_i, delta_f, num_nodes):
        self.perm = stable_argsort(time)
        self.inner = PendingTemporalLift(ei[:, self.perm].contiguous(), time[self.perm].contiguous(), mode, delta_i, delta_f,
                                         num_nodes)

    def finish(self) -> torch.Tensor:
        pairs = self.inner.finish()
        m = self.perm.numel()
        bits = max(1, (m - 1).bit_length())
        keys = (pairs[0] << bits) | self.perm[pairs[1]]
        del pairs
        sort_pairs_u64(keys, 2 * bits)
        return torch.stack([self.perm[keys >> bits], keys & ((1 << bits) - 1)])


def lift_order_temporal_begin(edge_index: torch.Tensor, time: torch.Tensor, delta, num_nodes: int,
                              grouped_ws: torch.Tensor | None = None, assume_sorted: bool = False,
                              limit_sources: int | None = None):
    """``assume_sorted``: the caller knows that ``time`` ascends (a ``TemporalGraph`` whose constructor put it in
    order); otherwise one device pass checks it, and an unordered stream takes ``PendingUnsortedTemporalLift``."""
    ei = _edge_index_arg(edge_index)
    _require_cuda(ei, time)
    time, mode, delta_i, delta_f = _time_mode(time.contiguous(), delta)
    if time.numel() != ei.size(1):
        raise ValueError("time and edge_index disagree on the number of edges")
    if ei.size(1) == 0 or num_nodes <= 0:
        raise _lib.EmptyLiftError("torch.cat(): expected a non-empty list of Tensors (lift_order_temporal: empty input)")
    if not assume_sorted and grouped_ws is None and time.numel() > 1 and not bool((time[1:] >= time[:-1]).all()):
        return PendingUnsortedTemporalLift(ei, time, mode, delta_i, delta_f, int(num_nodes))
    return PendingTemporalLift(ei, time, mode, delta_i, delta_f, int(num_nodes), grouped_ws, limit_sources)


def lift_order_temporal(edge_index: torch.Tensor, time: torch.Tensor, delta, num_nodes: int,
                        assume_sorted: bool = False, limit_sources: int | None = None, allow_empty: bool = False) -> torch.Tensor:
    pending = lift_order_temporal_begin(edge_index, time, delta, num_nodes, assume_sorted=assume_sorted,
                                        limit_sources=limit_sources)
    return pending.finish(allow_empty=allow_empty) if isinstance(pending, PendingTemporalLift) else pending.finish()


# --------------------------------------------------------------------------------------- a4
def rows_minmax(rows: torch.Tensor):
    lib = _lib.load()
    dev = _require_cuda(rows)
    M, k = rows.shape
    mins = (ctypes.c_int64 * k)()
    maxs = (ctypes.c_int64 * k)()
    asc = ctypes.c_int(0)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_rows_minmax_workspace_bytes(k), dev)
        _lib.check(lib.ppg_rows_minmax(_ptr(rows), M, k, _ptr(ws), ws.numel(), mins, maxs, ctypes.byref(asc), _stream(dev)))
    return list(mins), list(maxs), bool(asc.value)


def _unique_stage(rows: torch.Tensor, mins, bits):
    """One packed-key sort: returns (inverse, num_unique, workspace, total_bits)."""
    lib = _lib.load()
    dev = rows.device
    M, k = rows.shape
    total_bits = sum(bits)
    shifts, acc = [], total_bits
    for b in bits:  # column 0 is the most significant field
        acc -= b
        shifts.append(acc)
    inverse = torch.empty(M, dtype=torch.int64, device=dev)
    n = ctypes.c_int64(0)
    ws = _workspace(lib.ppg_unique_rows_workspace_bytes(M, total_bits), dev)
    _lib.check(lib.ppg_unique_rows_sort(_ptr(rows), M, k, (ctypes.c_int64 * k)(*mins), (ctypes.c_int * k)(*shifts),
                                        total_bits, _ptr(ws), ws.numel(), _ptr(inverse), ctypes.byref(n), _stream(dev)))
    return inverse, n.value, ws, total_bits


def unique_rows(node_sequence: torch.Tensor):
    """Distinct rows in lexicographic order and the inverse index (``torch.unique(dim=0, return_inverse=True)``)."""
    lib = _lib.load()
    rows = node_sequence.as_subclass(torch.Tensor)
    if rows.dim() != 2:
        raise ValueError(f"node_sequence must be 2-D, got shape {tuple(rows.shape)}")
    if rows.dtype != torch.int64:
        rows = rows.long()
    rows = rows.contiguous()
    dev = _require_cuda(rows)
    M, k = rows.shape
    if M == 0 or k == 0:
        return rows.new_empty((0 if M == 0 else 1, k)), torch.zeros(M, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        mins, maxs, ascending = rows_minmax(rows)
        if ascending:  # already the sorted distinct rows (e.g. arange(N)[:, None] of layer 1)
            return rows.clone(), torch.arange(M, dtype=torch.int64, device=dev)
        bits = [int(hi - lo).bit_length() for lo, hi in zip(mins, maxs)]
        if any(b > 63 for b in bits):
            raise ValueError("node_sequence value range exceeds 63 bits")
        cur, cur_mins, cur_bits = rows, mins, bits
        while sum(cur_bits) > 64 or cur.size(1) > 64:
            # rank refinement: the rank of a row is the rank of (rank of its leading columns, the rest)
            take, acc = 0, 0
            while take < min(len(cur_bits), 64) and acc + cur_bits[take] <= 64:
                acc += cur_bits[take]
                take += 1
            head_inv, head_n, _, _ = _unique_stage(cur[:, :take].contiguous(), cur_mins[:take], cur_bits[:take])
            cur = torch.cat([head_inv.unsqueeze(1), cur[:, take:]], dim=1)
            cur_mins = [0] + cur_mins[take:]
            cur_bits = [max(head_n - 1, 0).bit_length()] + cur_bits[take:]
        inverse, n, ws, total_bits = _unique_stage(cur, cur_mins, cur_bits)
        unique = torch.empty((n, k), dtype=torch.int64, device=dev)
        # `rep` (first occurrence of every distinct row) indexes the ORIGINAL rows in every stage
        _lib.check(lib.ppg_unique_rows_gather(_ptr(rows), M, k, _ptr(ws), total_bits, n, _ptr(unique), _stream(dev)))
    return unique, inverse


class PendingCoalesce:
    def __init__(self, ei, remap, num_nodes, edge_weight, reduce, return_inverse):
        self.dev, self.E, self.num_nodes = ei.device, ei.size(1), num_nodes
        self.edge_weight, self.reduce, self.return_inverse = edge_weight, reduce, return_inverse
        self.keep = (ei, remap)
        self.ws, self.inverse = None, None
        if return_inverse:
            self.inverse = torch.empty(self.E, dtype=torch.int64, device=self.dev)
        if self.E:
            lib = _lib.load()
            with torch.cuda.device(self.dev):
                self.ws = _workspace(lib.ppg_coalesce_workspace_bytes(self.E, num_nodes), self.dev)
                _lib.check(lib.ppg_coalesce_sort(_ptr(ei), self.E, _ptr(remap), 0 if remap is None else remap.numel(), num_nodes,
                                                 _ptr(self.ws), self.ws.numel(), _ptr(self.inverse), None, _stream(self.dev)))

    def finish(self):
        n_out = _read_result(self.ws, self.dev, "coalesce (EdgeIndex.validate)") if self.E else 0
        w_dtype = torch.float32 if self.edge_weight is None else self.edge_weight.dtype
        out_ei = torch.empty((2, n_out), dtype=torch.int64, device=self.dev)
        out_w = torch.empty(n_out, dtype=w_dtype, device=self.dev)
        if n_out:
            with torch.cuda.device(self.dev):
                _lib.check(_lib.load().ppg_coalesce_fill(_ptr(self.ws), self.E, self.num_nodes, n_out, _ptr(self.edge_weight),
                                                         _DTYPE_CODES[w_dtype], _lib.REDUCTIONS[self.reduce], _ptr(out_ei), _ptr(out_w),
                                                         _stream(self.dev)))
        if self.return_inverse:
            return out_ei, out_w, self.inverse
        return out_ei, out_w


def coalesce_begin(edge_index: torch.Tensor, remap: torch.Tensor | None, num_nodes: int,
                   edge_weight: torch.Tensor | None, reduce: str = "sum", return_inverse: bool = False) -> PendingCoalesce:
    if reduce not in _lib.REDUCTIONS:
        raise ValueError(f"Unknown reduce {reduce}")
    ei = _edge_index_arg(edge_index)
    _require_cuda(ei, remap, edge_weight)
    E = ei.size(1)
    if remap is not None:
        remap = remap.as_subclass(torch.Tensor).contiguous()
        if remap.dtype != torch.int64:
            remap = remap.long()
    if edge_weight is not None:
        edge_weight = edge_weight.contiguous()
        if edge_weight.dim() != 1 or edge_weight.numel() != E:
            raise ValueError("edge_weight must be 1-D with one entry per edge")
        if edge_weight.dtype not in _DTYPE_CODES:
            raise TypeError(f"edge_weight dtype {edge_weight.dtype} not supported (float32/float64/int64/int32)")
    return PendingCoalesce(ei, remap, int(num_nodes), edge_weight, reduce, return_inverse)


def coalesce(edge_index: torch.Tensor, remap: torch.Tensor | None, num_nodes: int,
             edge_weight: torch.Tensor | None, reduce: str = "sum", return_inverse: bool = False):
    """Map edge ids through ``remap`` (or not), merge duplicate (row, col) pairs reducing their weights;
    result is (row, col)-sorted.  ``edge_weight=None`` means unit float32 weights.  With
    ``return_inverse`` a third result gives, per input edge, the output edge it was merged into."""
    return coalesce_begin(edge_index, remap, num_nodes, edge_weight, reduce, return_inverse).finish()


def extend_rows(prev_rows: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """``cat([prev[ei[0]], prev[ei[1]][:, -1:]], 1)`` (multi_order_model.py:114) in one kernel."""
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    prev = prev_rows.as_subclass(torch.Tensor).contiguous()
    dev = _require_cuda(ei, prev)
    n, w = ei.size(1), prev.size(1)
    out = torch.empty((n, w + 1), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_extend_rows(_ptr(prev), prev.size(0), w, _ptr(ei), n, _ptr(out), _stream(dev)))
    return out


def sort_pairs_u64(keys: torch.Tensor, end_bit: int, time_passes: bool = False):
    """Stable in-place sort of the low ``end_bit`` bits of int64/uint64 keys; returns (perm int32, per-pass ms or None)."""
    lib = _lib.load()
    dev = _require_cuda(keys)
    n = keys.numel()
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    passes = max(1, -(-end_bit // 8))
    ms = (ctypes.c_float * passes)() if time_passes else None
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_sort_pairs_workspace_bytes(n, end_bit), dev)
        _lib.check(lib.ppg_sort_pairs_u64(_ptr(keys), _ptr(perm), n, end_bit, _ptr(ws), ws.numel(), ms, _stream(dev)))
    return perm, (list(ms) if time_passes else None)


def stable_argsort(values: torch.Tensor) -> torch.Tensor:
    """Permutation (int64) that puts a 1-D int64 / float64 CUDA tensor in ascending order, equal values keeping
    their input order.  Keys are made unsigned and as narrow as the value range allows, so time stamps spanning
    2^b cost ceil(b / 8) digit passes of the onesweep radix sort instead of 8 (TemporalGraph.__init__,
    reference core/temporal_graph.py:58)."""
    _require_cuda(values)
    v = values.as_subclass(torch.Tensor).contiguous()
    if v.dim() != 1:
        raise ValueError("stable_argsort expects a 1-D tensor")
    n = v.numel()
    if n >= 1 << 31:
        raise ValueError("stable_argsort: more than 2^31 - 1 elements")
    if n < 2:
        return torch.arange(n, device=v.device)
    if v.dtype == torch.float64:
        if bool(torch.isnan(v).any()):
            raise ValueError("stable_argsort: NaN time stamps have no order")
        bits = (v + 0.0).view(torch.int64)                                  # + 0.0 turns -0.0 into +0.0
        # order-preserving map to unsigned: negative values flip all bits, the others flip the sign bit
        keys, end_bit = bits ^ ((bits >> 63) | torch.tensor(-1 << 63, dtype=torch.int64, device=v.device)), 64
    elif v.dtype in (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8):
        v = v.long()
        lo, hi = int(v.min()), int(v.max())
        span = hi - lo
        if span >= 1 << 63:
            keys, end_bit = v ^ torch.tensor(-1 << 63, dtype=torch.int64, device=v.device), 64
        else:
            keys, end_bit = v - lo, max(1, span.bit_length())
    else:
        raise TypeError(f"stable_argsort supports int64 and float64 (got {v.dtype})")
    perm, _ = sort_pairs_u64(keys, end_bit)  # keys is a fresh tensor in every branch (sorted in place)
    return perm.long()


# --------------------------------------------------------------------------------------- e: cross-partition exchange
class RoutePlan:
    """Stable partition of the line-graph edges of one level by the rank that owns their source row
    (``csrc/exchange.cu``).  ``counts`` [world] (device int64) is available right after construction."""

    def __init__(self, line_index: torch.Tensor, node_info: torch.Tensor | None, offsets: torch.Tensor, world: int):
        lib = _lib.load()
        self.li = _edge_index_arg(line_index)
        self.dev = _require_cuda(self.li, node_info, offsets)
        self.E, self.world = self.li.size(1), int(world)
        self.node_info = None if node_info is None else node_info.contiguous()
        self.offsets = offsets.contiguous()
        if self.offsets.dtype != torch.int64 or self.offsets.numel() != world + 1:
            raise ValueError("offsets must be int64 with world + 1 entries")
        self.counts = torch.empty(world, dtype=torch.int64, device=self.dev)
        self.slot = self.last = None
        with torch.cuda.device(self.dev):
            self.ws = _workspace(lib.ppg_route_workspace_bytes(self.E), self.dev)
            _lib.check(lib.ppg_route_count(_ptr(self.li), self.E, _ptr(self.node_info), _ptr(self.offsets), self.world,
                                           _ptr(self.ws), self.ws.numel(), _ptr(self.counts), _stream(self.dev)))

    def pack(self, weights: torch.Tensor | None, own_prefix: int, peer_slots: list[int] | None = None):
        """Records [E, 2] int64 (16 bytes each) grouped by destination rank, edge order inside a destination.
        ``peer_slots``: per destination rank the DEVICE address (mapped peer memory) of the first record slot reserved
        for this sender in that rank's receive buffer -- the kernel then stores straight into the owners' memory and
        nothing is returned."""
        if weights is not None:
            weights = weights.contiguous()
            if weights.dtype != torch.float32 or weights.numel() != self.E:
                raise TypeError("route weights must be float32 with one entry per edge")
        records = torch.empty((self.E, 2), dtype=torch.int64, device=self.dev) if peer_slots is None else None
        peers = None if peer_slots is None else (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(int(p)) for p in peer_slots])
        self.slot = torch.empty(self.E, dtype=torch.int32, device=self.dev)
        self.last = torch.empty(self.E, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().ppg_route_pack(_ptr(self.li), self.E, _ptr(self.node_info), _ptr(weights), int(own_prefix),
                                                  _ptr(self.offsets), self.world, _ptr(self.ws), _ptr(records), peers,
                                                  _ptr(self.slot), _ptr(self.last), _stream(self.dev)))
        return records

    def unpack(self, back: torch.Tensor, edge_offsets: torch.Tensor) -> torch.Tensor:
        """``back`` [E] int32: merged-edge index of every record as returned by its owner; ``edge_offsets`` [world + 1]
        device int64.  Result [E] int64: ``global merged-edge id << 32 | last node`` per edge = next level's node_info."""
        out = torch.empty(self.E, dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().ppg_route_unpack(_ptr(self.ws), self.E, _ptr(back), _ptr(self.slot), _ptr(self.last),
                                                    _ptr(edge_offsets.contiguous()), self.world, _ptr(out), _stream(self.dev)))
        return out


def route_plan(line_index, node_info, offsets, world) -> RoutePlan:
    return RoutePlan(line_index, node_info, offsets, world)


class PendingMerge:
    """Owner side of the exchange: merge the received records of the rows [row_lo, row_lo + rows_owned).
    ``result_words`` (device int64 [2]: merged count, status bits) is valid once the stream has run;
    ``inverse`` [R] int32 = merged-edge index of every record."""

    def __init__(self, records: torch.Tensor, row_lo: int, rows_owned: int, total_nodes: int):
        lib = _lib.load()
        self.records = records.contiguous()
        self.dev = _require_cuda(self.records)
        self.R = self.records.size(0)
        self.row_lo, self.rows_owned, self.total_nodes = int(row_lo), int(rows_owned), int(total_nodes)
        self.inverse = torch.empty(self.R, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            self.ws = _workspace(lib.ppg_merge_records_workspace_bytes(self.R, self.rows_owned, self.total_nodes), self.dev)
            _lib.check(lib.ppg_merge_records_sort(_ptr(self.records), self.R, self.row_lo, self.rows_owned, self.total_nodes,
                                                  _ptr(self.ws), self.ws.numel(), _ptr(self.inverse), _stream(self.dev)))
        self.result_words = self.ws[:16].view(torch.int64)

    def finish(self, num_out: int):
        """(edge_index [2, num_out] global ids, edge_weight float32, last node int64) of the merged edges."""
        out_ei = torch.empty((2, num_out), dtype=torch.int64, device=self.dev)
        out_w = torch.empty(num_out, dtype=torch.float32, device=self.dev)
        out_last = torch.empty(num_out, dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().ppg_merge_records_fill(_ptr(self.ws), _ptr(self.records), self.R, self.row_lo, self.rows_owned,
                                                          self.total_nodes, int(num_out), _ptr(out_ei), _ptr(out_w), _ptr(out_last),
                                                          _stream(self.dev)))
        return out_ei, out_w, out_last


def merge_records_begin(records, row_lo, rows_owned, total_nodes) -> PendingMerge:
    return PendingMerge(records, row_lo, rows_owned, total_nodes)


def extend_owned_rows(prev_rows: torch.Tensor, prev_row_lo: int, src_ids: torch.Tensor, last: torch.Tensor) -> torch.Tensor:
    """``cat([prev_rows[src_ids - prev_row_lo], last[:, None]], 1)`` in one kernel."""
    prev = prev_rows.contiguous()
    dev = _require_cuda(prev, src_ids, last)
    n, w = src_ids.numel(), prev.size(1)
    out = torch.empty((n, w + 1), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().ppg_extend_owned_rows(_ptr(prev), w, int(prev_row_lo), _ptr(src_ids.contiguous()), _ptr(last.contiguous()),
                                                     n, _ptr(out), _stream(dev)))
    return out


# --------------------------------------------------------------------------------------- a7 / a9 bookkeeping
def counts_to_offsets(counts: torch.Tensor, min_count: int = 0, what: str = "counts"):
    """(offsets [n + 1] int64, total): exclusive prefix sums of int64 counts (``torch.cumsum`` with a leading zero);
    counts below ``min_count`` raise ValueError.  One host synchronisation (the total sizes what follows)."""
    lib = _lib.load()
    dev = _require_cuda(counts)
    counts = counts.contiguous().long()
    n = counts.numel()
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_counts_to_offsets_workspace_bytes(n), dev)
        _lib.check(lib.ppg_counts_to_offsets(_ptr(counts), n, int(min_count), _ptr(ws), ws.numel(), _ptr(offsets), _stream(dev)))
        total, status = ctypes.c_int64(0), ctypes.c_int(0)
        _lib.check(lib.ppg_result_read(_ptr(ws), ctypes.byref(total), ctypes.byref(status), _stream(dev)))
    if status.value & 1:
        raise ValueError(f"{what}: a count is below {min_count}")
    return offsets, total.value


def repeat_by_count(values: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """``values.repeat_interleave(counts)`` for 1-D values of 4 or 8 bytes per element (multi_order_model.py:217,402)."""
    lib = _lib.load()
    dev = _require_cuda(values, counts)
    values = values.contiguous()
    if values.dim() != 1 or values.numel() != counts.numel() or values.element_size() not in (4, 8):
        raise ValueError(f"repeat_by_count: values {tuple(values.shape)} {values.dtype}, counts {tuple(counts.shape)}")
    offsets, total = counts_to_offsets(counts, 0, "repeat_interleave")
    out = torch.empty(total, dtype=values.dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_expand_offsets(_ptr(offsets), counts.numel(), total, _ptr(values), values.element_size(), _ptr(out),
                                          None, _stream(dev)))
    return out


def walk_chain(lengths: torch.Tensor, base: int = 0) -> torch.Tensor:
    """edge_index [2, sum(lengths) - len(lengths)] of the links p -> p + 1 (+ base) inside walks laid end to end
    (path_data.py:139-159: arange, stack, drop the links between consecutive walks)."""
    lib = _lib.load()
    dev = _require_cuda(lengths)
    offsets, total = counts_to_offsets(lengths, 1, "walk lengths")
    w = lengths.numel()
    out = torch.empty((2, total - w), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_walk_chain(_ptr(offsets), w, total, int(base), _ptr(out), _stream(dev)))
    return out


def bincount(ids: torch.Tensor, num_bins: int | None = None) -> torch.Tensor:
    """``torch.bincount(ids, minlength=num_bins)`` with exactly ``num_bins`` bins (default: largest id + 1); ids outside
    [0, num_bins) raise ValueError."""
    lib = _lib.load()
    dev = _require_cuda(ids)
    ids = ids.reshape(-1).contiguous().long()
    if num_bins is None:
        num_bins = int(ids.max()) + 1 if ids.numel() else 0
    counts = torch.empty(int(num_bins), dtype=torch.int64, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_bincount(_ptr(ids), ids.numel(), int(num_bins), _ptr(counts), _ptr(status), _stream(dev)))
    if int(status[0]) & 1:
        raise ValueError(f"bincount: id outside [0, {num_bins})")
    return counts


# --------------------------------------------------------------------------------------- a10 / a11
class TargetGroupedEdges:
    """CSC view of an edge list: incoming edges of every target node, original order inside a target."""

    __slots__ = ("colptr", "src", "eid", "num_sources", "num_targets", "val", "self_val", "val_edge", "edge_index",
                 "_transposed", "pending_ws")

    def __init__(self, colptr, src, eid, num_sources, num_targets):
        self.colptr, self.src, self.eid = colptr, src, eid
        self.num_sources, self.num_targets = num_sources, num_targets
        self.val = None       # per-slot coefficient (None = 1)
        self.self_val = None  # per-target coefficient of the node's own row (None = no self term)
        self.val_edge = None  # the per-slot coefficient indexed by original edge id (kept for the backward view)
        self.edge_index = None
        self._transposed = None
        self.pending_ws = None  # workspace whose status word has not been checked yet (deferred csc_build)

    def check(self) -> None:
        """Raise ValueError if the deferred build saw a node id out of range (synchronises the stream)."""
        if self.pending_ws is not None:
            ws, self.pending_ws = self.pending_ws, None
            _read_result(ws, ws.device, "csc_build")

    def transposed(self) -> "TargetGroupedEdges":
        """The same edges grouped by SOURCE node (what the backward pass reduces over): A^T."""
        if self._transposed is None:
            t = csc_build(self.edge_index.flip(0), self.num_targets, self.num_sources)
            if self.val_edge is not None:
                t.val = gather_f32(self.val_edge, t.eid)
            t.self_val = self.self_val
            t._transposed = self
            self._transposed = t
        return self._transposed


def csc_build(edge_index: torch.Tensor, num_sources: int, num_targets: int, defer_check: bool = False) -> TargetGroupedEdges:
    """``defer_check``: only enqueue the kernels (no host synchronisation); the caller validates the ids later with
    ``graph.check()`` -- out-of-range ids are clamped meanwhile, so consumers stay memory-safe."""
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    dev = _require_cuda(ei)
    E = ei.size(1)
    colptr = torch.empty(num_targets + 1, dtype=torch.int32, device=dev)
    src = torch.empty(E, dtype=torch.int32, device=dev)
    eid = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_csc_workspace_bytes(E, num_targets), dev)
        build = lib.ppg_csc_build_async if defer_check else lib.ppg_csc_build
        _lib.check(build(_ptr(ei), E, num_sources, num_targets, _ptr(ws), ws.numel(), _ptr(colptr), _ptr(src), _ptr(eid),
                         _stream(dev)))
    g = TargetGroupedEdges(colptr, src, eid, num_sources, num_targets)
    g.edge_index = ei
    if defer_check:
        g.pending_ws = ws
    return g


def gcn_prepare(edge_index: torch.Tensor, edge_weight: torch.Tensor | None, num_nodes: int,
                keep_edge_values: bool = False, defer_check: bool = False) -> TargetGroupedEdges:
    """CSC view + symmetric GCN normalisation with remaining self-loops (PyG gcn_norm)."""
    lib = _lib.load()
    g = csc_build(edge_index, num_nodes, num_nodes, defer_check=defer_check)
    dev = g.colptr.device
    E = g.src.numel()
    if edge_weight is not None:
        edge_weight = edge_weight.contiguous().float()
    dis = torch.empty(num_nodes, dtype=torch.float32, device=dev)
    g.val = torch.empty(E, dtype=torch.float32, device=dev)
    g.self_val = torch.empty(num_nodes, dtype=torch.float32, device=dev)
    if keep_edge_values:
        g.val_edge = torch.empty(E, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_gcn_norm(_ptr(g.colptr), _ptr(g.src), _ptr(g.eid), _ptr(edge_weight), num_nodes, E, _ptr(dis),
                                    _ptr(g.val), _ptr(g.self_val), _ptr(g.val_edge), _stream(dev)))
    return g


def gather_f32(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[i] = src[idx[i]] (float32 values, int32 indices)."""
    lib = _lib.load()
    dev = _require_cuda(src, idx)
    out = torch.empty(idx.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_gather_f32(_ptr(src), _ptr(idx), idx.numel(), _ptr(out), _stream(dev)))
    return out


def act_backward(dy: torch.Tensor, y: torch.Tensor | None, act: int, rowscale: torch.Tensor | None = None,
                 want_dpre: bool = True):
    """(dPre, dPreScaled, colsum): dPre = dy * act'(.), dPreScaled = rowscale[:, None] * dPre (None without
    rowscale), colsum = column sums of dPreScaled if rowscale is given else of dPre."""
    lib = _lib.load()
    dev = _require_cuda(dy, y, rowscale)
    dy = dy.contiguous()
    M, H = dy.shape
    dpre = torch.empty_like(dy) if want_dpre else None
    scaled = torch.empty_like(dy) if rowscale is not None else None
    colsum = torch.empty(H, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_act_backward_workspace_bytes(M, H), dev)
        _lib.check(lib.ppg_act_backward(_ptr(dy), _ptr(y), _ptr(rowscale), M, H, act, _ptr(dpre), _ptr(scaled), _ptr(colsum),
                                        _ptr(ws), ws.numel(), _stream(dev)))
    return dpre, scaled, colsum


def atb(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a^T @ b for tall-skinny float32 a [M,H], b [M,F] (weight gradients: reduction over the node dimension)."""
    lib = _lib.load()
    dev = _require_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    M, H = a.shape
    F = b.size(1)
    if b.size(0) != M:
        raise ValueError(f"shape mismatch: {tuple(a.shape)}.T @ {tuple(b.shape)}")
    out = torch.empty((H, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_atb_workspace_bytes(M, H, F), dev)
        _lib.check(lib.ppg_atb(_ptr(a), _ptr(b), M, H, F, _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
    return out


def colptr_counts(g: TargetGroupedEdges) -> torch.Tensor:
    lib = _lib.load()
    dev = g.colptr.device
    out = torch.empty(g.num_targets, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_colptr_counts(_ptr(g.colptr), g.num_targets, _ptr(out), _stream(dev)))
    return out


def spmm_csc(g: TargetGroupedEdges, x: torch.Tensor, bias: torch.Tensor | None = None, act: int = _lib.ACT_NONE) -> torch.Tensor:
    lib = _lib.load()
    dev = _require_cuda(x, g.colptr)
    x = x.contiguous()
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError("spmm_csc expects a 2-D float32 feature matrix")
    if x.size(0) < g.num_sources:
        raise ValueError(f"feature matrix has {x.size(0)} rows but the graph has {g.num_sources} source nodes")
    F = x.size(1)
    out = torch.empty((g.num_targets, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_spmm_csc(_ptr(g.colptr), _ptr(g.src), _ptr(g.val), _ptr(g.self_val), _ptr(x), g.num_targets, F,
                                    _ptr(bias), act, _ptr(out), _stream(dev)))
    return out


def linear(a1: torch.Tensor, w1: torch.Tensor, bias: torch.Tensor | None = None, act: int = _lib.ACT_NONE,
           a2: torch.Tensor | None = None, w2: torch.Tensor | None = None, rowscale: torch.Tensor | None = None) -> torch.Tensor:
    """act(a1 @ w1.T + rowscale[:, None] * (a2 @ w2.T + bias))."""
    lib = _lib.load()
    dev = _require_cuda(a1, w1, a2, w2, bias, rowscale)
    a1, w1 = a1.contiguous(), w1.contiguous()
    M, K1 = a1.shape
    N = w1.size(0)
    if w1.size(1) != K1:
        raise ValueError(f"shape mismatch: {tuple(a1.shape)} @ {tuple(w1.shape)}.T")
    K2 = 0
    if a2 is not None:
        a2, w2 = a2.contiguous(), w2.contiguous()
        K2 = a2.size(1)
        if a2.size(0) != M or w2.shape != (N, K2):
            raise ValueError("shape mismatch in the second operand pair")
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_linear(_ptr(a1), _ptr(w1), M, K1, _ptr(a2), _ptr(w2), K2, _ptr(bias), _ptr(rowscale), N, act,
                                  _ptr(out), _stream(dev)))
    return out


def fused_supported(in_width: int, out_width: int) -> bool:
    return bool(_lib.load().ppg_gcn_fused_supported(in_width, out_width))


def gcn_layer_fused(g: TargetGroupedEdges, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None,
                    act: int = _lib.ACT_NONE) -> torch.Tensor:
    """act((A_norm x) W^T + b) in one kernel (widths in {16, 32, 64})."""
    lib = _lib.load()
    dev = _require_cuda(x, weight, bias, g.colptr)
    x, weight = x.contiguous(), weight.contiguous()
    H, F = weight.shape
    if x.size(1) != F or x.size(0) < g.num_sources:
        raise ValueError(f"shape mismatch: features {tuple(x.shape)}, weight {tuple(weight.shape)}, graph sources {g.num_sources}")
    out = torch.empty((g.num_targets, H), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_gcn_layer_fused(_ptr(g.colptr), _ptr(g.src), _ptr(g.val), _ptr(g.self_val), _ptr(x), _ptr(weight),
                                           _ptr(bias), g.num_targets, F, H, act, _ptr(out), _stream(dev)))
    return out


def tc_supported(in_width: int, out_width: int) -> bool:
    """Widths for which the layer runs its dense transform on the tcgen05 tensor cores.
    ``PPG_GCN_FMA=1`` in the environment selects the FMA-pipe kernel instead (A/B comparison)."""
    if os.environ.get("PPG_GCN_FMA", "0") == "1":
        return False
    return bool(_lib.load().ppg_gcn_tc_supported(in_width, out_width))


def gcn_layer_tc(g: TargetGroupedEdges, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None,
                 act: int = _lib.ACT_NONE) -> torch.Tensor:
    """act((A_norm x) W^T + b): segment-reduce gather + tcgen05 (3xTF32) transform in one kernel."""
    lib = _lib.load()
    dev = _require_cuda(x, weight, bias, g.colptr)
    x, weight = x.contiguous(), weight.contiguous()
    H, F = weight.shape
    if x.size(1) != F or x.size(0) < g.num_sources:
        raise ValueError(f"shape mismatch: features {tuple(x.shape)}, weight {tuple(weight.shape)}, graph sources {g.num_sources}")
    out = torch.empty((g.num_targets, H), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_gcn_layer_tc(_ptr(g.colptr), _ptr(g.src), _ptr(g.val), _ptr(g.self_val), _ptr(x), _ptr(weight),
                                        _ptr(bias), g.num_targets, g.src.numel(), F, H, act, _ptr(out), _stream(dev)))
    return out


def bipartite_fused(g: TargetGroupedEdges, x_h: torch.Tensor, x: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor,
                    bias12: torch.Tensor, act: int = _lib.ACT_NONE) -> torch.Tensor:
    """act((sum_u x_h[u]) W1^T + indeg (x W2^T + b1 + b2)) in one kernel (widths in {16, 32, 64})."""
    lib = _lib.load()
    dev = _require_cuda(x_h, x, w1, w2, bias12, g.colptr)
    x_h, x, w1, w2 = x_h.contiguous(), x.contiguous(), w1.contiguous(), w2.contiguous()
    H, F = w1.shape
    out = torch.empty((g.num_targets, H), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_bipartite_fused(_ptr(g.colptr), _ptr(g.src), _ptr(x_h), _ptr(x), _ptr(w1), _ptr(w2), _ptr(bias12),
                                           g.num_targets, F, H, act, _ptr(out), _stream(dev)))
    return out


# --------------------------------------------------------------------------------------- layer consumers
def sorted_ids_ptr(sorted_ids: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """CSR pointer (int32, ``num_nodes + 1`` entries) of an ascending int64 id column."""
    lib = _lib.load()
    ids = sorted_ids.as_subclass(torch.Tensor).contiguous()
    dev = _require_cuda(ids)
    ptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_sorted_ids_ptr(_ptr(ids), ids.numel(), num_nodes, _ptr(ptr), _stream(dev)))
    return ptr


def segment_sum(ptr: torch.Tensor, weights: torch.Tensor | None, perm: torch.Tensor | None = None) -> torch.Tensor:
    """float32 sums of ``weights`` over the slots of every segment of ``ptr`` (counts if ``weights`` is None)."""
    lib = _lib.load()
    dev = _require_cuda(ptr, weights, perm)
    n = ptr.numel() - 1
    if weights is not None:
        weights = weights.contiguous().float()
    slots = (perm.numel() if perm is not None else weights.numel()) if weights is not None else 0
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_segment_sum_workspace_bytes(slots), dev)
        _lib.check(lib.ppg_segment_sum(_ptr(ptr), _ptr(perm), _ptr(weights), n, slots, _ptr(ws), ws.numel(), _ptr(out), _stream(dev)))
    return out


def edge_ratio(ids: torch.Tensor, weights: torch.Tensor | None, denom: torch.Tensor) -> torch.Tensor:
    """out[e] = (weights[e] or 1) / denom[ids[e]] (float32)."""
    lib = _lib.load()
    ids = ids.as_subclass(torch.Tensor).contiguous()
    dev = _require_cuda(ids, weights, denom)
    if weights is not None:
        weights = weights.contiguous().float()
    denom = denom.contiguous().float()
    out = torch.empty(ids.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_edge_ratio(_ptr(ids), _ptr(weights), _ptr(denom), ids.numel(), _ptr(out), _stream(dev)))
    return out


def walk_counts(edge_index: torch.Tensor, num_nodes: int, max_len: int):
    """([#walks with k edges], [#nodes starting one]) for k = 1..max_len, as Python ints."""
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    dev = _require_cuda(ei)
    walks, sources = (ctypes.c_int64 * max_len)(), (ctypes.c_int64 * max_len)()
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_walk_counts_workspace_bytes(num_nodes, max_len), dev)
        _lib.check(lib.ppg_walk_counts(_ptr(ei), ei.size(1), num_nodes, max_len, _ptr(ws), ws.numel(), walks, sources, _stream(dev)))
    return list(walks), list(sources)


def weighted_log_sum(freq: torch.Tensor, prob: torch.Tensor, idx: torch.Tensor | None = None,
                     idx2: torch.Tensor | None = None) -> float:
    """sum_i freq[i] * log(prob[j]) with j = i, idx[i] or idx2[idx[i]]; fp32 terms, fp64 fixed-order accumulation."""
    lib = _lib.load()
    dev = _require_cuda(freq, prob, idx, idx2)
    freq, prob = freq.contiguous().float(), prob.contiguous().float()
    idx = None if idx is None else idx.as_subclass(torch.Tensor).contiguous().long()
    idx2 = None if idx2 is None else idx2.as_subclass(torch.Tensor).contiguous().long()
    out = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_weighted_log_sum_workspace_bytes(), dev)
        try:
            _lib.check(lib.ppg_weighted_log_sum(_ptr(freq), _ptr(prob), _ptr(idx), _ptr(idx2), freq.numel(), prob.numel(),
                                                0 if idx2 is None else idx2.numel(), _ptr(ws), ws.numel(), ctypes.byref(out), _stream(dev)))
        except ValueError as e:  # torch raises IndexError for an out-of-range index
            raise IndexError(str(e)) from None
    return out.value


# --------------------------------------------------------------------------------------- shortest time-respecting paths
def temporal_paths(edge_index: torch.Tensor, event_graph: torch.Tensor | None, num_nodes: int,
                   max_workspace_bytes: int = 8 << 30):
    """(dist [n, n] float64, pred [n, n] int64) on the device: bit-parallel multi-source BFS over the event graph,
    sources processed in chunks (multiples of 32) sized so that the frontier state stays below
    ``max_workspace_bytes``."""
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    dev = _require_cuda(ei, event_graph)
    m, n = ei.size(1), int(num_nodes)
    eg = _edge_index_arg(event_graph) if event_graph is not None and event_graph.numel() else None
    pairs = 0 if eg is None else eg.size(1)
    dist = torch.empty((n, n), dtype=torch.float64, device=dev)
    pred = torch.empty((n, n), dtype=torch.int64, device=dev)
    if n == 0:
        return dist, pred
    per_word = 12 * max(m, 1) + 32 * 8 * n                      # frontier + visited + next words, best rows
    chunk = max(32, min((n + 31) // 32 * 32, max_workspace_bytes // per_word * 32))
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_temporal_paths_workspace_bytes(m, n, min(chunk, n)), dev)
        for s0 in range(0, n, chunk):
            s1 = min(n, s0 + chunk)
            _lib.check(lib.ppg_temporal_paths(_ptr(ei), m, n, _ptr(eg), pairs, s0, s1, _ptr(ws), ws.numel(),
                                              ctypes.c_void_p(dist.data_ptr() + s0 * n * 8),
                                              ctypes.c_void_p(pred.data_ptr() + s0 * n * 8), None, _stream(dev)))
    return dist, pred


def temporal_closeness(dist: torch.Tensor) -> torch.Tensor:
    """closeness[v] = sum over x != v of (n - 1) / dist[x, v] (float64, added in ascending x)."""
    lib = _lib.load()
    dev = _require_cuda(dist)
    dist = dist.contiguous()
    n = dist.size(0)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ppg_temporal_closeness(_ptr(dist), n, _ptr(out), _stream(dev)))
    return out


def temporal_betweenness(edge_index: torch.Tensor, time: torch.Tensor, event_graph: torch.Tensor, num_nodes: int,
                         max_workspace_bytes: int = 4 << 30) -> torch.Tensor:
    """Temporal betweenness of every node (float64 [n]) from the time-sorted events and their event graph: Brandes'
    dependency accumulation on the event DAG, one CTA per source node, sources in batches."""
    lib = _lib.load()
    ei = _edge_index_arg(edge_index)
    eg = _edge_index_arg(event_graph)
    dev = _require_cuda(ei, time, eg)
    m, n = ei.size(1), int(num_nodes)
    bw = torch.zeros(n, dtype=torch.float64, device=dev)
    if m == 0 or n == 0:
        return bw
    # index plumbing: time groups, CSR / CSC of the event graph, events grouped by target, source list
    _, counts = torch.unique_consecutive(time, return_counts=True)
    group_off = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=dev)
    group_off[1:] = torch.cumsum(counts, 0)
    succ_ptr = sorted_ids_ptr(eg[0], m)
    preds = csc_build(eg, m, m)
    incoming = csc_build(ei, n, n)
    sources = torch.unique(ei[0]).int()
    per_source = 28 * m + 20 * n
    batch_cap = max(1, min(296, max_workspace_bytes // max(per_source, 1)))
    with torch.cuda.device(dev):
        ws = _workspace(lib.ppg_temporal_betweenness_workspace_bytes(m, n, min(batch_cap, sources.numel())), dev)
        for b0 in range(0, sources.numel(), batch_cap):
            batch = sources[b0:b0 + batch_cap].contiguous()
            _lib.check(lib.ppg_temporal_betweenness(_ptr(ei), m, n, _ptr(group_off), group_off.numel() - 1, _ptr(succ_ptr),
                                                    _ptr(eg[1].contiguous()), _ptr(preds.colptr), _ptr(preds.src),
                                                    _ptr(incoming.colptr), _ptr(incoming.eid), _ptr(batch), batch.numel(),
                                                    _ptr(ws), ws.numel(), _ptr(bw), _stream(dev)))
    return bw
