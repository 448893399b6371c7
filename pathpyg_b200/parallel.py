"""Multi-GPU plumbing of the lift -> DBGNN path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests of the host logic).

The reference (pathpy/pathpyG) is single-process, single-device: nothing here has a counterpart in
``/root/reference``; the contract is that the distributed result, gathered, is IDENTICAL (bit-exact) to
``MultiOrderModel.from_temporal_graph`` / ``from_path_data`` on one device.

* ``allreduce_gradients``          -- data-parallel DBGNN training (BASELINE config 4): walks are sharded by walk id,
  every rank lifts its own shard locally (independent units, no exchange) and trains on its own layers;
  one flat all-reduce carries all weight gradients.
* ``shard_walks``                  -- contiguous walk-id ranges balanced by node count.
* ``distributed_temporal_layers``  -- BASELINE config 5: the time-sorted edge stream is split into contiguous
  ranges; a ghost zone of (K-1)*delta is exchanged once (all-to-all-v); every rank lifts its extended range
  locally and counts the causal paths whose FIRST edge it owns (every path is counted exactly once); per
  order ONE all-to-all-v carries the line-graph edges to the owner of their source row, where they are merged
  (local radix sort + run detection); the global index of a merged edge -- all-gather of counts + exclusive
  scan -- is returned to the senders and IS the De Bruijn node id of the next order.

The local compute goes through ``pathpyg_b200.ops`` (CUDA only).  The CPU tests inject an object with the
same functions backed by the oracle, so that the partition / exchange / id-assignment logic is
exercised with gloo at world size 2 without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------
# data-parallel training
# ------------------------------------------------------------------------------------------------
def allreduce_gradients(module: torch.nn.Module, group=None, average: bool = True) -> None:
    """Sum (or average) the gradients of all parameters over the ranks with ONE flat all-reduce.
    DBGNN has a few tens of thousands of weights: the collective is latency-bound, so everything goes
    into a single bucket."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    offset = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[offset:offset + n].view_as(p.grad))
        offset += n


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Start all replicas from rank ``src``'s weights."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = list(module.parameters()) + list(module.buffers())
    if not tensors:
        return
    flat = torch.cat([t.detach().reshape(-1) for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    offset = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[offset:offset + n].view_as(t))
            offset += n


def shard_walks(lengths: torch.Tensor, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of walk ids for ``rank``, balanced by the number of nodes in the walks."""
    total = int(lengths.sum())
    ends = torch.cumsum(lengths, 0)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(torch.searchsorted(ends, torch.tensor(total * r // world, device=ends.device), right=False)))
    bounds.append(lengths.numel())
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds[rank], bounds[rank + 1]


# ------------------------------------------------------------------------------------------------
# collectives on variable-size tensors
# ------------------------------------------------------------------------------------------------
def _all_gather_int(values: list[int], device, group) -> torch.Tensor:
    """[world, len(values)] int64 on the host."""
    world = dist.get_world_size(group)
    mine = torch.tensor(values, dtype=torch.int64, device=device)
    out = torch.empty((world, len(values)), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine.unsqueeze(0), group=group) if device.type == "cuda" else \
        dist.all_gather(list(out.unbind(0)), mine, group=group)
    return out.cpu()


def all_to_all_rows(rows: torch.Tensor, send_counts: list[int], group=None,
                    recv_counts: list[int] | None = None) -> tuple[torch.Tensor, list[int]]:
    """Variable all-to-all of the leading dimension: ``rows`` is ordered by destination rank,
    ``send_counts[p]`` rows go to rank p.  Returns (received rows ordered by source rank, recv_counts).
    ``recv_counts``: pass them when they are already known (the answer to an earlier exchange travels the same
    routes backwards) -- saves the all-gather of the counts and its host synchronisation."""
    world = dist.get_world_size(group)
    dev = rows.device
    if recv_counts is None:
        counts = _all_gather_int(send_counts, dev, group)          # counts[q, p] = rows q sends to p
        recv_counts = counts[:, dist.get_rank(group)].tolist()
    width = rows.shape[1:]
    out = torch.empty((sum(recv_counts),) + tuple(width), dtype=rows.dtype, device=dev)
    per_row = 1
    for w in width:
        per_row *= w
    dist.all_to_all_single(out.reshape(-1), rows.contiguous().reshape(-1),
                           [c * per_row for c in recv_counts], [c * per_row for c in send_counts], group=group)
    assert len(recv_counts) == world
    return out, recv_counts


def _by_owner(owner: torch.Tensor, world: int, local_ops=None):
    """Stable order that groups items by owner rank + how many go to each rank.  On the GPU the order comes from the
    library's radix sort over the ceil(log2(world)) significant bits (one digit pass), and the counts from a binary
    search in the sorted owners (a histogram over `world` bins would be `n` atomics on a handful of addresses)."""
    if owner.is_cuda and hasattr(local_ops, "sort_pairs_u64") and owner.numel() > 0:
        keys = owner.clone()
        order = local_ops.sort_pairs_u64(keys, max(1, (world - 1).bit_length()))[0].long()
        grouped = keys
    else:
        grouped, order = torch.sort(owner, stable=True)
    bounds = torch.searchsorted(grouped, torch.arange(world + 1, device=owner.device, dtype=owner.dtype))
    return order, (bounds[1:] - bounds[:-1]).tolist()


def _take_rows(rows: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``rows[index]`` for a [N, k] matrix with a few int64 columns.  torch's row gather launches one CTA per 16-byte
    row (5 ms for 20M rows on B200); gathering column by column runs at memory speed."""
    if rows.dim() != 2 or not rows.is_cuda or rows.size(1) > 8:
        return rows[index]
    out = torch.empty((index.numel(), rows.size(1)), dtype=rows.dtype, device=rows.device)
    for c in range(rows.size(1)):
        out[:, c] = rows[:, c][index]
    return out


# ------------------------------------------------------------------------------------------------
# distributed temporal lift (BASELINE config 5)
# ------------------------------------------------------------------------------------------------
@dataclass
class DistributedLayer:
    """Rows [row_offset, row_offset + node_sequence.size(0)) of the global De Bruijn layer of one order.
    Concatenating ``node_sequence`` / ``edge_index`` / ``edge_weight`` over the ranks in rank order gives
    exactly the single-device layer."""

    order: int
    num_nodes: int               # global
    row_offset: int              # first global node id owned by this rank
    node_sequence: torch.Tensor  # [owned, k] first-order node ids
    edge_index: torch.Tensor     # [2, owned edges] GLOBAL node ids, (row, col)-sorted, rows in the owned range
    edge_weight: torch.Tensor

    def gather(self, group=None) -> "DistributedLayer":
        """The full layer on every rank (tests / small graphs)."""
        ns = _all_gather_var(self.node_sequence, group)
        ei = _all_gather_var(self.edge_index.t().contiguous(), group).t().contiguous()
        w = _all_gather_var(self.edge_weight, group)
        return DistributedLayer(self.order, self.num_nodes, 0, ns, ei, w)


def _all_gather_var(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    counts = _all_gather_int([t.size(0)], t.device, group)[:, 0].tolist()
    pad = max(counts) if counts else 0
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[:t.size(0)] = t
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def partition_stream(num_edges: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range of positions of the time-sorted stream owned by ``rank``."""
    return num_edges * rank // world, num_edges * (rank + 1) // world


def exchange_ghost_zone(edge_index: torch.Tensor, time: torch.Tensor, weight: torch.Tensor | None, horizon,
                        group=None):
    """Append to the local range the edges of the FOLLOWING ranks with ``t <= t_last(own) + horizon``.

    Every rank holds a contiguous range of the globally time-sorted stream, so what rank q needs is a
    prefix of the concatenation of the later ranks' ranges; rank r > q sends the prefix of its range with
    ``t <= t_last(q) + horizon`` (empty ranges in between contribute nothing).  Returns the extended
    (edge_index, time, weight) -- own edges first, then the ghosts in stream order."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = time.device
    m = time.numel()
    is_float = time.is_floating_point()
    # last own time stamp per rank (ranks with an empty range: "nothing needed")
    last = torch.zeros(world, dtype=torch.float64 if is_float else torch.int64, device=dev)
    has = torch.zeros(world, dtype=torch.int64, device=dev)
    if m > 0:
        last[rank] = time[-1].to(last.dtype)
        has[rank] = 1
    dist.all_reduce(last, group=group)
    dist.all_reduce(has, group=group)
    send_counts = [0] * world
    for q in range(rank):
        if has[q].item() and m > 0:
            limit = (last[q] + horizon).to(time.dtype) if not is_float else last[q] + horizon
            send_counts[q] = int(torch.searchsorted(time, limit.reshape(1).to(time.dtype), right=True))
    # rows are ordered by destination rank: prefix for rank 0, prefix for rank 1, ...
    packed = torch.cat([time.to(torch.float64).view(torch.int64) if is_float else time,
                        edge_index[0], edge_index[1]]).reshape(3, m).t().contiguous() if m else \
        torch.empty((0, 3), dtype=torch.int64, device=dev)
    cols = [packed]
    if weight is not None:
        cols.append(weight.to(torch.float64).view(torch.int64).unsqueeze(1))
    packed = torch.cat(cols, dim=1)
    to_send = torch.cat([packed[:send_counts[q]] for q in range(world)], dim=0)
    got, recv_counts = all_to_all_rows(to_send, send_counts, group)
    # a later rank's prefix is only usable if all ranks in between were sent COMPLETELY; by construction
    # (sorted stream, same limit) they were, unless they are empty.
    ghost_t = got[:, 0].view(torch.float64).to(time.dtype) if is_float else got[:, 0]
    ext_ei = torch.cat([edge_index, got[:, 1:3].t()], dim=1)
    ext_t = torch.cat([time, ghost_t])
    ext_w = torch.cat([weight, got[:, 3].view(torch.float64).to(weight.dtype)]) if weight is not None else None
    return ext_ei.contiguous(), ext_t.contiguous(), ext_w


def _owner_of_id(gid: torch.Tensor, offsets: torch.Tensor, world: int) -> torch.Tensor:
    """Rank whose owned id range contains ``gid`` (``offsets`` [world+1] cumulative sizes, host tensor)."""
    return (torch.searchsorted(offsets.to(gid.device), gid, right=True) - 1).clamp_(0, world - 1)


def _exchange_coalesce(gsrc, gdst, w, last, num_nodes, offsets, local_ops, group):
    """Send every edge (global node ids) to the rank that owns its source row and merge duplicates there.

    Returns, for the owner: the merged (row, col)-sorted edges, their summed weights and the ``last`` value of every
    merged edge (identical for all duplicates); for the sender: the GLOBAL index of the merged edge every input edge
    fell into -- owners hold ascending row ranges, so the concatenation of their merged lists is the global
    (row, col) order, and an index into it is the id of the next layer's De Bruijn node (the distinct edges of a
    layer, in (row, col) order, are the k-grams of the next one in lexicographic order); and the cumulative merged
    edge counts per rank [world + 1] (host)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = gsrc.device
    order, counts = _by_owner(_owner_of_id(gsrc, offsets, world), world, local_ops)
    narrow = w.dtype == torch.float32 and num_nodes < (1 << 31)
    if narrow:
        # ids and first-order nodes fit 32 bits: (source, target) and (weight bits, last node) travel as two words
        low = (1 << 32) - 1
        packed = torch.stack([(gsrc << 32) | gdst, (w.view(torch.int32).to(torch.int64) << 32) | (last & low)], dim=1)
        got, recv_counts = all_to_all_rows(_take_rows(packed, order), counts, group)
        ei = torch.stack([got[:, 0] >> 32, got[:, 0] & low])
        ww = (got[:, 1] >> 32).to(torch.int32).view(torch.float32)
        got_last = got[:, 1] & low
    else:
        payload = torch.stack([gsrc[order], gdst[order], w[order].to(torch.float64).view(torch.int64), last[order]], dim=1)
        got, recv_counts = all_to_all_rows(payload, counts, group)
        ei = got[:, :2].t().contiguous()
        ww = got[:, 2].view(torch.float64).to(w.dtype)
        got_last = got[:, 3]
    if ei.size(1):
        out_ei, out_w, inv = local_ops.coalesce(ei, None, num_nodes, ww, "sum", return_inverse=True)
    else:
        out_ei, out_w, inv = ei, ww, torch.empty(0, dtype=torch.int64, device=dev)
    out_last = torch.empty(out_ei.size(1), dtype=torch.int64, device=dev)
    out_last[inv] = got_last
    sizes = _all_gather_int([out_ei.size(1)], dev, group)[:, 0]
    edge_offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(sizes, 0)])
    # ids return to the senders along the same routes: what came from rank q goes back to q
    back, _ = all_to_all_rows((inv + int(edge_offsets[rank])).unsqueeze(1), recv_counts, group, recv_counts=counts)
    gid = torch.empty(gsrc.size(0), dtype=torch.int64, device=dev)
    gid[order] = back[:, 0]
    return out_ei, out_w, out_last, gid, edge_offsets


def distributed_temporal_layers(edge_index: torch.Tensor, time: torch.Tensor, num_nodes: int, delta, max_order: int,
                                edge_weight: torch.Tensor | None = None, group=None, local_ops=None) -> dict:
    """``MultiOrderModel.from_temporal_graph`` over a stream that is split across the ranks.

    ``edge_index`` [2, m_local] / ``time`` [m_local]: this rank's CONTIGUOUS range of the globally
    time-sorted stream (ranges in rank order).  Returns ``{order: DistributedLayer}``.

    Per order there is ONE exchange: the line-graph edges of the extended (own + ghost) range travel to the owner
    of their source row as (source id, target id, weight, last node) and are merged there; the owner returns the
    global index of the merged edge, which IS the node id of the next order (see ``_exchange_coalesce``), so no
    k-gram row ever has to be ranked or sent.  Edges whose path starts with a ghost event are sent with weight 0:
    they only collect their ids (their owner contributes the weight), every path is counted exactly once."""
    if local_ops is None:
        from . import ops as local_ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = edge_index.device
    m_own = edge_index.size(1)
    w_own = edge_weight if edge_weight is not None else torch.ones(m_own, dtype=torch.float32, device=dev)
    layers: dict[int, DistributedLayer] = {}

    # ---- ghost zone once: everything after it is local lifting + one exchange per order
    if max_order > 1:
        ext_ei, ext_t, ext_w = exchange_ghost_zone(edge_index, time, w_own, delta * (max_order - 1), group)
    else:
        ext_ei, ext_t, ext_w = edge_index, time, w_own
    counted = ext_w.clone()
    counted[m_own:] = 0                                  # ghost events: ids only

    # ---- order 1: nodes are the first-order nodes themselves, rows owned by node-id range
    n1_bounds = torch.tensor([-(-num_nodes * p // world) for p in range(world + 1)], dtype=torch.int64)
    lo, hi = int(n1_bounds[rank]), int(n1_bounds[rank + 1])
    ei_k, w_k, last_k, gid_line, offsets = _exchange_coalesce(ext_ei[0], ext_ei[1], counted, ext_ei[1], num_nodes, n1_bounds,
                                                              local_ops, group)
    rows_k = torch.arange(lo, hi, device=dev).unsqueeze(1)
    layers[1] = DistributedLayer(1, num_nodes, lo, rows_k, ei_k, w_k)
    if max_order == 1:
        return layers

    # ---- order 2: line-graph nodes are the (own + ghost) events, their ids the merged first-order edges
    try:
        line_index = local_ops.lift_order_temporal(ext_ei, ext_t, delta, num_nodes)
    except (RuntimeError, ValueError):                # no pair in this range (the single-device call fails only if NO rank has one)
        line_index = torch.empty((2, 0), dtype=torch.int64, device=dev)
    line_w = local_ops.pair_attributes(line_index, counted, "src") if line_index.size(1) else counted[:0]
    last_line = ext_ei[1]                             # last first-order node of every line-graph node's k-gram
    num_line_nodes = ext_ei.size(1)
    row_lo = lo

    for k in range(2, max_order + 1):
        # rows of the nodes this rank owns at order k = its merged edges of order k - 1
        prev_rows, prev_lo = rows_k, row_lo
        rows_k = torch.cat([_take_rows(prev_rows, ei_k[0] - prev_lo), last_k.unsqueeze(1)], dim=1)
        total, row_lo = int(offsets[-1]), int(offsets[rank])
        edge_last = last_line[line_index[1]]
        ei_k, w_k, last_k, gid_next, next_offsets = _exchange_coalesce(gid_line[line_index[0]], gid_line[line_index[1]], line_w,
                                                                       edge_last, total, offsets, local_ops, group)
        layers[k] = DistributedLayer(k, total, row_lo, rows_k, ei_k, w_k)
        if k == max_order:
            break
        nxt = local_ops.lift_order_edge_index(line_index, num_line_nodes)
        line_w = local_ops.pair_attributes(nxt, line_w, "src") if nxt.size(1) else line_w[:0]
        gid_line, last_line, offsets = gid_next, edge_last, next_offsets
        num_line_nodes, line_index = line_index.size(1), nxt
    return layers
