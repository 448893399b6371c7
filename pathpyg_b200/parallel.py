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
  locally and keeps the causal paths whose FIRST edge it owns (every path is counted exactly once);
  De Bruijn node ids are made global by range-partitioning the k-gram rows on their first node
  (all-to-all-v -> local radix sort + unique -> all-gather of counts -> exclusive scan = lexicographic rank)
  and the mapped edges are exchanged to the owner of their source row and coalesced there.

The local compute goes through ``pathpyg_b200.ops`` (CUDA only).  The CPU tests inject an object with the
same five functions backed by the oracle, so that the partition / exchange / id-assignment logic is
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


def all_to_all_rows(rows: torch.Tensor, send_counts: list[int], group=None) -> tuple[torch.Tensor, list[int]]:
    """Variable all-to-all of the leading dimension: ``rows`` is ordered by destination rank,
    ``send_counts[p]`` rows go to rank p.  Returns (received rows ordered by source rank, recv_counts)."""
    world = dist.get_world_size(group)
    dev = rows.device
    counts = _all_gather_int(send_counts, dev, group)          # counts[q, p] = rows q sends to p
    rank = dist.get_rank(group)
    recv_counts = counts[:, rank].tolist()
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


class _GlobalIds:
    """Global lexicographic ranks of k-gram rows, range-partitioned on the first node of the row."""

    def __init__(self, num_first_order_nodes: int, local_ops, group):
        self.n1, self.ops, self.group = num_first_order_nodes, local_ops, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def owner_of(self, first_node: torch.Tensor) -> torch.Tensor:
        return (first_node * self.world) // max(self.n1, 1)

    def resolve(self, rows: torch.Tensor):
        """``rows`` [r, k] (any order, duplicates allowed).  Returns (global id of every row,
        owned distinct rows sorted, first global id owned, global number of distinct rows)."""
        dev = rows.device
        k = rows.size(1)
        uniq, inv = self.ops.unique_rows(rows) if rows.size(0) else (rows, torch.empty(0, dtype=torch.int64, device=dev))
        order, counts = _by_owner(self.owner_of(uniq[:, 0]).clamp_(max=self.world - 1), self.world, self.ops)
        got, recv_counts = all_to_all_rows(_take_rows(uniq, order), counts, self.group)
        if got.size(0):
            owned, got_inv = self.ops.unique_rows(got)
        else:
            owned, got_inv = got.reshape(0, k), torch.empty(0, dtype=torch.int64, device=dev)
        sizes = _all_gather_int([owned.size(0)], dev, self.group)[:, 0]
        offset = int(sizes[:self.rank].sum())
        total = int(sizes.sum())
        back, _ = all_to_all_rows((got_inv + offset).unsqueeze(1), recv_counts, self.group)   # ids return to the askers
        gid_sorted_by_owner = back[:, 0]
        gid_of_uniq = torch.empty_like(gid_sorted_by_owner)
        gid_of_uniq[order] = gid_sorted_by_owner
        return gid_of_uniq[inv], owned, offset, total

    def owner_of_id(self, gid: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        """Rank whose owned id range contains ``gid`` (``offsets`` [world+1] cumulative sizes)."""
        return (torch.searchsorted(offsets.to(gid.device), gid, right=True) - 1).clamp_(0, self.world - 1)


def _coalesce_at_owner(gsrc, gdst, w, num_nodes, offsets, ids: _GlobalIds, local_ops):
    """Send every mapped edge to the rank that owns its source row; merge duplicates there."""
    dev = gsrc.device
    order, counts = _by_owner(ids.owner_of_id(gsrc, offsets), ids.world, local_ops)
    payload = torch.stack([gsrc[order], gdst[order], w[order].to(torch.float64).view(torch.int64)], dim=1)
    got, _ = all_to_all_rows(payload, counts, ids.group)
    ei = got[:, :2].t().contiguous()
    ww = got[:, 2].view(torch.float64).to(w.dtype)
    if ei.size(1) == 0:
        return ei, ww
    return local_ops.coalesce(ei, None, num_nodes, ww, "sum")


def distributed_temporal_layers(edge_index: torch.Tensor, time: torch.Tensor, num_nodes: int, delta, max_order: int,
                                edge_weight: torch.Tensor | None = None, group=None, local_ops=None) -> dict:
    """``MultiOrderModel.from_temporal_graph`` over a stream that is split across the ranks.

    ``edge_index`` [2, m_local] / ``time`` [m_local]: this rank's CONTIGUOUS range of the globally
    time-sorted stream (ranges in rank order).  Returns ``{order: DistributedLayer}``."""
    if local_ops is None:
        from . import ops as local_ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = edge_index.device
    ids = _GlobalIds(num_nodes, local_ops, group)
    m_own = edge_index.size(1)
    w_own = edge_weight if edge_weight is not None else torch.ones(m_own, dtype=torch.float32, device=dev)
    layers: dict[int, DistributedLayer] = {}

    # ---- order 1: nodes are the first-order nodes themselves, rows owned by node-id range
    n1_bounds = torch.tensor([-(-num_nodes * p // world) for p in range(world + 1)], dtype=torch.int64)  # ceil: matches owner_of
    lo, hi = int(n1_bounds[rank]), int(n1_bounds[rank + 1])
    ei1, w1 = _coalesce_at_owner(edge_index[0], edge_index[1], w_own, num_nodes, n1_bounds, ids, local_ops)
    layers[1] = DistributedLayer(1, num_nodes, lo, torch.arange(lo, hi, device=dev).unsqueeze(1), ei1, w1)
    if max_order == 1:
        return layers

    # ---- ghost zone once, then purely local lifts on the extended range
    horizon = delta * (max_order - 1)
    ext_ei, ext_t, ext_w = exchange_ghost_zone(edge_index, time, w_own, horizon, group)
    node_sequence = ext_ei.t().contiguous()           # 2-gram of every (own + ghost) temporal edge
    own_count = m_own                                 # line-graph nodes whose path starts with an own edge: a prefix
    try:
        line_index = local_ops.lift_order_temporal(ext_ei, ext_t, delta, num_nodes)
    except (RuntimeError, ValueError):                # no pair in this range (the single-device call fails only if NO rank has one)
        line_index = torch.empty((2, 0), dtype=torch.int64, device=dev)
    line_w = local_ops.pair_attributes(line_index, ext_w, "src") if line_index.size(1) else ext_w[:0]
    num_line_nodes = ext_ei.size(1)

    for k in range(2, max_order + 1):
        # own line-graph edges of this level: columns whose source is an own line-graph node (a prefix, sources ascend)
        own_edges = int(torch.searchsorted(line_index[0].contiguous(), torch.tensor([own_count], device=dev), right=False)) \
            if line_index.size(1) else 0
        dst_rows = _take_rows(node_sequence, line_index[1, :own_edges])
        cand = torch.cat([node_sequence[:own_count], dst_rows], dim=0)    # node candidates + look-ups (all are real nodes)
        gid, owned_rows, offset, total = ids.resolve(cand)
        sizes = _all_gather_int([owned_rows.size(0)], dev, group)[:, 0]
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(sizes, 0)])
        # source k-gram of an own edge = node_sequence row of its source, which is among the first own_count candidates
        gsrc = gid[line_index[0, :own_edges]]
        gdst = gid[own_count:]
        ei_k, w_k = _coalesce_at_owner(gsrc, gdst, line_w[:own_edges], total, offsets, ids, local_ops)
        layers[k] = DistributedLayer(k, total, offset, owned_rows, ei_k, w_k)
        if k == max_order:
            break
        # next level on the extended range
        nxt = local_ops.lift_order_edge_index(line_index, num_line_nodes)
        line_w = local_ops.pair_attributes(nxt, line_w, "src") if nxt.size(1) else line_w[:0]
        if hasattr(local_ops, "extend_rows") and node_sequence.is_cuda:
            node_sequence = local_ops.extend_rows(node_sequence, line_index)
        else:
            node_sequence = torch.cat([node_sequence[line_index[0]], node_sequence[line_index[1]][:, -1:]], dim=1)
        num_line_nodes, own_count, line_index = line_index.size(1), own_edges, nxt
    return layers
