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
  locally -- only the prefix of every line graph that later orders still need -- and counts the causal paths
  whose FIRST edge it owns (every path is counted exactly once); per order ONE all-to-all-v carries the
  line-graph edges as 16-byte records to the owner of their source row (stable partition by owner, ``csrc/
  exchange.cu``), where they are merged (radix sort + run detection); the global index of a merged edge --
  all-gather of the merged counts + exclusive scan -- returns to the senders as 4 bytes per edge and IS the
  De Bruijn node id of the next order.  The local lift of the next order runs while the records are in flight.

The local compute goes through ``pathpyg_b200.ops`` (CUDA only).  The CPU tests inject an object with the
same functions written in torch, so that the partition / exchange / id-assignment / pruning logic is
exercised with gloo at world size 2 and 3 without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _lib


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
    edge_offset: int = 0         # global index of the first owned edge (= id of the next order's first owned node)
    num_edges: int = 0           # global

    def digest(self) -> torch.Tensor:
        """This rank's part of ``layer_digest`` (sum over the ranks = digest of the whole layer)."""
        return layer_digest(self.edge_index, self.edge_weight, self.node_sequence, self.edge_offset, self.row_offset)

    def gather(self, group=None) -> "DistributedLayer":
        """The full layer on every rank (tests / small graphs)."""
        ns = _all_gather_var(self.node_sequence, group)
        ei = _all_gather_var(self.edge_index.t().contiguous(), group).t().contiguous()
        w = _all_gather_var(self.edge_weight, group)
        return DistributedLayer(self.order, self.num_nodes, 0, ns, ei, w, 0, self.num_edges)


# ------------------------------------------------------------------------------------------------
# order-sensitive 64-bit digests: a distributed build is checked against a single-device one without gathering it
# ------------------------------------------------------------------------------------------------
_DIGEST_CHUNK = 1 << 25
_DIGEST_MULS = (0x2545F4914F6CDD1D, -0x61C8864680B583EB, 0x1B873593CC9E2D51, -0x3A39CE76F1A9D8B5, 0x27D4EB2F165667C5,
                -0x00B1A2C3D4E5F607, 0x5851F42D4C957F2D)


def _mix_sum(first_position: int, columns) -> torch.Tensor:
    """sum_i mix(first_position + i, columns[0][i], columns[1][i], ...) in wrapping int64 arithmetic: depends on the
    position of every element, and partial sums over disjoint position ranges add up to the sum over their union."""
    n, dev = columns[0].numel(), columns[0].device
    total = torch.zeros((), dtype=torch.int64, device=dev)
    for a in range(0, n, _DIGEST_CHUNK):
        b = min(n, a + _DIGEST_CHUNK)
        h = (torch.arange(a, b, device=dev, dtype=torch.int64) + (first_position + 1)) * _DIGEST_MULS[0]
        for j, col in enumerate(columns):
            h = (h ^ col[a:b]) * _DIGEST_MULS[1 + j % (len(_DIGEST_MULS) - 1)]
            h ^= h >> 29
        total += h.sum()
    return total


def layer_digest(edge_index, edge_weight, node_sequence, edge_offset: int = 0, row_offset: int = 0) -> torch.Tensor:
    """[2] int64 (edges, nodes) of the rows / edges held here, at their GLOBAL positions.  The digests of the parts of a
    ``DistributedLayer`` summed over the ranks (wrapping) equal the digest of the single-device layer iff -- up to
    64-bit hash collisions -- edge index, weights and node sequences agree element by element."""
    ei = edge_index.as_subclass(torch.Tensor)
    w = edge_weight.contiguous()
    wbits = w.view(torch.int32).long() if w.dtype == torch.float32 else w.to(torch.float64).view(torch.int64)
    edges = _mix_sum(edge_offset, [ei[0], ei[1], wbits])
    ns = node_sequence.as_subclass(torch.Tensor)
    nodes = _mix_sum(row_offset, [ns[:, c] for c in range(ns.size(1))])
    return torch.stack([edges, nodes])


def _all_gather_var(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    counts = _all_gather_int([t.size(0)], t.device, group)[:, 0].tolist()
    pad = max(counts) if counts else 0
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[:t.size(0)] = t
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def partition_stream(num_edges: int, rank: int, world: int, last_share: float = 1.0) -> tuple[int, int]:
    """Contiguous range of positions of the time-sorted stream owned by ``rank``.  ``last_share``: size of the LAST
    rank's range relative to the others -- it has no ghost zone (nothing follows it), so it can own ``1 + g`` times as
    many events as a rank whose ghost paths cost it a fraction ``g`` of extra work."""
    if world == 1:
        return 0, num_edges
    unit = num_edges / (world - 1 + last_share)
    cuts = [min(num_edges, int(round(unit * r))) for r in range(world)] + [num_edges]
    return cuts[rank], cuts[rank + 1]


def exchange_ghost_zone(edge_index: torch.Tensor, time: torch.Tensor, weight: torch.Tensor | None, horizon,
                        group=None):
    """Append to the local range the edges of the FOLLOWING ranks with ``t <= t_last(own) + horizon``.

    Every rank holds a contiguous range of the globally time-sorted stream, so what rank q needs is a
    prefix of the concatenation of the later ranks' ranges; rank r > q sends the prefix of its range with
    ``t <= t_last(q) + horizon`` (empty ranges in between contribute nothing).  Returns the extended
    (edge_index, time, weight) -- own edges first, then the ghosts in stream order.

    Two host synchronisations: the ranks' last time stamps, then the matrix of prefix lengths."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = time.device
    m = time.numel()
    is_float = time.is_floating_point()
    # last own time stamp per rank (ranks with an empty range need nothing), as raw 64-bit words
    mine = torch.zeros(2, dtype=torch.int64, device=dev)
    if m > 0:
        mine[0] = time[-1].to(torch.float64).view(torch.int64) if is_float else time[-1].to(torch.int64)
        mine[1] = 1
    lasts = _gather_counts(mine, group)
    last = lasts[:, 0].view(torch.float64) if is_float else lasts[:, 0]
    has = lasts[:, 1].tolist()
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    earlier = [q for q in range(rank) if has[q] and m > 0]
    if earlier:
        limits = (last[earlier] + horizon).to(time.dtype).to(dev)
        counts[earlier] = torch.searchsorted(time, limits, right=True)
    matrix = _gather_counts(counts, group)                      # matrix[r, q] = events r sends to q
    send_counts, recv_counts = matrix[rank].tolist(), matrix[:, rank].tolist()

    # one transfer: per destination the prefix of every column (source row, target row, time, weight) as 64-bit words
    columns = [edge_index[0], edge_index[1], time.view(torch.int64) if time.dtype in (torch.int64, torch.float64) else time.to(torch.int64)]
    if weight is not None:
        columns.append(weight.to(torch.float64).view(torch.int64))
    ncol = len(columns)
    parts = [col[:c] for c in send_counts if c for col in columns]
    to_send = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device=dev)
    got = torch.empty(ncol * sum(recv_counts), dtype=torch.int64, device=dev)
    dist.all_to_all_single(got, to_send, [ncol * c for c in recv_counts], [ncol * c for c in send_counts], group=group)
    # a later rank's prefix is only usable if all ranks in between were sent COMPLETELY; by construction
    # (sorted stream, same limit) they were, unless they are empty.
    blocks, at = [], 0
    for c in recv_counts:
        if c:
            blocks.append(got[at:at + ncol * c].view(ncol, c))
            at += ncol * c
    ghosts = torch.cat(blocks, dim=1) if blocks else torch.empty((ncol, 0), dtype=torch.int64, device=dev)
    ext_ei = torch.cat([edge_index, ghosts[:2]], dim=1)
    ghost_t = ghosts[2].view(time.dtype) if time.dtype in (torch.int64, torch.float64) else ghosts[2].to(time.dtype)
    ext_t = torch.cat([time, ghost_t])
    ext_w = torch.cat([weight, ghosts[3].view(torch.float64).to(weight.dtype)]) if weight is not None else None
    return ext_ei.contiguous(), ext_t.contiguous(), ext_w


class _Trace:
    """NVTX range per phase (always; free without a profiler) and, with ``PPG_DIST_TRACE=1``, CUDA-event time stamps
    of the phases of one call: ``parallel.last_trace`` then holds [(label, ms since the start of the call)]."""

    def __init__(self, dev):
        import os
        self.timed = os.environ.get("PPG_DIST_TRACE", "0") == "1" and dev.type == "cuda"
        self.cuda = dev.type == "cuda"
        self.marks = []
        self.sizes = []      # (level, sources, sources expanded, pairs) of every expansion, when timed
        self.open = False
        self.mark("start")

    def mark(self, label):
        if self.cuda:
            if self.open:
                torch.cuda.nvtx.range_pop()
            torch.cuda.nvtx.range_push(label)
            self.open = True
        if self.timed:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((label, ev))

    def close(self):
        global last_trace
        if self.cuda and self.open:
            torch.cuda.nvtx.range_pop()
            self.open = False
        if self.timed:
            torch.cuda.synchronize()
            t0 = self.marks[0][1]
            last_trace = [(label, t0.elapsed_time(ev)) for label, ev in self.marks]
            global last_sizes
            last_sizes = list(self.sizes)


last_trace = None
last_sizes = None


def _gather_counts(mine: torch.Tensor, group) -> torch.Tensor:
    """All ranks' copies of a small device vector -> [world, len] on the host (the one synchronisation of a step)."""
    world = dist.get_world_size(group)
    mine = mine.reshape(1, -1).contiguous()
    out = torch.empty((world, mine.size(1)), dtype=mine.dtype, device=mine.device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), mine[0], group=group)
    return out.cpu()


class _PeerArenas:
    """Two symmetric receive buffers per (process, device), mapped into every peer (``torch.distributed.
    _symmetric_memory``: CUDA virtual-memory handles exchanged once, NVLink peer access).  The pack kernel of a sender
    stores its records straight into the owner's buffer; a device-side barrier over the buffers' signal pads tells the
    owner that all senders are done.  Levels alternate between the two buffers: the barrier of level k + 1 is behind
    every rank's reads of level k's records in stream order, so level k + 2 may overwrite them."""

    _cache: dict = {}
    disabled = None   # None: not tried yet

    def __init__(self, group, dev):
        self.group, self.dev, self.capacity = group, dev, 0
        self.buffers, self.handles = [None, None], [None, None]

    @classmethod
    def get(cls, group, dev):
        key = (id(group) if group is not None else 0, dev.index)
        if key not in cls._cache:
            cls._cache[key] = cls(group, dev)
        return cls._cache[key]

    @classmethod
    def available(cls, dev) -> bool:
        import os
        if cls.disabled is None:
            cls.disabled = os.environ.get("PPG_DIST_P2P", "1") == "0" or dev.type != "cuda"
            if not cls.disabled:
                try:
                    import torch.distributed._symmetric_memory  # noqa: F401
                except Exception:
                    cls.disabled = True
        return not cls.disabled

    def ensure(self, records: int) -> None:
        """Collective when it grows: every rank passes the same number (the largest receive count of the level)."""
        if records <= self.capacity:
            return
        import torch.distributed._symmetric_memory as symm_mem
        capacity = (int(records * 1.02) + 4096) // 1024 * 1024
        group = self.group if self.group is not None else dist.group.WORLD
        for i in range(2):
            self.handles[i] = self.buffers[i] = None                # release the old mapping first
        for i in range(2):
            self.buffers[i] = symm_mem.empty(capacity * 2, dtype=torch.int64, device=self.dev)
            self.handles[i] = symm_mem.rendezvous(self.buffers[i], group)
        self.capacity = capacity

    def peer_slot(self, which: int, rank: int, record_offset: int) -> int:
        return int(self.handles[which].buffer_ptrs[rank]) + 16 * int(record_offset)

    def barrier(self, which: int) -> None:
        self.handles[which].barrier(channel=0)

    def received(self, which: int, count: int) -> torch.Tensor:
        return self.buffers[which][:2 * count].view(count, 2)


class _PendingIds:
    """The merged-edge ids of one level on their way back to the senders."""

    def __init__(self, plan, back, work, edge_offsets_dev, mark, level):
        self.plan, self.back, self.work, self.edge_offsets_dev, self.mark, self.level = plan, back, work, edge_offsets_dev, mark, level

    def finish(self) -> torch.Tensor:
        """node_info of the next level: one word per input edge, merged id << 32 | last node."""
        self.mark(f"ids_wait[{self.level}]")
        self.work.wait()
        self.mark(f"route_unpack[{self.level}]")
        return self.plan.unpack(self.back, self.edge_offsets_dev)


def _exchange_merge(line_index, node_info, weights, own_prefix, offsets, offsets_dev, total_nodes, local_ops, group,
                    lift_begin=None, trace=None, level=0):
    """One level of the exchange.  Every line-graph edge travels as a 16-byte record to the rank that owns its source
    row (``offsets``: first row of every rank), is merged there with its duplicates, and the owner returns the index of
    the merged edge.  Owners hold ascending row ranges, so the concatenation of their merged lists is the global
    (row, col) order and an index into it is the id of the next layer's De Bruijn node (the distinct edges of a layer in
    (row, col) order are the k-grams of the next one in lexicographic order).

    ``lift_begin``: callable that enqueues the count pass of the NEXT level's local lift; its column count is collected
    in the same synchronisation as the record counts, and its fill pass runs while the records are in flight.

    Returns (merged edge_index (global ids), weights, last nodes) of the owned rows; a ``_PendingIds`` whose
    ``finish()`` gives the ``node_info`` of the next level; the cumulative merged counts [world + 1] as a host list and
    as a device tensor; the next level's line graph (or None).  Two host synchronisations."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = line_index.device
    mark = trace.mark if trace is not None else (lambda label: None)
    mark(f"route_count[{level}]")
    plan = local_ops.route_plan(line_index, node_info, offsets_dev, world)
    pending = lift_begin() if lift_begin is not None else None
    payload = plan.counts if pending is None else torch.cat([plan.counts, pending.result_words])
    gathered = _gather_counts(payload, group)                        # [q, p] = records q sends to p (+ lift count)  (sync 1)
    matrix = gathered[:, :world]
    send, recv = matrix[rank].tolist(), matrix[:, rank].tolist()
    mark(f"route_pack[{level}]")
    peer_to_peer = world > 1 and _PeerArenas.available(dev)
    if peer_to_peer:
        # fused partition + transfer: the pack kernel stores every record into its owner's receive buffer over NVLink
        arenas = _PeerArenas.get(group, dev)
        arenas.ensure(int(matrix.sum(0).max()))
        which = level & 1
        ahead = matrix[:rank].sum(0).tolist()                        # records of the lower ranks in every owner's buffer
        plan.pack(weights, own_prefix, peer_slots=[arenas.peer_slot(which, d, ahead[d]) for d in range(world)])
    else:
        records = plan.pack(weights, own_prefix)
        received = torch.empty((sum(recv), 2), dtype=torch.int64, device=dev)
        work = dist.all_to_all_single(received.view(-1), records.view(-1), [2 * c for c in recv], [2 * c for c in send],
                                      group=group, async_op=True)
    mark(f"lift_next[{level}]")
    carried = None
    if pending is not None:                                          # overlaps the transfer / absorbs the ranks' skew
        carried = pending.finish(total=int(gathered[rank, world]), status=int(gathered[rank, world + 1]), allow_empty=True)
    mark(f"records_wait[{level}]")
    if peer_to_peer:
        arenas.barrier(which)                                        # every sender's records have landed
        received = arenas.received(which, sum(recv))
    else:
        work.wait()
        del records
    mark(f"merge_sort[{level}]")
    row_lo, rows_owned = int(offsets[rank]), int(offsets[rank + 1] - offsets[rank])
    merge = local_ops.merge_records_begin(received, row_lo, rows_owned, total_nodes)
    mark(f"merge_sync[{level}]")
    # (collectives of one communicator run in issue order: the small all-gather goes first, the bulk transfer after it)
    results = _gather_counts(merge.result_words, group)              # [world, 2]: merged count, status   (sync 2)
    # the merged-edge indices go back along the same routes while the owner writes its merged edges
    back = torch.empty(plan.E, dtype=torch.int32, device=dev)
    work = dist.all_to_all_single(back, merge.inverse, send, recv, group=group, async_op=True)
    if int(results[:, 1].max()) & 1:
        raise ValueError("distributed lift: a node id outside its layer reached an owner (inconsistent inputs)")
    sizes = results[:, 0]
    edge_offsets = [0]
    for c in sizes.tolist():
        edge_offsets.append(edge_offsets[-1] + int(c))
    if edge_offsets[-1] > (1 << 32):
        raise ValueError(f"distributed lift: {edge_offsets[-1]} merged edges exceed the 32-bit record fields")
    edge_offsets_dev = torch.tensor(edge_offsets, dtype=torch.int64, device=dev)
    mark(f"merge_fill[{level}]")
    out_ei, out_w, out_last = merge.finish(int(sizes[rank]))
    ids = _PendingIds(plan, back, work, edge_offsets_dev, mark, level)
    return out_ei, out_w, out_last, ids, edge_offsets, edge_offsets_dev, carried


# ------------------------------------------------------------------------------------------------
# distributed lift on the generation-order chain (csrc/chain.cu)
# ------------------------------------------------------------------------------------------------
class _ChainBuffers:
    """Record and answer buffers of the distributed chain on one rank: two record buffers (levels alternate: the tiles of
    level k + 1 read their weights from level k's records while they write their own) and one answer buffer.  With peer
    access (``torch.distributed._symmetric_memory``: CUDA virtual-memory handles exchanged once, NVLink) the buffers are
    symmetric allocations mapped into every peer: an owner's merge kernel loads the senders' records and stores the merged
    indices straight through those mappings, and a device-side barrier over the signal pads orders the phases."""

    _cache: dict = {}

    def __init__(self, group, dev, peer: bool):
        self.group, self.dev, self.peer, self.capacity = group, dev, peer, 0
        self.rec, self.back, self.rec_handles, self.back_handle = [None, None], None, [None, None], None

    @classmethod
    def get(cls, group, dev, peer: bool):
        key = (id(group) if group is not None else 0, dev.index, peer)
        if key not in cls._cache:
            cls._cache[key] = cls(group, dev, peer)
        return cls._cache[key]

    def ensure(self, slots: int, keep: int | None = None, keep_slots: int = 0) -> None:
        """Collective when it grows (every rank passes the same number).  ``keep``: record buffer whose first
        ``keep_slots`` records must survive the growth."""
        if slots <= self.capacity:
            return
        capacity = (int(slots * 1.05) + 4096) // 1024 * 1024
        saved = self.rec[keep][:2 * keep_slots].clone() if keep is not None and keep_slots else None
        torch.cuda.synchronize(self.dev)   # kernels in flight may still use the buffers that are about to be released
        if self.peer:
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            self.rec_handles, self.back_handle, self.rec, self.back = [None, None], None, [None, None], None   # release first
            for i in range(2):
                self.rec[i] = symm_mem.empty(capacity * 2, dtype=torch.int64, device=self.dev)
                self.rec_handles[i] = symm_mem.rendezvous(self.rec[i], group)
            self.back = symm_mem.empty(capacity, dtype=torch.int32, device=self.dev)
            self.back_handle = symm_mem.rendezvous(self.back, group)
        else:
            self.rec = [torch.empty(capacity * 2, dtype=torch.int64, device=self.dev) for _ in range(2)]
            self.back = torch.empty(capacity, dtype=torch.int32, device=self.dev)
        if saved is not None:
            self.rec[keep][:saved.numel()] = saved
        self.capacity = capacity

    def rec_ptr(self, which: int, rank: int, slot: int) -> int:
        base = int(self.rec_handles[which].buffer_ptrs[rank]) if self.peer else self.rec[which].data_ptr()
        return base + 16 * int(slot)

    def back_ptr(self, rank: int, slot: int) -> int:
        base = int(self.back_handle.buffer_ptrs[rank]) if self.peer else self.back.data_ptr()
        return base + 4 * int(slot)

    def peer_records(self, which: int, rank: int, lo: int, n: int) -> torch.Tensor:
        """[n, 2] int64 view of records lo .. lo + n of ``rank``'s buffer (the overflow fallback copies them)."""
        if not self.peer:
            return self.rec[which][2 * lo:2 * (lo + n)].view(n, 2)
        return self.rec_handles[which].get_buffer(rank, (self.capacity * 2,), torch.int64)[2 * lo:2 * (lo + n)].view(n, 2)

    def peer_back(self, rank: int, lo: int, n: int) -> torch.Tensor:
        if not self.peer:
            return self.back[lo:lo + n]
        return self.back_handle.get_buffer(rank, (self.capacity,), torch.int32)[lo:lo + n]

    def barrier(self) -> None:
        if self.peer:
            self.back_handle.barrier(channel=0)


class _MergeSorted:
    """Owner side on the chain: one (row, col)-sorted run of records per sender, merged in shared-memory tiles of whole
    row ranges (``ppg_merge_sorted``); the merged-edge index of every record is stored where ``back_ptrs`` say (the
    senders' answer buffers).  ``result_words`` [merged count, status]; ``finish(num_out)`` as ``ops.PendingMerge``."""

    def __init__(self, run_ptrs: list[int], run_lens: list[int], back_ptrs: list[int], row_lo: int, rows_owned: int, total_nodes: int,
                 dev):
        import ctypes
        from .ops import _ptr, _stream
        lib = _lib.load()
        world = len(run_lens)
        self.dev, self.R = dev, int(sum(run_lens))
        n = max(self.R, 1)
        tiles = int(lib.ppg_merge_sorted_tiles(self.R))
        self.compact = [torch.empty(n, dtype=torch.float32 if i == 2 else torch.int32, device=dev) for i in range(4)]
        self.result_words = torch.zeros(2, dtype=torch.int64, device=dev)
        self.keep = (torch.empty((tiles + 1) * world + 1, dtype=torch.int32, device=dev), torch.empty(tiles + 1, dtype=torch.int64, device=dev))
        row_m, col_m, w_m, last_m = self.compact
        runs = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p)) for p in run_ptrs])
        backs = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p)) for p in back_ptrs])
        lens = (ctypes.c_int64 * world)(*[int(c) for c in run_lens])
        with torch.cuda.device(dev):
            _lib.check(lib.ppg_merge_sorted(runs, lens, backs, world, int(row_lo), int(rows_owned), int(total_nodes), _ptr(self.keep[0]),
                                            _ptr(self.keep[1]), _ptr(row_m), _ptr(col_m), _ptr(w_m), _ptr(last_m), _ptr(self.result_words),
                                            _stream(dev)))

    def finish(self, num_out: int):
        from .ops import _ptr, _stream
        out_ei = torch.empty((2, num_out), dtype=torch.int64, device=self.dev)
        out_w = torch.empty(num_out, dtype=torch.float32, device=self.dev)
        out_last = torch.empty(num_out, dtype=torch.int64, device=self.dev)
        row_m, col_m, w_m, last_m = self.compact
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().ppg_merge_sorted_fill(_ptr(row_m), _ptr(col_m), _ptr(w_m), _ptr(last_m), int(num_out), _ptr(out_ei),
                                                         _ptr(out_w), _ptr(out_last), _stream(self.dev)))
        return out_ei, out_w, out_last


def _chain_supported(edge_index, edge_weight, num_nodes, local_ops) -> bool:
    import os
    return (local_ops is None and edge_index.is_cuda and os.environ.get("PPG_CHAIN", "1") != "0" and edge_index.dtype == torch.int64
            and (edge_weight is None or edge_weight.dtype == torch.float32) and num_nodes < (1 << 31))


def _distributed_chain_layers(ext_ei, ext_t, ext_w, m_own, num_nodes, delta, K, cuts, group, trace) -> dict:
    """``distributed_temporal_layers`` on the generation-order chain.  Every rank expands its items in the order of their
    GLOBAL merged ids (known from the previous level's exchange), so its pairs leave as 16-byte records sorted by
    (row, col) and grouped by owner -- no partition pass, no pack pass.  With peer access the owner's merge kernel READS
    the senders' runs over NVLink, merges them in shared-memory tiles (``ppg_merge_sorted``) and WRITES the merged-edge
    index of every record into the sender's answer buffer: the exchange in both directions is the merge kernel's own load
    / store stream, ordered by two device-side barriers per order.  Without peer access (``PPG_DIST_P2P=0``, or one
    rank) the same kernels run on copies moved by two all-to-all-v.

    Per order two host synchronisations: (1) all-gather of the per-destination record counts, the first slot of every
    destination, the heavy-row counts and the sizes of the next level; (2) all-gather of the merged counts.

    Paths that start with a ghost event carry weight 0 from level 1 on (a path inherits the weight of its first event),
    so they only collect ids.  At level k only the items that start before cut k are expanded (``limit``); how many items
    start before the later cuts is read off the row pointer of the level (``starts_before``)."""
    import ctypes
    from . import chain as chain_mod
    from . import ops
    from .ops import _ptr, _stream

    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = ext_ei.device
    m_ext = ext_ei.size(1)
    heavy, tile = chain_mod.heavy_threshold(), lib.ppg_chain_tile_slots()
    u32 = torch.int32
    mark = trace.mark
    res = torch.zeros((K + 2, 8), dtype=torch.int64, device=dev)
    peer = world > 1 and _PeerArenas.available(dev)
    bufs = _ChainBuffers.get(group, dev, peer)
    layers: dict[int, DistributedLayer] = {}

    def empty32(n, dtype=u32):
        return torch.empty(max(int(n), 1), dtype=dtype, device=dev)

    def scan_ws(n):
        return torch.empty(lib.ppg_chain_scan_workspace_bytes(int(n)), dtype=torch.uint8, device=dev)

    def word(t: torch.Tensor, index: int) -> ctypes.c_void_p:
        return ctypes.c_void_p(t.data_ptr() + 4 * index)

    with torch.cuda.device(dev):
        stream = _stream(dev)
        main, side = torch.cuda.current_stream(dev), chain_mod.side_stream(dev)
        # the buffers hold the largest level of any rank; level 1 = the events of the extended range
        most = int(_gather_counts(torch.tensor([m_ext], dtype=torch.int64, device=dev), group).max())
        bufs.ensure(max(most, 1))

        # ---- level 1: events grouped by source node, ranked by target node; ghost events carry weight 0
        w_event = torch.ones(m_ext, dtype=torch.float32, device=dev) if ext_w is None else ext_w.to(torch.float32).clone()
        w_event[m_own:] = 0.0
        n_slots = m_ext
        arrays = {name: empty32(n_slots, torch.float32 if name == "w" else u32) for name in ("row", "col", "lab", "w")}
        heavy_list = torch.empty((n_slots // (heavy + 1) + 2, 2), dtype=u32, device=dev)
        labS, firstS, degS = arrays["lab"], None, None
        info = torch.empty((max(n_slots, 1) + 1, 2), dtype=torch.int64, device=dev)    # by label: {id, last, count -> pointer, -}
        tws = views = None
        if m_ext:
            tws = ops.lift_order_temporal_group(ext_ei, num_nodes)
            views = (ctypes.c_void_p * 6)()
            _lib.check(lib.ppg_lift_temporal_views(_ptr(tws), m_ext, num_nodes, views))
            state = torch.empty(-(-m_ext // tile) + 1, dtype=torch.int64, device=dev)
            _lib.check(lib.ppg_chain_first_tiles(_ptr(ext_ei), m_ext, num_nodes, ctypes.c_void_p(views[0]), ctypes.c_void_p(views[1]),
                                                 ctypes.c_void_p(views[2]), _ptr(w_event), heavy, _ptr(arrays["row"]), _ptr(arrays["col"]),
                                                 _ptr(arrays["lab"]), _ptr(arrays["w"]), None, None, None, _ptr(state), _ptr(heavy_list),
                                                 _ptr(res[1]), stream))
        # rows of layer 1 = first-order nodes, owned by node-id range
        offsets = [-(-num_nodes * p // world) for p in range(world + 1)]
        offsets_dev = torch.tensor(offsets, dtype=torch.int64, device=dev)
        total_nodes = num_nodes
        rows_k = torch.arange(offsets[rank], offsets[rank + 1], device=dev).unsqueeze(1)
        starts_before = {j: cuts[j] for j in range(2, K + 1)}   # items of the current level that start before cut j

        for k in range(1, K + 1):
            more = k < K
            which = k & 1
            rec = bufs.rec[which]
            mark(f"route_count[{k}]")
            dstart = torch.empty(world + 1, dtype=torch.int64, device=dev)
            counts = torch.empty(world, dtype=torch.int64, device=dev)
            rows_at, stride = (_ptr(arrays["row"]), 1) if k == 1 else (word(rec, 1), 4)
            _lib.check(lib.ppg_chain_dest_bounds(rows_at, stride, n_slots, _ptr(offsets_dev), world, _ptr(dstart), _ptr(counts), stream))
            # sizes of the next level: row pointer of this level's items in label order
            later = list(range(k + 1, K + 1))
            sel = torch.zeros(len(later) + 1, dtype=torch.int64, device=dev)     # starts_before of the next level + status bits
            if more and n_slots:
                at_items = torch.tensor([min(starts_before[j], n_slots) for j in later], dtype=torch.int64, device=dev)
                if k == 1:
                    time, mode, delta_i, delta_f = ops._time_mode(ext_t.contiguous(), delta)
                    _lib.check(lib.ppg_lift_temporal_count(_ptr(ext_ei), _ptr(time), m_ext, num_nodes, mode | _lib.TIME_GROUPED, delta_i,
                                                           delta_f, _ptr(tws), tws.numel(), None, stream))
                    _lib.check(lib.ppg_chain_node_ptr(ctypes.c_void_p(views[4]), m_ext, _ptr(info), 4, 2, stream))
                    status_word = tws[8:16].view(torch.int64)
                else:
                    ws = scan_ws(n_slots)
                    _lib.check(lib.ppg_chain_scan_nodes(_ptr(info), 4, 2, n_slots, _ptr(ws), ws.numel(),
                                                        ctypes.c_void_p(res[k].data_ptr() + 8 * 4), stream))
                    status_word = res[k, 1:2]
                sel = torch.cat([info.view(torch.int32).view(-1)[at_items * 4 + 2].to(torch.int64), status_word])
            gathered = _gather_counts(torch.cat([counts, dstart, res[k, :4], sel]), group)          # (sync 1)
            matrix, starts, mine = gathered[:, :world], gathered[:, world:2 * world + 1], gathered[rank].tolist()
            recv = matrix[:, rank].tolist()
            base = 2 * world + 1
            if (int(gathered[:, base + 1].max()) | int(gathered[:, -1].max())) & 1:
                raise ValueError("distributed lift: node id outside [0, num_nodes)")
            heavy_slots, heavy_rows = mine[base + 2], mine[base + 3]
            next_starts = {j: mine[base + 4 + i] for i, j in enumerate(later)}
            next_most = int(gathered[:, base + 4].max()) if later else 0      # largest next level of any rank

            mark(f"route_pack[{k}]")
            if k == 1:
                if heavy_rows:   # hub rows: the tiles left them in generation order
                    ws = torch.empty(lib.ppg_chain_heavy_workspace_bytes(heavy_slots, heavy_rows, n_slots), dtype=torch.uint8, device=dev)
                    _lib.check(lib.ppg_chain_heavy_fix(_ptr(heavy_list), heavy_rows, heavy_slots, n_slots, _ptr(arrays["col"]),
                                                       _ptr(arrays["lab"]), _ptr(arrays["w"]), None, None, None, _ptr(ws), ws.numel(), stream))
                _lib.check(lib.ppg_chain_pack(_ptr(arrays["row"]), _ptr(arrays["col"]), _ptr(arrays["col"]), _ptr(arrays["w"]), n_slots,
                                              _ptr(rec), stream))
            elif heavy_rows:
                ws = torch.empty(lib.ppg_chain_heavy_records_workspace_bytes(heavy_slots, heavy_rows, n_slots), dtype=torch.uint8, device=dev)
                _lib.check(lib.ppg_chain_heavy_fix_records(_ptr(heavy_list), heavy_rows, heavy_slots, n_slots, _ptr(rec), _ptr(labS),
                                                           _ptr(firstS), _ptr(degS), _ptr(ws), ws.numel(), stream))
            if k > 1:   # owned rows of this layer = merged edges of the previous one: output work, on the side stream
                with torch.cuda.stream(side):
                    rows_k = ops.extend_owned_rows(rows_k, prev_row_lo, prev_ei[0], prev_last)
                    rows_k.record_stream(main)
            mark(f"records_wait[{k}]")
            lo_at = starts[:, rank].tolist()            # first record for this owner in every sender's buffer
            received = send_back = None
            if peer:
                bufs.barrier()                          # every sender's records are complete (and put in order)
                run_ptrs = [bufs.rec_ptr(which, s, lo_at[s]) for s in range(world)]
                back_ptrs = [bufs.back_ptr(s, lo_at[s]) for s in range(world)]
            elif world == 1:
                run_ptrs, back_ptrs = [bufs.rec_ptr(which, 0, 0)], [bufs.back_ptr(0, 0)]
            else:                                       # no peer access: the records travel through one all-to-all-v
                send = matrix[rank].tolist()
                received = torch.empty((sum(recv), 2), dtype=torch.int64, device=dev)
                dist.all_to_all_single(received.view(-1), rec[:2 * n_slots], [2 * c for c in recv], [2 * c for c in send], group=group)
                send_back = torch.empty(max(sum(recv), 1), dtype=torch.int32, device=dev)
                seg = [0]
                for c in recv:
                    seg.append(seg[-1] + c)
                run_ptrs = [received.data_ptr() + 16 * seg[s] for s in range(world)]
                back_ptrs = [send_back.data_ptr() + 4 * seg[s] for s in range(world)]
            mark(f"merge_sort[{k}]")
            rows_owned = offsets[rank + 1] - offsets[rank]
            merge = _MergeSorted(run_ptrs, recv, back_ptrs, offsets[rank], rows_owned, total_nodes, dev)
            mark(f"merge_sync[{k}]")
            results = _gather_counts(merge.result_words, group)                                     # (sync 2)
            if int(results[:, 1].max()) & 2:   # a row range did not fit a tile somewhere: those owners sort their records
                if int(results[rank, 1]) & 2:
                    copy = torch.cat([bufs.peer_records(which, s, lo_at[s], recv[s]) for s in range(world)]) if received is None \
                        else received
                    merge = ops.merge_records_begin(copy, offsets[rank], rows_owned, total_nodes)
                    at = 0
                    for s in range(world):   # the merged indices go where the tiles would have put them
                        target = bufs.peer_back(s, lo_at[s], recv[s]) if send_back is None else send_back[at:at + recv[s]]
                        target.copy_(merge.inverse[at:at + recv[s]])
                        at += recv[s]
                results = _gather_counts(merge.result_words, group)
            if int(results[:, 1].max()) & 1:
                raise ValueError("distributed lift: a node id outside its layer reached an owner (inconsistent inputs)")
            if peer:
                bufs.barrier()                          # every owner's merged indices have landed in the senders' buffers
            elif send_back is not None:
                dist.all_to_all_single(bufs.back[:n_slots], send_back[:sum(recv)], matrix[rank].tolist(), recv, group=group)
            edge_offsets = [0]
            for c in results[:, 0].tolist():
                edge_offsets.append(edge_offsets[-1] + int(c))
            if edge_offsets[-1] > (1 << 31):
                raise ValueError(f"distributed lift: {edge_offsets[-1]} merged edges exceed the 31-bit ids of the chain")
            edge_offsets_dev = torch.tensor(edge_offsets, dtype=torch.int64, device=dev)
            mark(f"merge_fill[{k}]")
            side.wait_stream(main)   # the merged edges are written out under the next level's expansion
            for t in getattr(merge, "compact", ()):
                t.record_stream(side)
            with torch.cuda.stream(side):
                out_ei, out_w, out_last = merge.finish(edge_offsets[rank + 1] - edge_offsets[rank])
                for t in (out_ei, out_w, out_last):
                    t.record_stream(main)
            if k == 2 and edge_offsets[-1] == 0:
                raise _lib.EmptyLiftError("torch.cat(): expected a non-empty list of Tensors (distributed lift: no time-respecting pair)")
            layers[k] = DistributedLayer(k, total_nodes, offsets[rank], rows_k, out_ei, out_w, edge_offsets[rank], edge_offsets[-1])
            prev_ei, prev_last, prev_row_lo = out_ei, out_last, offsets[rank]
            if not more:
                break

            # ---- the ids are back: local rows of the next level, info words of this level's items, then expand in that order
            mark(f"route_unpack[{k}]")
            rowid, run_start, row_value = empty32(n_slots), empty32(n_slots + 1), empty32(n_slots)
            ws = scan_ws(n_slots)
            _lib.check(lib.ppg_chain_unpack(_ptr(bufs.back), n_slots, _ptr(dstart), _ptr(edge_offsets_dev), world, _ptr(labS),
                                            word(rec, 2), 4, None, _ptr(ws), ws.numel(), _ptr(rowid), _ptr(run_start), _ptr(row_value),
                                            _ptr(info), ctypes.c_void_p(res[k].data_ptr() + 8 * 5), stream))
            mark(f"lift_next[{k}]")
            n_next = next_starts.get(k + 1, 0) if n_slots else 0
            if trace.timed:
                trace.sizes.append((k + 1, n_slots, min(starts_before[k + 1], n_slots), n_next))
            bufs.ensure(next_most, keep=which, keep_slots=n_slots)
            rec = bufs.rec[which]
            nxt_lab = empty32(n_next)
            nxt_first = nxt_deg = nxt_info = None
            nxt_heavy = torch.empty((n_next // (heavy + 1) + 2, 2), dtype=u32, device=dev)
            if k + 1 < K:
                nxt_first, nxt_deg = empty32(n_next), empty32(n_next)
                nxt_info = torch.empty((max(n_next, 1) + 1, 2), dtype=torch.int64, device=dev)
            if n_next:
                ns = n_slots
                offP = torch.empty(ns + 1, dtype=torch.int64, device=dev)
                lblP = empty32(ns)
                srcbound = torch.empty((-(-n_next // tile), 2), dtype=u32, device=dev)
                ws = scan_ws(ns)
                limit = min(starts_before[k + 1], ns)
                if k == 1:   # the counts of the events are in label order (temporal count): gather them into merged order
                    firstP = empty32(ns)
                    _lib.check(lib.ppg_chain_count_sorted(_ptr(labS), ns, ctypes.c_void_p(views[3]), ctypes.c_void_p(views[4]), None, limit,
                                                          _ptr(rowid), _ptr(run_start), _ptr(ws), ws.numel(), _ptr(offP), _ptr(firstP),
                                                          _ptr(lblP), None, _ptr(srcbound), stream))
                    via = ctypes.c_void_p(views[1])
                else:        # this level's tiles left them in merged order
                    firstP = firstS
                    _lib.check(lib.ppg_chain_count_sorted_next(_ptr(labS), ns, _ptr(degS), _ptr(info), 4, 2, limit, _ptr(rowid),
                                                               _ptr(run_start), _ptr(ws), ws.numel(), _ptr(offP), _ptr(lblP),
                                                               _ptr(srcbound), stream))
                    via = None
                state = torch.empty(-(-n_next // tile) + 1, dtype=torch.int64, device=dev)
                _lib.check(lib.ppg_chain_tiles_dist(ns, n_next, _ptr(offP), _ptr(firstP), _ptr(lblP), word(rec, 3), 4, _ptr(run_start),
                                                    _ptr(rowid), _ptr(row_value), _ptr(info), via, _ptr(srcbound), heavy,
                                                    _ptr(bufs.rec[1 - which]), _ptr(nxt_lab), _ptr(nxt_first), _ptr(nxt_deg), _ptr(nxt_info),
                                                    _ptr(state), _ptr(nxt_heavy), _ptr(res[k + 1]), stream))
            labS, firstS, degS, heavy_list, n_slots = nxt_lab, nxt_first, nxt_deg, nxt_heavy, n_next
            if nxt_info is not None:
                info = nxt_info
            starts_before = next_starts
            offsets, offsets_dev, total_nodes = edge_offsets, edge_offsets_dev, edge_offsets[-1]
        main.wait_stream(side)
    return layers


def distributed_temporal_layers(edge_index: torch.Tensor, time: torch.Tensor, num_nodes: int, delta, max_order: int,
                                edge_weight: torch.Tensor | None = None, group=None, local_ops=None) -> dict:
    """``MultiOrderModel.from_temporal_graph`` over a stream that is split across the ranks.

    ``edge_index`` [2, m_local] / ``time`` [m_local]: this rank's CONTIGUOUS range of the globally
    time-sorted stream (ranges in rank order).  Returns ``{order: DistributedLayer}``.

    Per order there is ONE exchange (``_exchange_merge``): the line-graph edges of the extended (own + ghost) range
    travel to the owner of their source row as (source id, target id, weight, last node) and are merged there; the
    owner returns the global index of the merged edge, which IS the node id of the next order, so no k-gram row ever
    has to be ranked or sent.  Paths that start with a ghost event travel with weight 0: they only collect their ids
    (the rank that owns their first event contributes the weight), so every path is counted exactly once.

    Only what later orders need is lifted: at order j a path is needed if it starts no later than
    ``t_last(own) + (K - j) delta`` (its id is the target of a path one order up), and such paths are a PREFIX of the
    line graph, whose columns are ascending in the source (``limit_sources``).

    Limits: float32 weights; fewer than 2^32 nodes per layer (32-bit record fields)."""
    if local_ops is None:
        from . import ops as local_ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = edge_index.device
    K = int(max_order)
    m_own = edge_index.size(1)
    if edge_weight is not None and edge_weight.dtype != torch.float32:
        raise TypeError("distributed_temporal_layers: edge weights must be float32")
    if num_nodes > (1 << 32):
        raise ValueError("distributed_temporal_layers: more than 2^32 first-order nodes")
    layers: dict[int, DistributedLayer] = {}
    trace = _Trace(dev)
    trace.mark("ghost_zone")

    # ---- ghost zone once: everything after it is local lifting + one exchange per order
    if K > 1:
        ext_ei, ext_t, ext_w = exchange_ghost_zone(edge_index, time, edge_weight, delta * (K - 1), group)
    else:
        ext_ei, ext_t, ext_w = edge_index.contiguous(), time, edge_weight
    m_ext = ext_ei.size(1)

    # ---- cuts: events that may START a path needed at order j (see above); exact for integer time stamps, no pruning otherwise
    exact = (not time.is_floating_point()) and isinstance(delta, int)
    cuts = {K: m_own}
    if 1 < K:
        if exact and m_own > 0:
            limits = torch.tensor([(K - j) * delta for j in range(2, K)], dtype=time.dtype, device=dev) + time[-1]
            found = torch.searchsorted(ext_t, limits, right=True).tolist() if K > 2 else []
        else:
            found = [m_ext if m_own > 0 else 0] * (K - 2)
        for j, c in zip(range(2, K), found):
            cuts[j] = int(c)

    if _chain_supported(ext_ei, ext_w, num_nodes, None if getattr(local_ops, "__name__", "") == "pathpyg_b200.ops" else local_ops):
        layers = _distributed_chain_layers(ext_ei, ext_t, ext_w, m_own, num_nodes, delta, K, cuts, group, trace)
        trace.close()
        return layers

    # ---- order 1: nodes are the first-order nodes themselves, rows owned by node-id range
    n1_bounds = [-(-num_nodes * p // world) for p in range(world + 1)]
    n1_dev = torch.tensor(n1_bounds, dtype=torch.int64, device=dev)
    lo = n1_bounds[rank]

    def first_lift():
        if m_ext == 0:
            return None
        return local_ops.lift_order_temporal_begin(ext_ei, ext_t, delta, num_nodes, assume_sorted=True, limit_sources=cuts[2])

    ei_k, w_k, last_k, ids, offsets, offsets_dev, line_index = _exchange_merge(
        ext_ei, None, ext_w, m_own, n1_bounds, n1_dev, num_nodes, local_ops, group,
        lift_begin=first_lift if K > 1 and m_ext > 0 else None, trace=trace, level=1)
    rows_k, row_lo = torch.arange(lo, n1_bounds[rank + 1], device=dev).unsqueeze(1), lo
    layers[1] = DistributedLayer(1, num_nodes, lo, rows_k, ei_k, w_k, offsets[rank], offsets[-1])
    if K == 1:
        trace.close()
        return layers
    if line_index is None:
        line_index = torch.empty((2, 0), dtype=torch.int64, device=dev)

    # ---- orders 2..K: the line-graph nodes of level k are the edges of level k - 1 (level 1: the events)
    line_w = None
    if ext_w is not None:
        line_w = local_ops.pair_attributes(line_index, ext_w, "src", index_bound=m_ext) if line_index.size(1) else ext_w[:0]
    num_line_nodes = m_ext
    # prefix[j] = number of level-k line-graph NODES that start before cut j (level 2: the events themselves)
    prefix = dict(cuts)
    for k in range(2, K + 1):
        # the owned rows of layer k = the merged edges of layer k - 1 (local work while the ids of level k - 1 travel)
        trace.mark(f"rows[{k}]")
        rows_k = local_ops.extend_owned_rows(rows_k, row_lo, ei_k[0], last_k)
        total, row_lo = offsets[-1], offsets[rank]
        info = ids.finish()
        # number of level-k EDGES (= level-(k+1) nodes) that start before every later cut
        later = [j for j in range(k + 1, K + 1)]
        if later and line_index.size(1):
            at = torch.tensor([prefix[j] for j in later], dtype=torch.int64, device=dev)
            nxt_prefix = dict(zip(later, torch.searchsorted(line_index[0].contiguous(), at, right=False).tolist()))
        else:
            nxt_prefix = {j: 0 for j in later}

        def next_lift(line_index=line_index, num_line_nodes=num_line_nodes, limit=nxt_prefix.get(k + 1, 0)):
            return local_ops.lift_order_edge_index_begin(line_index, num_line_nodes, limit_sources=limit)

        ei_k, w_k, last_k, ids, offsets, offsets_dev, nxt = _exchange_merge(
            line_index, info, line_w, prefix[K], offsets, offsets_dev, total, local_ops, group,
            lift_begin=next_lift if k < K and line_index.size(1) else None, trace=trace, level=k)
        if k == 2 and offsets[-1] == 0:
            # no time-respecting pair on ANY rank: the single-device build fails in lift_order_temporal (temporal.py:53)
            raise _lib.EmptyLiftError("torch.cat(): expected a non-empty list of Tensors (distributed lift: no time-respecting pair)")
        layers[k] = DistributedLayer(k, total, row_lo, rows_k, ei_k, w_k, offsets[rank], offsets[-1])
        if k == K:
            break
        if nxt is None:
            nxt = torch.empty((2, 0), dtype=torch.int64, device=dev)
        if line_w is not None:
            line_w = local_ops.pair_attributes(nxt, line_w, "src", index_bound=line_index.size(1)) if nxt.size(1) else line_w[:0]
        num_line_nodes, line_index, prefix = line_index.size(1), nxt, nxt_prefix
    trace.close()
    return layers
