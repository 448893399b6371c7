"""Layer chain of ``MultiOrderModel.from_temporal_graph`` without a global sort per order (``csrc/chain.cu``).

Reference: ``src/pathpyG/core/multi_order_model.py:124-192`` -- per order one ``lift_order_temporal`` /
``lift_order_edge_index`` (temporal.py:17-54, lift_order.py:48-79) and one ``aggregate_edge_index``
(lift_order.py:109-152).  Here the line graph of order k is never materialised as an int64 edge index and never sorted:
its pairs are generated in the merged order of their source items, which leaves them grouped by De Bruijn row, and are
ranked by column inside shared-memory tiles of whole rows.  What a level keeps on the device:

    slot arrays (final (row, col) order)   rowS, colS, labS (= merged order P of the next level), idS, wS
    item arrays (reference line-graph order) id_item (= ``inverse_idx`` of the next layer), tail, w_item
    run_start                                first slot of every merged edge

One host synchronisation per order (merged edges of this order + pairs of the next, read together).
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib, ops
from .ops import _ptr, _stream

_u32 = torch.int32  # 32-bit words; the kernels read them unsigned (all values < 2^31)

_RES_HEADS, _RES_STATUS, _RES_HEAVY_SLOTS, _RES_HEAVY_ROWS, _RES_NEXT = 0, 1, 2, 3, 4


_side_streams: dict = {}


def side_stream(dev: torch.device) -> torch.cuda.Stream:
    """One side stream per device for the output work of a finished layer (fill, node sequences)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(dev)
    return _side_streams[key]


def heavy_threshold() -> int:
    """Rows with more pairs than this are ordered by the radix-sort fallback (``PPG_CHAIN_HEAVY`` overrides, tests)."""
    lib = _lib.load()
    return max(1, min(int(os.environ.get("PPG_CHAIN_HEAVY", lib.ppg_chain_heavy_default())), lib.ppg_chain_heavy_default()))


class _Level:
    """Device arrays of one level (see the module docstring)."""

    def __init__(self, items: int, dev, weighted: bool, keep_items: bool, heavy: int):
        self.items = items
        n = max(items, 1)
        self.rowS = torch.empty(n, dtype=_u32, device=dev)
        self.colS = torch.empty(n, dtype=_u32, device=dev)
        self.labS = torch.empty(n, dtype=_u32, device=dev)
        self.wS = torch.empty(n, dtype=torch.float32, device=dev) if weighted else None
        self.run_start = torch.empty(n + 1, dtype=_u32, device=dev)
        self.heavy_list = torch.empty((n // (heavy + 1) + 2, 2), dtype=_u32, device=dev)
        # needed by the next level only
        self.idS = torch.empty(n, dtype=_u32, device=dev) if keep_items else None
        self.node = torch.empty(n + 1, dtype=torch.int64, device=dev) if keep_items else None   # by label: id | count / pointer << 32
        self.firstS = self.degS = None   # slot order: first continuation and their number (levels >= 2 with a level above)
        self.merged = 0


class TemporalChain:
    """Builds the layers 1..K of a time-sorted event stream on one device."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, edge_weight: torch.Tensor | None, max_order: int):
        self.lib = _lib.load()
        self.ei, self.n, self.w, self.K = edge_index, int(num_nodes), edge_weight, int(max_order)
        self.dev, self.m = edge_index.device, edge_index.size(1)
        self.heavy = heavy_threshold()
        self.tile = self.lib.ppg_chain_tile_slots()
        self.res = torch.zeros((self.K + 2, 8), dtype=torch.int64, device=self.dev)
        self.weighted = edge_weight is not None

    # ------------------------------------------------------------------ helpers
    def _scan_ws(self, n: int) -> torch.Tensor:
        return torch.empty(self.lib.ppg_chain_scan_workspace_bytes(n), dtype=torch.uint8, device=self.dev)

    def _tile_state(self, slots: int) -> torch.Tensor:
        return torch.empty(-(-slots // self.tile) + 1, dtype=torch.int64, device=self.dev)

    def _heavy_fix(self, level: _Level, k: int, rows: int, slots: int) -> int:
        """Order the rows the tiles skipped, redo the run heads; returns the merged count (one more synchronisation)."""
        ws = torch.empty(self.lib.ppg_chain_heavy_workspace_bytes(slots, rows, level.items), dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.ppg_chain_heavy_fix(_ptr(level.heavy_list), rows, slots, level.items, _ptr(level.colS), _ptr(level.labS),
                                                _ptr(level.wS), _ptr(level.firstS), _ptr(level.degS), None, _ptr(ws), ws.numel(),
                                                _stream(self.dev)))
        ws = self._scan_ws(level.items)
        _lib.check(self.lib.ppg_chain_heads(_ptr(level.rowS), _ptr(level.colS), _ptr(level.labS), level.items, _ptr(ws), ws.numel(),
                                            _ptr(level.idS), _ptr(level.node), 2, _ptr(level.run_start), _ptr(self.res[k]),
                                            _stream(self.dev)))
        return int(self.res[k, _RES_HEADS].item())

    def _fill(self, level: _Level):
        out_ei = torch.empty((2, level.merged), dtype=torch.int64, device=self.dev)
        out_w = torch.empty(level.merged, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.ppg_chain_fill(_ptr(level.rowS), _ptr(level.colS), _ptr(level.wS), _ptr(level.run_start), level.merged,
                                           _ptr(out_ei), _ptr(out_w), _stream(self.dev)))
        return out_ei, out_w

    def _emit(self, store, k: int, level: _Level, nodes: int, inverse) -> None:
        """Finish layer k: merged edges + whatever ``store`` derives from them.  (On one device this output work stays on
        the main stream: moved to a side stream under the next level's expansion it gained nothing at 10M and 50M events
        -- both are bound by the same memory system -- and cost 80 us of stream hand-offs at 1M.)"""
        out_ei, out_w = self._fill(level)
        store(k, out_ei, out_w, nodes, inverse)

    def inverse_idx(self, level: _Level) -> torch.Tensor:
        """``inverse_idx`` of the layer above ``level``: the merged id of every item (the id word of its node word),
        int64 as in the reference."""
        out = torch.empty(level.items, dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.ppg_chain_widen(_ptr(level.node), 2, level.items, _ptr(out), _stream(self.dev)))
        return out

    # ------------------------------------------------------------------ the build
    def run(self, time: torch.Tensor, delta, store, grouped_ws: torch.Tensor | None = None, before_time=None,
            cached: bool = True) -> None:
        """``store(order, edge_index, edge_weight, num_nodes, inverse_idx_or_None)`` receives every finished layer
        (inverse_idx None for order 1, and for the layers a non-``cached`` build drops).  ``grouped_ws``: workspace of
        ``ops.lift_order_temporal_group`` if the caller grouped the events already (staged upload); ``before_time()`` is
        called before the first kernel that reads the time stamps."""
        lib, dev, m, n, K = self.lib, self.dev, self.m, self.n, self.K
        with torch.cuda.device(dev):
            stream = _stream(dev)
            tws = grouped_ws if grouped_ws is not None else ops.lift_order_temporal_group(self.ei, n)
            views = (ctypes.c_void_p * 6)()
            _lib.check(lib.ppg_lift_temporal_views(_ptr(tws), m, n, views))
            ptr1, grouped, sorted_src, first2, off2, _ = (ctypes.c_void_p(v) for v in views)
            # ---- level 1: events grouped by source node, ranked by target node inside every source group
            torch.cuda.nvtx.range_push("chain.level[1]")     # NVTX range per order (free without a profiler attached)
            l1 = _Level(m, dev, self.weighted, K > 1, self.heavy)
            _lib.check(lib.ppg_chain_first_tiles(_ptr(self.ei), m, n, ptr1, grouped, sorted_src, _ptr(self.w), self.heavy,
                                                 _ptr(l1.rowS), _ptr(l1.colS), _ptr(l1.labS), _ptr(l1.wS), _ptr(l1.idS), _ptr(l1.node),
                                                 _ptr(l1.run_start), _ptr(self._tile_state(m)), _ptr(l1.heavy_list), _ptr(self.res[1]),
                                                 stream))
            if K > 1:
                if before_time is not None:
                    before_time()
                time, mode, delta_i, delta_f = ops._time_mode(time.contiguous(), delta)
                _lib.check(lib.ppg_lift_temporal_count(_ptr(self.ei), _ptr(time), m, n, mode | _lib.TIME_GROUPED, delta_i, delta_f,
                                                       _ptr(tws), tws.numel(), None, stream))
                _lib.check(lib.ppg_chain_node_ptr(off2, m, _ptr(l1.node), 2, 1, stream))     # node word of an event: id | first pair
                words = torch.cat([self.res[1, :4], tws[:16].view(torch.int64)]).tolist()      # the synchronisation of order 1
            else:
                words = self.res[1, :4].tolist() + [0, 0]
            if (words[_RES_STATUS] | words[5]) & 1:
                raise ValueError("from_temporal_graph: node id outside [0, num_nodes)")
            l1.merged = words[_RES_HEADS]
            if words[_RES_HEAVY_ROWS]:
                l1.merged = self._heavy_fix(l1, 1, words[_RES_HEAVY_ROWS], words[_RES_HEAVY_SLOTS])
            self._emit(store, 1, l1, n, None)
            torch.cuda.nvtx.range_pop()
            if K == 1:
                return
            pairs = words[4]
            if pairs == 0:
                raise _lib.EmptyLiftError("torch.cat(): expected a non-empty list of Tensors "
                                          "(lift_order_temporal: no time-respecting pair for this delta)")

            # ---- levels 2..K
            prev = l1
            for k in range(2, K + 1):
                more = k < K
                if pairs == 0:   # nothing continues: this layer has nodes but no edges, the ones above are empty
                    for j in range(k, K + 1):
                        nodes = prev.merged if j == k else 0
                        inv = self.inverse_idx(prev) if j == k else torch.empty(0, dtype=torch.int64, device=dev)
                        store(j, torch.empty((2, 0), dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.float32, device=dev),
                              nodes, inv)
                    return
                torch.cuda.nvtx.range_push(f"chain.level[{k}]")
                cur = _Level(pairs, dev, self.weighted, more, self.heavy)
                ns = prev.items
                offP = torch.empty(ns + 1, dtype=torch.int64, device=dev)
                lblP = torch.empty(ns, dtype=_u32, device=dev)
                srcbound = torch.empty((-(-pairs // self.tile), 2), dtype=_u32, device=dev)
                ws = self._scan_ws(ns)
                if k == 2:   # the counts of the events are in label order (temporal count): gather them into merged order
                    firstP = torch.empty(ns, dtype=_u32, device=dev)
                    _lib.check(lib.ppg_chain_count_sorted(_ptr(prev.labS), ns, first2, off2, None, ns, _ptr(prev.idS),
                                                          _ptr(prev.run_start), _ptr(ws), ws.numel(), _ptr(offP), _ptr(firstP),
                                                          _ptr(lblP), None, _ptr(srcbound), stream))
                    via = grouped
                else:        # the previous level's tiles left them in merged order
                    firstP = prev.firstS
                    _lib.check(lib.ppg_chain_count_sorted_next(_ptr(prev.labS), ns, _ptr(prev.degS), _ptr(prev.node), 2, 1, ns, _ptr(prev.idS),
                                                               _ptr(prev.run_start), _ptr(ws), ws.numel(), _ptr(offP), _ptr(lblP),
                                                               _ptr(srcbound), stream))
                    via = None
                if more:
                    cur.firstS = torch.empty(pairs, dtype=_u32, device=dev)
                    cur.degS = torch.empty(pairs, dtype=_u32, device=dev)
                # a pair inherits the weight of its source item, which the previous level holds in slot (= merged) order
                _lib.check(lib.ppg_chain_tiles(ns, prev.merged, pairs, _ptr(offP), _ptr(firstP), _ptr(lblP), _ptr(prev.wS),
                                               _ptr(prev.run_start), _ptr(prev.idS), _ptr(prev.node), via, _ptr(srcbound), self.heavy,
                                               _ptr(cur.rowS), _ptr(cur.colS), _ptr(cur.labS), _ptr(cur.wS), _ptr(cur.idS),
                                               _ptr(cur.node), _ptr(cur.firstS), _ptr(cur.degS), _ptr(cur.run_start),
                                               _ptr(self._tile_state(pairs)), _ptr(cur.heavy_list), _ptr(self.res[k]), stream))
                if more:     # label-order scan of the counts: row pointer of the next level + its number of pairs
                    ws2 = self._scan_ws(pairs)
                    _lib.check(lib.ppg_chain_scan_nodes(_ptr(cur.node), 2, 1, pairs, _ptr(ws2), ws2.numel(),
                                                        ctypes.c_void_p(self.res[k].data_ptr() + 8 * _RES_NEXT), stream))
                inverse = self.inverse_idx(prev) if cached or not more else None
                words = self.res[k].tolist()                                                      # the synchronisation of order k
                cur.merged = words[_RES_HEADS]
                if words[_RES_HEAVY_ROWS]:
                    cur.merged = self._heavy_fix(cur, k, words[_RES_HEAVY_ROWS], words[_RES_HEAVY_SLOTS])
                self._emit(store, k, cur, prev.merged, inverse)
                del offP, firstP, lblP, srcbound
                torch.cuda.nvtx.range_pop()
                prev, pairs = cur, words[_RES_NEXT]
            del tws
