"""Plain-torch stand-ins (CPU or CUDA tensors) for the functions of ``pathpyg_b200.ops`` that the distributed lift
calls: same call signatures and the same 16-byte record format as ``csrc/exchange.cu``.  Test infrastructure only:
the gloo tests inject it as ``local_ops``; the GPU tests compare the CUDA kernels against it."""
import torch

from oracle import lift


class _TorchRoutePlan:
    def __init__(self, line_index, node_info, offsets, world):
        self.li, self.node_info, self.offsets, self.world = line_index, node_info, offsets, world
        self.E = line_index.size(1)
        ids = line_index[0] if node_info is None else node_info[line_index[0]] >> 32
        self.owner = (torch.searchsorted(offsets, ids, right=True) - 1).clamp_(0, world - 1)
        self.counts = torch.bincount(self.owner, minlength=world)
        self.order = torch.sort(self.owner, stable=True).indices

    def pack(self, weights, own_prefix):
        s, t = self.li[0], self.li[1]
        if self.node_info is None:
            src, dst, last = s, t, t
            ghost = torch.arange(self.E, device=s.device) >= own_prefix
        else:
            src, dst, last = self.node_info[s] >> 32, self.node_info[t] >> 32, self.node_info[t] & 0xffffffff
            ghost = s >= own_prefix
        w = torch.ones(self.E, device=s.device) if weights is None else weights.clone()
        w[ghost] = 0.0
        bits = w.view(torch.int32).to(torch.int64) & 0xffffffff
        records = torch.stack([(src << 32) | dst, (bits << 32) | last], dim=1)
        self.slot = torch.empty(self.E, dtype=torch.int64, device=s.device)
        self.slot[self.order] = torch.arange(self.E, device=s.device)
        self.last = last
        return records[self.order].contiguous()

    def unpack(self, back, edge_offsets):
        ids = back.long()[self.slot] + edge_offsets[self.owner]
        return (ids << 32) | self.last


class _TorchMerge:
    def __init__(self, records, row_lo, rows_owned, total_nodes):
        self.row_lo, self.total = row_lo, max(total_nodes, 1)
        src, dst = records[:, 0] >> 32, records[:, 0] & 0xffffffff
        self.w = ((records[:, 1] >> 32) & 0xffffffff).to(torch.int32).view(torch.float32) if records.size(0) else torch.empty(0, device=records.device)
        self.last = records[:, 1] & 0xffffffff
        bad = bool(((src < row_lo) | (src >= row_lo + rows_owned) | (dst >= total_nodes)).any()) if records.size(0) else False
        keys = (src - row_lo) * self.total + dst
        self.uniq, inverse = torch.unique(keys, return_inverse=True)
        self.inverse = inverse.int()
        self.result_words = torch.tensor([self.uniq.numel(), int(bad)], dtype=torch.int64, device=records.device)

    def finish(self, num_out):
        assert num_out == self.uniq.numel()
        ei = torch.stack([self.uniq // self.total + self.row_lo, self.uniq % self.total])
        # weights are summed in arrival order, as the stable device sort does
        w = torch.zeros(num_out, device=self.w.device).index_add_(0, self.inverse.long(), self.w)
        last = torch.zeros(num_out, dtype=torch.int64, device=self.w.device)
        last[self.inverse.long()] = self.last
        return ei, w, last


class _TorchPendingLift:
    def __init__(self, out):
        self.out = out
        self.result_words = torch.tensor([out.size(1), 0], dtype=torch.int64, device=out.device)

    def finish(self, total=None, status=0, allow_empty=False):
        assert total is None or total == self.out.size(1)
        if self.out.size(1) == 0 and not allow_empty:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
        return self.out


class TorchOps:
    """CPU stand-in (plain torch + the oracle's lift) for the functions of pathpyg_b200.ops that the distributed lift
    calls, with the same call signatures and record format."""

    @staticmethod
    def lift_order_temporal_begin(edge_index, time, delta, num_nodes, assume_sorted=True, limit_sources=None):
        try:
            out = lift.lift_order_temporal(edge_index, time, delta)
        except (RuntimeError, ValueError):
            out = edge_index.new_empty((2, 0))
        return _TorchPendingLift(out if limit_sources is None else out[:, out[0] < limit_sources].contiguous())

    @staticmethod
    def lift_order_edge_index_begin(edge_index, num_nodes, limit_sources=None):
        if edge_index.size(1) == 0:
            return _TorchPendingLift(edge_index.new_empty((2, 0)))
        out = lift.lift_order_edge_index(edge_index, num_nodes)
        return _TorchPendingLift(out if limit_sources is None else out[:, out[0] < limit_sources].contiguous())

    @staticmethod
    def pair_attributes(edge_index, attr, rule, index_bound=None):
        return lift.aggregate_node_attributes(edge_index, attr, rule)

    route_plan = staticmethod(_TorchRoutePlan)
    merge_records_begin = staticmethod(_TorchMerge)

    @staticmethod
    def extend_owned_rows(prev_rows, prev_row_lo, src_ids, last):
        return torch.cat([prev_rows[src_ids - prev_row_lo], last.unsqueeze(1)], dim=1)
