"""The kernels of the cross-partition exchange (csrc/exchange.cu) against their plain-torch restatement
(tests/torch_local_ops.py), and the whole distributed lift on ONE device (a one-rank NCCL group: every record is
routed to the rank itself, so the count / pack / merge / unpack chain runs exactly as on N ranks) against the
single-device ``MultiOrderModel.from_temporal_graph``.  The N-rank NCCL run is in test_parallel_gpu.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

import pathpyg_b200 as pp
from pathpyg_b200 import ops, parallel
from torch_local_ops import TorchOps

pytestmark = pytest.mark.gpu


def _line_graph(gen, n_line, E, cuda):
    src = torch.sort(torch.randint(0, n_line, (E,), generator=gen)).values
    dst = torch.randint(0, n_line, (E,), generator=gen)
    return torch.stack([src, dst]).to(cuda)


@pytest.mark.parametrize("E,n_line,world,first_level", [
    (0, 10, 3, False), (1, 5, 1, True), (1000, 300, 4, False), (5000, 50, 16, True),
    (300_000, 80_000, 8, False), (2_500_000, 1_000_000, 5, False), (1_100_000, 4_000, 2, True),
])
def test_route_pack_unpack_matches_torch(cuda, E, n_line, world, first_level):
    gen = torch.Generator().manual_seed(E + world)
    li = _line_graph(gen, n_line, E, cuda)
    total_ids = 3 * n_line + 7
    if first_level:
        info, id_space = None, n_line
    else:
        ids = torch.randint(0, total_ids, (n_line,), generator=gen)
        last = torch.randint(0, 1 << 20, (n_line,), generator=gen)
        info, id_space = ((ids << 32) | last).to(cuda), total_ids
    cutsr = torch.sort(torch.randint(0, id_space + 1, (world - 1,), generator=gen)).values.tolist()
    offsets = torch.tensor([0] + cutsr + [id_space], dtype=torch.int64, device=cuda)      # includes empty ranges
    w = torch.randint(1, 9, (E,), generator=gen).float().to(cuda)
    own_prefix = (E if first_level else n_line) * 2 // 3

    got_plan, want_plan = ops.route_plan(li, info, offsets, world), TorchOps.route_plan(li, info, offsets, world)
    assert torch.equal(got_plan.counts, want_plan.counts)
    for weights in (w, None):
        got, want = got_plan.pack(weights, own_prefix), want_plan.pack(weights, own_prefix)
        assert torch.equal(got, want)
        assert torch.equal(got_plan.slot.long(), want_plan.slot)
    back = torch.randint(0, 1 << 20, (E,), generator=gen).int().to(cuda)
    edge_offsets = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64), torch.randint(0, 1 << 25, (world,), generator=gen)]), 0).to(cuda)
    assert torch.equal(got_plan.unpack(back, edge_offsets), want_plan.unpack(back, edge_offsets))


@pytest.mark.parametrize("R,rows,total,row_lo", [(0, 5, 9, 3), (1, 1, 1, 0), (4000, 60, 200, 1000), (700_000, 5000, 40_000, 123_456),
                                                 (3_000_000, 2_000_000, 9_000_000, 5_000_000)])
def test_merge_records_matches_torch(cuda, R, rows, total, row_lo):
    gen = torch.Generator().manual_seed(R)
    src = torch.randint(row_lo, row_lo + rows, (R,), generator=gen)
    dst = torch.randint(0, total, (R,), generator=gen)
    if R > 10:  # force duplicates
        src[R // 2:] = src[:R - R // 2]
        dst[R // 2:] = dst[:R - R // 2]
    w = torch.randint(0, 5, (R,), generator=gen).float()
    # `last` is a function of the (src, dst) pair in the real exchange (duplicates carry the same value)
    last = (src * 31 + dst * 17) % 1000
    bits = w.view(torch.int32).to(torch.int64) & 0xffffffff
    records = torch.stack([(src << 32) | dst, (bits << 32) | last], dim=1).to(cuda)
    got, want = ops.merge_records_begin(records, row_lo, rows, total), TorchOps.merge_records_begin(records, row_lo, rows, total)
    assert torch.equal(got.result_words, want.result_words.to(cuda))
    assert torch.equal(got.inverse, want.inverse.to(cuda))
    n_out = int(want.result_words[0])
    for a, b in zip(got.finish(n_out), want.finish(n_out)):
        assert torch.equal(a, b.to(cuda))


def test_merge_records_flags_foreign_rows(cuda):
    records = torch.tensor([[(7 << 32) | 1, 0], [(2 << 32) | 1, 0]], dtype=torch.int64, device=cuda)
    m = ops.merge_records_begin(records, 5, 10, 4)      # row 2 is not in [5, 15)
    assert int(m.result_words[1]) & 1


def test_extend_owned_rows(cuda):
    gen = torch.Generator().manual_seed(1)
    prev = torch.randint(0, 99, (500, 3), generator=gen).to(cuda)
    src = torch.randint(40, 540, (2000,), generator=gen).to(cuda)
    last = torch.randint(0, 99, (2000,), generator=gen).to(cuda)
    assert torch.equal(ops.extend_owned_rows(prev, 40, src, last), torch.cat([prev[src - 40], last[:, None]], 1))


@pytest.mark.parametrize("limit", [0, 1, 777, 10_000])
def test_prefix_limited_lifts(cuda, limit):
    gen = torch.Generator().manual_seed(3)
    n, m = 300, 10_000
    ei = torch.randint(0, n, (2, m), generator=gen).to(cuda)
    t = torch.sort(torch.randint(0, 500, (m,), generator=gen)).values.to(cuda)
    full = ops.lift_order_temporal(ei, t, 7, n)
    part = ops.lift_order_temporal(ei, t, 7, n, assume_sorted=True, limit_sources=limit, allow_empty=True)
    assert torch.equal(part, full[:, full[0] < limit])
    full3 = ops.lift_order_edge_index(full, m)
    cut = min(limit * 3, full.size(1))
    assert torch.equal(ops.lift_order_edge_index(full, m, limit_sources=cut), full3[:, full3[0] < cut])


@pytest.fixture(scope="module")
def one_rank_group():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("n,m,horizon,delta,K,weighted", [(40, 3000, 300, 4, 4, True), (3000, 400_000, 4000, 60, 3, False),
                                                          (100_000, 1_000_000, 1000, 200, 2, False), (25, 400, 50, 1.5, 3, True)])
def test_distributed_lift_on_one_rank_equals_single_device(cuda, one_rank_group, n, m, horizon, delta, K, weighted):
    gen = torch.Generator().manual_seed(m)
    ei = torch.randint(0, n, (2, m), generator=gen).to(cuda)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=gen)).values.to(cuda)
    attrs = {"edge_weight": torch.randint(1, 4, (m,), generator=gen).float().to(cuda)} if weighted else {}
    want = pp.MultiOrderModel.from_temporal_graph(pp.TemporalGraph.from_tensors(ei, t, n, **attrs), delta=delta, max_order=K)
    got = parallel.distributed_temporal_layers(ei, t, n, delta, K, edge_weight=attrs.get("edge_weight"))
    for k, layer in want.layers.items():
        assert got[k].num_nodes == layer.n and got[k].row_offset == 0, k
        assert torch.equal(got[k].node_sequence, layer.data.node_sequence), k
        assert torch.equal(got[k].edge_index, layer.data.edge_index.as_tensor()), k
        assert torch.equal(got[k].edge_weight, layer.data.edge_weight), k


def test_distributed_lift_without_any_pair_raises(cuda, one_rank_group):
    ei = torch.tensor([[0, 1], [1, 2]], device=cuda)
    t = torch.tensor([5, 5], device=cuda)
    with pytest.raises((RuntimeError, ValueError)):
        parallel.distributed_temporal_layers(ei, t, 3, 1, 2)
