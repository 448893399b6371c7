"""The generation-order layer chain (csrc/chain.cu, pathpyg_b200/chain.py) against the oracle: every layer of
``MultiOrderModel.from_temporal_graph`` bit for bit -- edge index, weights, node sequences, inverse index -- with unit,
integer and arbitrary float32 weights, with the heavy-row fallback forced on, on hub-dominated streams, and against the
sort-per-order chain it replaces (``PPG_CHAIN=0``)."""
import pytest
import torch

import pathpyg_b200 as pp
from oracle import lift, mom

pytestmark = pytest.mark.gpu


def assert_layers_equal(model, want):
    assert sorted(model.layers) == sorted(want)
    for k, layer in want.items():
        got = model.layers[k].data
        assert got.num_nodes == layer.num_nodes, k
        assert torch.equal(got.edge_index.as_tensor().cpu(), layer.edge_index), k
        assert torch.equal(got.edge_weight.cpu(), layer.edge_weight), k
        assert torch.equal(got.node_sequence.cpu(), layer.node_sequence), k
        assert torch.equal(got.inverse_idx.cpu(), layer.inverse_idx), k


def stream(seed, n, m, horizon):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=g)).values
    return ei, t, g


@pytest.mark.parametrize("heavy", [None, 1, 3])
@pytest.mark.parametrize("seed,n,m,horizon,delta,K", [(0, 20, 100, 50, 2, 2), (1, 30, 400, 40, 3, 4), (2, 200, 6000, 300, 4, 3),
                                                       (3, 12, 300, 25, 2, 5), (4, 3, 500, 60, 2, 3), (5, 500, 3000, 20, 1, 4)])
def test_chain_vs_oracle(cuda, monkeypatch, seed, n, m, horizon, delta, K, heavy):
    if heavy is not None:
        monkeypatch.setenv("PPG_CHAIN_HEAVY", str(heavy))   # rows above `heavy` pairs take the radix-sort fallback
    ei, t, g = stream(seed, n, m, horizon)
    want = mom.from_temporal_graph(ei, t, n, delta=delta, max_order=K)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    assert_layers_equal(pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=K), want)
    top = pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=K, cached=False)
    assert list(top.layers) == [K]
    assert torch.equal(top.layers[K].data.edge_index.as_tensor().cpu(), want[K].edge_index)
    assert torch.equal(top.layers[K].data.inverse_idx.cpu(), want[K].inverse_idx)


@pytest.mark.parametrize("heavy", [None, 2])
@pytest.mark.parametrize("kind", ["integer", "float"])
def test_chain_weighted_vs_oracle(cuda, monkeypatch, kind, heavy):
    """Weights travel with the first event of a path; a run is summed in stream order, like the reference's stable
    coalesce, so arbitrary float32 weights agree to the bit as well."""
    if heavy is not None:
        monkeypatch.setenv("PPG_CHAIN_HEAVY", str(heavy))
    ei, t, g = stream(11, 15, 900, 80, )
    w = torch.randint(1, 4, (900,), generator=g).float() if kind == "integer" else torch.rand(900, generator=g) * 3 + 0.1
    want = mom.from_temporal_graph(ei, t, 15, delta=3, max_order=4, edge_weight=w)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), 15, edge_weight=w.to(cuda))
    assert_layers_equal(pp.MultiOrderModel.from_temporal_graph(tg, delta=3, max_order=4), want)


def test_chain_hub_stream(cuda):
    """A star: every event touches node 0, so single rows hold thousands of pairs (several tiles long)."""
    g = torch.Generator().manual_seed(5)
    n, m = 40, 4000
    other = torch.randint(1, n, (m,), generator=g)
    flip = torch.rand(m, generator=g) < 0.5
    ei = torch.stack([torch.where(flip, other, torch.zeros_like(other)), torch.where(flip, torch.zeros_like(other), other)])
    t = torch.sort(torch.randint(0, 400, (m,), generator=g)).values
    want = mom.from_temporal_graph(ei, t, n, delta=6, max_order=3)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    assert_layers_equal(pp.MultiOrderModel.from_temporal_graph(tg, delta=6, max_order=3), want)


def test_chain_runs_dry(cuda):
    """No pair continues beyond order 3: the layer above keeps its nodes and has no edges, the ones after it are empty."""
    ei = torch.tensor([[0, 1, 2, 5, 6], [1, 2, 3, 6, 7]])
    t = torch.tensor([1, 2, 3, 10, 11])
    want = mom.from_temporal_graph(ei, t, 8, delta=2, max_order=5)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), 8)
    got = pp.MultiOrderModel.from_temporal_graph(tg, delta=2, max_order=5)
    assert {k: (v.n, v.m) for k, v in got.layers.items()} == {k: (v.num_nodes, v.edge_index.size(1)) for k, v in want.items()}
    assert_layers_equal(got, want)
    with pytest.raises((RuntimeError, ValueError)):
        pp.MultiOrderModel.from_temporal_graph(tg, delta=0, max_order=2)   # no pair at all: temporal.py:53


def test_chain_equals_sort_per_order_chain(cuda, monkeypatch):
    """Mid-size stream, orders 1-4: the generation-order chain and the radix-sort chain build identical layers."""
    ei, t, g = stream(21, 20_000, 400_000, 2_000)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), 20_000)
    new = pp.MultiOrderModel.from_temporal_graph(tg, delta=60, max_order=4)
    monkeypatch.setenv("PPG_CHAIN", "0")
    old = pp.MultiOrderModel.from_temporal_graph(tg, delta=60, max_order=4)
    for k in range(1, 5):
        a, b = new.layers[k].data, old.layers[k].data
        assert a.num_nodes == b.num_nodes
        assert torch.equal(a.edge_index.as_tensor(), b.edge_index.as_tensor()), k
        assert torch.equal(a.edge_weight, b.edge_weight), k
        assert torch.equal(a.node_sequence, b.node_sequence), k
        assert torch.equal(a.inverse_idx, b.inverse_idx), k


def test_chain_bad_ids(cuda):
    ei = torch.tensor([[0, 1, 2], [1, 9, 0]], device=cuda)
    tg = pp.TemporalGraph.from_tensors(ei, torch.tensor([1, 2, 3], device=cuda), 3)
    with pytest.raises(ValueError):
        pp.MultiOrderModel.from_temporal_graph(tg, delta=2, max_order=2)


@pytest.mark.parametrize("R,rows,total,row_lo,world", [(0, 5, 9, 3, 2), (1, 1, 1, 0, 1), (4000, 60, 200, 1000, 3), (700_000, 50_000, 400_000, 123_456, 8),
                                                       (300_000, 200_000, 9_000_000, 5_000_000, 16), (50_000, 7, 30, 10, 4),
                                                       (3_000, 10_000_000, 9_000_000, 0, 5)])
def test_merge_sorted_runs_matches_torch(cuda, R, rows, total, row_lo, world):
    """Owner side of the distributed chain: `world` runs of records, each sorted by (row, col), merged in tiles; against
    the torch restatement of the owner's merge (stable: equal keys keep sender order, then arrival order).  One case
    crams all records into 7 rows: the row ranges overflow their tiles and the status word asks for the sort; the last
    one spreads few records over many rows, which takes the search-based ranking instead of the per-row counts."""
    from pathpyg_b200.parallel import _MergeSorted  # noqa
    from torch_local_ops import TorchOps

    gen = torch.Generator().manual_seed(R + world)
    cuts = torch.sort(torch.randint(0, R + 1, (world - 1,), generator=gen)).values.tolist()
    seg = [0] + cuts + [R]
    recv = [b - a for a, b in zip(seg[:-1], seg[1:])]
    src = torch.randint(row_lo, row_lo + rows, (R,), generator=gen)
    dst = torch.randint(0, total, (R,), generator=gen)
    for a, b in zip(seg[:-1], seg[1:]):   # every run sorted by (row, col)
        order = torch.argsort(src[a:b] * (total + 1) + dst[a:b], stable=True)
        src[a:b], dst[a:b] = src[a:b][order], dst[a:b][order]
    last = dst % 1000
    w = torch.rand(R, generator=gen)
    bits = w.view(torch.int32).to(torch.int64) & 0xffffffff
    records = torch.stack([(src << 32) | dst, (bits << 32) | last], dim=1).to(cuda)
    back = torch.empty(max(R, 1), dtype=torch.int32, device=cuda)
    got = _MergeSorted([records.data_ptr() + 16 * a for a in seg[:-1]], recv, [back.data_ptr() + 4 * a for a in seg[:-1]], row_lo, rows,
                       total, cuda)
    want = TorchOps.merge_records_begin(records.cpu(), row_lo, rows, total)   # on the host: index_add_ runs in arrival order there
    words = got.result_words.tolist()
    if rows == 7:
        assert words[1] & 2
        return
    assert words == want.result_words.tolist()
    assert torch.equal(back[:R].cpu(), want.inverse)
    n_out = words[0]
    for a, b in zip(got.finish(n_out), want.finish(n_out)):
        assert torch.equal(a.cpu(), b)
