"""MultiOrderModel builders on the GPU vs the oracle and the reference's known answers."""
import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import lift, mom

pytestmark = pytest.mark.gpu


def assert_layers_equal(model, want):
    assert sorted(model.layers) == sorted(want)
    for k, layer in want.items():
        got = model.layers[k].data
        assert torch.equal(got.edge_index.as_tensor().cpu(), layer.edge_index), k
        assert torch.equal(got.edge_weight.cpu(), layer.edge_weight), k
        assert torch.equal(got.node_sequence.cpu(), layer.node_sequence), k
        assert torch.equal(got.inverse_idx.cpu(), layer.inverse_idx), k
        assert got.num_nodes == layer.num_nodes, k


def test_from_temporal_graph_known_answer(cuda):  # reference tests/core/test_multi_order_model.py:176-190
    g = pp.TemporalGraph.from_edge_list([("a", "b", 1), ("b", "c", 5), ("c", "d", 9), ("c", "e", 9)]).to(cuda)
    m = pp.MultiOrderModel.from_temporal_graph(g, max_order=3, delta=4)
    assert torch.equal(m.layers[1].data.edge_index, pp.EdgeIndex([[0, 1, 2, 2], [1, 2, 3, 4]]).to(cuda))
    assert torch.equal(m.layers[2].data.edge_index, pp.EdgeIndex([[0, 1, 1], [1, 2, 3]]).to(cuda))
    assert torch.equal(m.layers[3].data.edge_index, pp.EdgeIndex([[0, 0], [1, 2]]).to(cuda))
    data = m.to_dbgnn_data(max_order=3)
    assert torch.equal(data.edge_index, pp.EdgeIndex([[0, 1, 2, 2], [1, 2, 3, 4]]).to(cuda))
    assert torch.equal(data.edge_index_higher_order, pp.EdgeIndex([[0, 0], [1, 2]]).to(cuda))
    assert m.layers[2].mapping.to_id(1) == ("b", "c") and m.layers[3].mapping.to_idx(("a", "b", "c")) == 0
    assert str(m) == "MultiOrderModel with max. order 3"
    with pytest.raises(ValueError):
        m.to_dbgnn_data(max_order=7)


def test_trp_tutorial_known_answer(cuda):  # docs/tutorial/trp_higher_order.ipynb:67,712,1252-1256,1796,2338,2887
    tedges = [("a", "b", 1), ("a", "b", 2), ("b", "a", 3), ("b", "c", 3), ("d", "c", 4), ("a", "b", 4), ("c", "b", 4),
              ("c", "d", 5), ("b", "a", 5), ("c", "b", 6)]
    g = pp.TemporalGraph.from_edge_list(tedges)  # host tensors: staged to the GPU and back
    m = pp.MultiOrderModel.from_temporal_graph(g, delta=1, max_order=5)
    assert {k: (v.n, v.m) for k, v in m.layers.items()} == {1: (4, 6), 2: (6, 6), 3: (6, 4), 4: (4, 2), 5: (2, 0)}
    assert sorted(m.layers[2].data.edge_weight.tolist(), reverse=True) == [2.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    assert not m.layers[2].data.edge_index.is_cuda


def test_iterate_lift_order_known_answer(cuda):  # reference tests/core/test_multi_order_model.py:29-42
    g = pp.Graph.from_edge_list([("a", "b"), ("b", "c"), ("a", "c"), ("a", "b")]).to(cuda)
    ho, ns, w, gk = pp.MultiOrderModel.iterate_lift_order(g.data.edge_index, torch.arange(g.n, device=cuda).unsqueeze(1),
                                                          mapping=g.mapping, save=True)
    assert ho.tolist() == [[0, 2], [3, 3]]
    assert ns.tolist() == [[0, 1], [0, 2], [0, 1], [1, 2]]
    assert w is None
    assert gk.data.edge_index.as_tensor().tolist() == [[0], [2]]
    assert gk.data.node_sequence.tolist() == [[0, 1], [0, 2], [1, 2]]
    assert gk.data.edge_weight.tolist() == [2.0]
    assert gk.order == 2


def test_from_path_data_known_answer(cuda):  # reference tests/core/test_multi_order_model.py:165-173, tests/nn/test_dbgnn.py:11-30
    p = pp.PathData(pp.IndexMap(["A", "B", "C", "D", "E"]))
    p.append_walk(("A", "C", "D"), weight=2.0)
    p.append_walk(("B", "C", "E"), weight=2.0)
    m = pp.MultiOrderModel.from_path_data(p.to(cuda), max_order=2)
    g1, g2 = m.layers[1], m.layers[2]
    assert torch.equal(g1.data.edge_index, pp.EdgeIndex([[0, 1, 2, 2], [2, 2, 3, 4]]).to(cuda))
    assert torch.equal(g1.data.edge_weight.cpu(), torch.tensor([2.0, 2.0, 2.0, 2.0]))
    assert torch.equal(g2.data.edge_index, pp.EdgeIndex([[0, 1], [2, 3]]).to(cuda))
    assert torch.equal(g2.data.edge_weight.cpu(), torch.tensor([2.0, 2.0]))
    assert pp.utils.generate_bipartite_edge_index(g1, g2, mapping="last").tolist() == [[0, 1, 2, 3], [2, 2, 3, 4]]
    assert pp.utils.generate_bipartite_edge_index(g1, g2, mapping="first").tolist() == [[0, 1, 2, 3], [0, 1, 2, 2]]
    assert pp.utils.generate_bipartite_edge_index(g1, g2, mapping="both").shape == (2, 8)


@pytest.mark.parametrize("seed,n,m,horizon,delta,K", [(0, 20, 100, 50, 2, 2), (1, 30, 400, 40, 3, 4), (2, 200, 6000, 300, 4, 3), (3, 12, 300, 25, 2, 5)])
def test_from_temporal_graph_vs_oracle(cuda, seed, n, m, horizon, delta, K):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=g)).values
    want = mom.from_temporal_graph(ei, t, n, delta=delta, max_order=K)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    assert_layers_equal(pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=K), want)
    # cached=False keeps only the top layer (multi_order_model.py:158,173,188-191)
    top = pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=K, cached=False)
    assert list(top.layers) == [K]
    assert torch.equal(top.layers[K].data.edge_index.as_tensor().cpu(), want[K].edge_index)
    # integer edge weights from data.edge_weight, precomputed event graph
    w = torch.randint(1, 4, (m,), generator=g).float()
    want_w = mom.from_temporal_graph(ei, t, n, delta=delta, max_order=min(K, 3), edge_weight=w)
    tgw = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n, edge_weight=w.to(cuda))
    ev = pp.algorithms.lift_order_temporal(tgw, delta)
    assert_layers_equal(pp.MultiOrderModel.from_temporal_graph(tgw, delta=delta, max_order=min(K, 3), event_graph=ev), want_w)


def test_from_temporal_graph_config2_size_vs_oracle(cuda):
    """BASELINE config 2 at full size (m = 1M, N = 100k, T = 1000, delta = 200, K = 2) against the closed-form oracle."""
    g = torch.Generator().manual_seed(0)
    n, m = 100_000, 1_000_000
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, 1000, (m,), generator=g)).values
    closed = lambda e, tt, d: torch.from_numpy(lift.lift_order_temporal_closed_form(e.numpy(), tt.numpy(), d))
    want = mom.from_temporal_graph(ei, t, n, delta=200, max_order=2, temporal_fn=closed)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    assert_layers_equal(pp.MultiOrderModel.from_temporal_graph(tg, delta=200, max_order=2), want)


@pytest.mark.parametrize("mode", ["propagation", "diffusion"])
def test_from_path_data_vs_oracle(cuda, mode):
    g = torch.Generator().manual_seed(7)
    n_nodes, walks = 30, 400
    lengths = torch.randint(1, 9, (walks,), generator=g)
    seqs = [torch.randint(0, n_nodes, (int(l),), generator=g).tolist() for l in lengths]
    seqs.append(list(range(n_nodes)))  # every node occurs (lift_order.py:133-143: num_nodes = #distinct nodes present)
    weights = torch.randint(1, 4, (len(seqs),), generator=g).float().tolist()
    want = mom.from_path_data(mom.append_walks(seqs, weights), max_order=4, mode=mode)
    p = pp.PathData()
    p.append_index_walks(torch.tensor([v for s in seqs for v in s]), torch.tensor([len(s) for s in seqs]), torch.tensor(weights))
    got = pp.MultiOrderModel.from_path_data(p.to(cuda), max_order=4, mode=mode)
    for k, layer in want.items():
        d = got.layers[k].data
        assert torch.equal(d.edge_index.as_tensor().cpu(), layer.edge_index), k
        assert torch.equal(d.node_sequence.cpu(), layer.node_sequence), k
        if mode == "propagation":
            assert torch.equal(d.edge_weight.cpu(), layer.edge_weight), k
        else:
            assert torch.allclose(d.edge_weight.cpu(), layer.edge_weight, rtol=1e-5), k


def test_from_temporal_graph_staged_from_host(cuda):
    """Host graph + ``device=``: edge index and time stamps are uploaded by the call (time stamps on a copy stream
    behind the layer-1 sort), the layers stay on the GPU and equal the device-resident build."""
    g = torch.Generator().manual_seed(12)
    n, m = 5000, 200_000
    ei = torch.randint(0, n, (2, m), generator=g).pin_memory()
    t = torch.sort(torch.randint(0, 600, (m,), generator=g)).values.pin_memory()
    w = torch.randint(1, 4, (m,), generator=g).float()
    host = pp.TemporalGraph.from_tensors(ei, t, n, edge_weight=w)
    want = pp.MultiOrderModel.from_temporal_graph(pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n, edge_weight=w.to(cuda)),
                                                  delta=6, max_order=3)
    for _ in range(3):
        got = pp.MultiOrderModel.from_temporal_graph(host, delta=6, max_order=3, device=cuda)
        for k in (1, 2, 3):
            a, b = got.layers[k].data, want.layers[k].data
            assert a.edge_index.is_cuda
            assert torch.equal(a.edge_index.as_tensor(), b.edge_index.as_tensor()) and torch.equal(a.edge_weight, b.edge_weight)
            assert torch.equal(a.node_sequence, b.node_sequence) and torch.equal(a.inverse_idx, b.inverse_idx)
    back = pp.MultiOrderModel.from_temporal_graph(host, delta=6, max_order=2)      # without device=: results return to the host
    assert not back.layers[2].data.edge_index.is_cuda
    assert torch.equal(back.layers[2].data.edge_index.as_tensor(), want.layers[2].data.edge_index.as_tensor().cpu())
