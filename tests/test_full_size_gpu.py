"""BASELINE.json configurations 3 and 4 at FULL size on one GPU, checked through size-independent properties
(the reference cannot run these sizes): exact event-graph equality with the closed form, sortedness /
uniqueness of every layer, weight conservation, consistency of ``inverse_idx`` and ``node_sequence`` with the
k-grams they encode, idempotence of the aggregation, and agreement of the one-sort-per-order layer chain with the
per-order ``aggregate_edge_index``."""
import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import lift
from pathpyg_b200 import ops
from pathpyg_b200.algorithms import aggregate_edge_index, lift_order_edge_index

pytestmark = pytest.mark.gpu


def check_layer_invariants(layer, k, total_weight=None):
    d = layer.data
    ei, ns, w = d.edge_index.as_tensor(), d.node_sequence, d.edge_weight
    n = int(d.num_nodes)
    assert ns.shape == (n, k)
    if ei.size(1):
        assert int(ei.min()) >= 0 and int(ei.max()) < n
        key = ei[0] * n + ei[1]
        assert bool((key[1:] > key[:-1]).all())                       # (row, col)-sorted, no duplicate edge
    if n > 1:                                                          # rows strictly ascending in lexicographic order
        a, b = ns[:-1], ns[1:]
        diff = a != b
        first = diff.float().argmax(dim=1)
        assert bool(diff.any(dim=1).all())
        assert bool((b.gather(1, first[:, None]) > a.gather(1, first[:, None])).all())
    if k > 1 and ei.size(1):                                           # De Bruijn property: suffix of row node == prefix of col node
        assert torch.equal(ns[ei[0]][:, 1:], ns[ei[1]][:, :-1])
    if total_weight is not None:
        assert float(w.double().sum()) == float(total_weight)         # integer-valued fp32 weights: exact


def test_config3_full_size_properties(cuda):
    """cfg3: m = 10M time-stamped edges, N = 100k, T = 250, delta = 5, orders 1-3."""
    gen = torch.Generator().manual_seed(0)
    n, m, T, delta = 100_000, 10_000_000, 250, 5
    ei = torch.randint(0, n, (2, m), generator=gen)
    t = torch.sort(torch.randint(0, T, (m,), generator=gen)).values
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    event_graph = pp.algorithms.lift_order_temporal(tg, delta)
    want = lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), delta)          # numpy searchsorted, ~10 s
    assert np.array_equal(event_graph.cpu().numpy(), want)
    del want

    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=3)
    e2 = event_graph.size(1)
    e3 = lift_order_edge_index(event_graph, m)
    check_layer_invariants(model.layers[1], 1, total_weight=m)
    check_layer_invariants(model.layers[2], 2, total_weight=e2)
    check_layer_invariants(model.layers[3], 3, total_weight=e3.size(1))
    # inverse_idx of layer k names, for every level-(k-1) line-graph node, the De Bruijn node carrying its k-gram
    l2, l3 = model.layers[2].data, model.layers[3].data
    dev_ei = tg.data.edge_index.as_tensor()
    assert torch.equal(l2.node_sequence[l2.inverse_idx], dev_ei.t())
    grams3 = torch.cat([dev_ei.t()[event_graph[0]], dev_ei[1][event_graph[1]][:, None]], dim=1)
    assert torch.equal(l3.node_sequence[l3.inverse_idx], grams3)
    # the layer chain (one sort per order) against the per-order aggregation of the reference's formulation
    direct = aggregate_edge_index(e3, grams3)
    assert torch.equal(direct.data.edge_index.as_tensor(), l3.edge_index.as_tensor())
    assert torch.equal(direct.data.node_sequence, l3.node_sequence)
    assert torch.equal(direct.data.edge_weight, l3.edge_weight)
    assert torch.equal(direct.data.inverse_idx, l3.inverse_idx)
    # idempotence: aggregating an aggregated layer changes nothing
    again = aggregate_edge_index(l3.edge_index.as_tensor(), l3.node_sequence, l3.edge_weight)
    assert torch.equal(again.data.edge_index.as_tensor(), l3.edge_index.as_tensor())
    assert torch.equal(again.data.edge_weight, l3.edge_weight)
    assert torch.equal(again.data.inverse_idx, torch.arange(l3.num_nodes, device=cuda))


def test_config4_full_size_properties(cuda):
    """cfg4 on one GPU: 5M walks of 3-11 nodes over 100k nodes, order 2, DBGNN(32) training step."""
    gen = torch.Generator().manual_seed(4)
    n, P, H = 100_000, 5_000_000, 32
    lengths = torch.randint(3, 12, (P,), generator=gen)
    total = int(lengths.sum())
    flat = torch.randint(0, n, (total,), generator=gen)
    flat[:n] = torch.arange(n)                                      # every node occurs (lift_order.py:133-143)
    paths = pp.PathData(device=cuda)
    paths.append_index_walks(flat.to(cuda), lengths.to(cuda), torch.ones(P, device=cuda))
    d = paths.data
    assert d.node_sequence.size(0) == total and d.edge_index.size(1) == total - P
    model = pp.MultiOrderModel.from_path_data(paths, max_order=2)
    check_layer_invariants(model.layers[1], 1, total_weight=total - P)
    check_layer_invariants(model.layers[2], 2, total_weight=total - 2 * P)
    l1, l2 = model.layers[1].data, model.layers[2].data
    # every walk edge maps to the layer-1 edge with its end points; layer-2 nodes are exactly the layer-1 edges
    walk_nodes = d.node_sequence.reshape(-1)
    assert torch.equal(l2.node_sequence[l2.inverse_idx], walk_nodes[d.edge_index.as_subclass(torch.Tensor)].t())
    assert torch.equal(l2.node_sequence, l1.edge_index.as_tensor().t())
    # out-degree sums of layer 1 reproduce the walk-edge counts per source node
    deg = model.layers[1].degrees("out", "edge_weight", True)
    want = torch.bincount(walk_nodes[d.edge_index[0]], minlength=n).float()
    assert torch.equal(deg, want)

    model.layers[1].data.x = torch.randn(n, H, generator=gen).to(cuda)
    data = model.to_dbgnn_data(max_order=2, x_h=torch.randn(l2.num_nodes, H, generator=gen).to(cuda))
    y = torch.randint(0, 16, (n,), generator=gen).to(cuda)
    net = pp.nn.DBGNN(num_classes=16, num_features=(H, H), hidden_dims=[H, H, H]).to(cuda)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        out = net(data)
        assert out.shape == (n, 16) and bool(torch.isfinite(out).all())
        loss = torch.nn.functional.cross_entropy(out, y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0]


def test_orders_1_to_5_chain_consistency(cuda):
    """cfg5's order range on a single-GPU slice: every layer of the chain equals the per-order aggregation."""
    gen = torch.Generator().manual_seed(5)
    n, m, T, delta = 200_000, 4_000_000, 2_000, 20            # continuation factor (m/n)(delta/T) = 0.2: E_k shrinks
    ei = torch.randint(0, n, (2, m), generator=gen).to(cuda)
    t = torch.sort(torch.randint(0, T, (m,), generator=gen)).values.to(cuda)
    tg = pp.TemporalGraph.from_tensors(ei, t, n)
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=5)
    line, grams, num = pp.algorithms.lift_order_temporal(tg, delta), ei.t().contiguous(), m
    for k in range(2, 6):
        layer = model.layers[k].data
        check_layer_invariants(model.layers[k], k, total_weight=line.size(1))
        direct = aggregate_edge_index(line, grams)
        assert torch.equal(direct.data.edge_index.as_tensor(), layer.edge_index.as_tensor()), k
        assert torch.equal(direct.data.node_sequence, layer.node_sequence), k
        assert torch.equal(direct.data.inverse_idx, layer.inverse_idx), k
        if k < 5:
            nxt = lift_order_edge_index(line, num)
            grams = ops.extend_rows(grams, line)
            num, line = line.size(1), nxt
