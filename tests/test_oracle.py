"""The oracle against (1) every known-answer vector the reference's tests hold for the path and
(2) the golden vectors produced by the reference's own source (tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest
import torch

from oracle import dbgnn as odbgnn
from oracle import lift, mom


# ---- reference known answers ---------------------------------------------------------------
def test_aggregate_node_attributes_known_answer():  # tests/algorithms/test_lift_order.py:12-31
    ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 0, 3, 0]])
    a = torch.tensor([1, 2, 3, 4])
    want = {"src": [1, 2, 3, 3, 4], "dst": [2, 3, 1, 4, 1], "max": [2, 3, 3, 4, 4], "mul": [2, 6, 3, 12, 4], "add": [3, 5, 4, 7, 5]}
    for rule, v in want.items():
        assert lift.aggregate_node_attributes(ei, a, rule).tolist() == v
    with pytest.raises(ValueError):
        lift.aggregate_node_attributes(ei, a, "unknown")


def test_lift_order_edge_index_known_answer():  # tests/algorithms/test_lift_order.py:34-57
    ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 0, 3, 0]])
    assert lift.lift_order_edge_index(ei, 4).tolist() == [[0, 1, 1, 2, 3, 4], [1, 2, 3, 0, 4, 0]]
    ho, w = lift.lift_order_edge_index_weighted(ei, torch.tensor([1, 2, 3, 4, 5]), 4)
    assert w.tolist() == [1, 2, 2, 3, 4, 5]


def test_aggregate_edge_index_known_answer():  # tests/algorithms/test_lift_order.py:60-79
    g = lift.aggregate_edge_index(torch.tensor([[0, 2, 2, 1], [1, 1, 3, 0]]), torch.tensor([[1, 2], [2, 3], [1, 2], [4, 5]]),
                                  torch.tensor([1, 2, 3, 4]))
    assert g.edge_index.tolist() == [[0, 0, 1], [1, 2, 0]]
    assert g.edge_weight.tolist() == [3, 3, 4]
    assert g.node_sequence.tolist() == [[1, 2], [2, 3], [4, 5]]


SIMPLE_TEMPORAL = (torch.tensor([[0, 1, 2, 2], [1, 2, 3, 4]]), torch.tensor([1, 5, 9, 9]), 5)  # tests/core/conftest.py:43-47


def test_lift_order_temporal_known_answer():  # tests/algorithms/test_temporal.py:11-17
    ei, t, n = SIMPLE_TEMPORAL
    assert lift.lift_order_temporal(ei, t, 5).tolist() == [[0, 1, 1], [1, 2, 3]]
    assert lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), 5).tolist() == [[0, 1, 1], [1, 2, 3]]
    # torch.cat([]) raises RuntimeError up to torch 2.8 (the reference's pin) and ValueError in newer releases
    with pytest.raises((RuntimeError, ValueError)):
        lift.lift_order_temporal(ei, t, 1)
    with pytest.raises((RuntimeError, ValueError)):
        lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), 1)


def test_iterate_lift_order_known_answer():  # tests/core/test_multi_order_model.py:29-42
    # edges a-b, b-c, a-c, a-b (tests/core/conftest.py:18-21) after Graph's stable row sort
    ei = torch.tensor([[0, 0, 0, 1], [1, 2, 1, 2]])
    ho, ns, w, gk = mom.iterate_lift_order(ei, torch.arange(3).unsqueeze(1))
    assert ho.tolist() == [[0, 2], [3, 3]]
    assert ns.tolist() == [[0, 1], [0, 2], [0, 1], [1, 2]]
    assert w is None
    assert gk.edge_index.tolist() == [[0], [2]]
    assert gk.node_sequence.tolist() == [[0, 1], [0, 2], [1, 2]]
    assert gk.edge_weight.tolist() == [2.0]


def test_from_path_data_known_answer():  # tests/core/test_multi_order_model.py:165-173
    walks = mom.append_walks([(0, 2, 3), (1, 2, 4)], [2.0, 2.0])
    layers = mom.from_path_data(walks, max_order=2)
    assert layers[1].edge_index.tolist() == [[0, 1, 2, 2], [2, 2, 3, 4]]
    assert layers[1].edge_weight.tolist() == [2.0, 2.0, 2.0, 2.0]
    assert layers[2].edge_index.tolist() == [[0, 1], [2, 3]]
    assert layers[2].edge_weight.tolist() == [2.0, 2.0]


def test_from_temporal_graph_known_answer():  # tests/core/test_multi_order_model.py:176-190
    ei, t, n = SIMPLE_TEMPORAL
    layers = mom.from_temporal_graph(ei, t, n, delta=4, max_order=3)
    assert layers[1].edge_index.tolist() == [[0, 1, 2, 2], [1, 2, 3, 4]]
    assert layers[2].edge_index.tolist() == [[0, 1, 1], [1, 2, 3]]
    assert layers[3].edge_index.tolist() == [[0, 0], [1, 2]]
    data = mom.to_dbgnn_data(layers, max_order=3)
    assert data["edge_index_higher_order"].tolist() == [[0, 0], [1, 2]]


def test_trp_tutorial_known_answer():
    """docs/tutorial/trp_higher_order.ipynb:67 (10-event toy, delta=1): layer sizes (4,6), (6,6) with
    weights [2,1,1,1,1,1], (6,4), (4,2), (2,0)  (:712, :1252-1256, :1796, :2338, :2887)."""
    tedges = [("a", "b", 1), ("a", "b", 2), ("b", "a", 3), ("b", "c", 3), ("d", "c", 4), ("a", "b", 4), ("c", "b", 4),
              ("c", "d", 5), ("b", "a", 5), ("c", "b", 6)]
    ids = {"a": 0, "b": 1, "c": 2, "d": 3}
    ei = torch.tensor([[ids[s] for s, _, _ in tedges], [ids[d] for _, d, _ in tedges]])
    t = torch.tensor([x for _, _, x in tedges])
    layers = mom.from_temporal_graph(ei, t, 4, delta=1, max_order=5)
    sizes = {k: (v.num_nodes, v.edge_index.size(1)) for k, v in layers.items()}
    assert sizes == {1: (4, 6), 2: (6, 6), 3: (6, 4), 4: (4, 2), 5: (2, 0)}
    assert sorted(layers[2].edge_weight.tolist(), reverse=True) == [2.0, 1.0, 1.0, 1.0, 1.0, 1.0]


def test_bipartite_known_answer():  # tests/nn/test_dbgnn.py:11-30
    walks = mom.append_walks([(0, 2, 3), (1, 2, 4)], [2.0, 2.0])
    layers = mom.from_path_data(walks, max_order=2)
    assert mom.generate_bipartite_edge_index(layers[2].node_sequence, "last").tolist() == [[0, 1, 2, 3], [2, 2, 3, 4]]
    assert mom.generate_bipartite_edge_index(layers[2].node_sequence, "first").tolist() == [[0, 1, 2, 3], [0, 1, 2, 2]]


def test_path_data_known_answer():  # tests/core/test_path_data.py:57-74
    w = mom.append_walks([(0, 1, 3), (0, 1), (2, 1, 3), (2, 1, 4)], [1.0] * 4)
    assert w.dag_num_nodes.tolist() == [3, 2, 3, 3]
    assert w.dag_num_edges.tolist() == [2, 1, 2, 2]
    assert w.edge_index.tolist() == [[0, 1, 3, 5, 6, 8, 9], [1, 2, 4, 6, 7, 9, 10]]


# ---- golden vectors from the reference's own source ----------------------------------------
@pytest.mark.parametrize("case", [f"lift{i}" for i in range(4)])
def test_lift_golden(golden, case):
    ei = torch.from_numpy(golden[f"{case}_edge_index"])
    n = int(golden[f"{case}_num_nodes"])
    assert np.array_equal(lift.lift_order_edge_index(ei, n).numpy(), golden[f"{case}_out"])
    w = torch.from_numpy(golden[f"{case}_weight"])
    for rule in ("src", "dst", "max", "mul", "add"):
        _, hw = lift.lift_order_edge_index_weighted(ei, w, n, rule)
        assert np.array_equal(hw.numpy(), golden[f"{case}_w_{rule}"])


@pytest.mark.parametrize("case", [f"agg{i}" for i in range(4)])
def test_aggregate_golden(golden, case):
    ei = torch.from_numpy(golden[f"{case}_edge_index"])
    ns = torch.from_numpy(golden[f"{case}_node_sequence"])
    w = torch.from_numpy(golden[f"{case}_weight"])
    g = lift.aggregate_edge_index(ei, ns, w)
    assert np.array_equal(g.edge_index.numpy(), golden[f"{case}_out_edge_index"])
    assert np.array_equal(g.edge_weight.numpy(), golden[f"{case}_out_weight"])
    assert np.array_equal(g.node_sequence.numpy(), golden[f"{case}_out_node_sequence"])
    assert np.array_equal(g.inverse_idx.numpy(), golden[f"{case}_out_inverse"])
    u, inv = lift.unique_rows_closed_form(ns.numpy())
    assert np.array_equal(u, golden[f"{case}_out_node_sequence"]) and np.array_equal(inv, golden[f"{case}_out_inverse"])


@pytest.mark.parametrize("case", [f"temp{i}" for i in range(6)])
def test_temporal_golden(golden, case):
    ei = torch.from_numpy(golden[f"{case}_edge_index"])
    t = torch.from_numpy(golden[f"{case}_time"])
    delta = golden[f"{case}_delta"].item()
    want = golden[f"{case}_out"]
    assert np.array_equal(lift.lift_order_temporal(ei, t, delta).numpy(), want)
    if not (t.dtype == torch.int64 and isinstance(delta, float)):  # closed form covers the exact promotions only
        assert np.array_equal(lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), delta), want)


# ---- DBGNN restatement: internal consistency (PARITY UNPINNED, see oracle/__init__.py) -------
def test_dbgnn_oracle_matches_dense_algebra():
    walks = mom.append_walks([(0, 2, 3), (1, 2, 4), (0, 2, 4), (3, 3, 2)], [2.0, 1.0, 3.0, 1.0])
    layers = mom.from_path_data(walks, max_order=2)
    data = mom.to_dbgnn_data(layers, max_order=2)
    params = odbgnn.init_params(3, (data["num_nodes"], data["num_ho_nodes"]), [16, 32, 8], seed=1, dtype=torch.float64)
    data = {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in data.items()}
    out = odbgnn.dbgnn_forward(params, data)
    assert out.shape == (data["num_nodes"], 3) and torch.isfinite(out).all()

    def dense_gcn(x, ei, w, W, b):
        n = x.size(0)
        A = torch.zeros(n, n, dtype=torch.float64)
        loops = torch.ones(n, dtype=torch.float64)
        for (r, c), v in zip(ei.t().tolist(), w.tolist()):
            if r == c:
                loops[r] = v
            else:
                A[c, r] += v
        A = A + torch.diag(loops)
        d = A.sum(1).pow(-0.5)
        return (d[:, None] * A * d[None, :]) @ (x @ W.t()) + b

    x = torch.nn.functional.elu(dense_gcn(data["x"], data["edge_index"], data["edge_weights"],
                                          params["first_order_layers.0.lin.weight"], params["first_order_layers.0.bias"]))
    got = torch.nn.functional.elu(odbgnn.gcn_conv(data["x"], data["edge_index"], data["edge_weights"],
                                                  params["first_order_layers.0.lin.weight"], params["first_order_layers.0.bias"]))
    assert torch.allclose(x, got, rtol=1e-12, atol=1e-12)
