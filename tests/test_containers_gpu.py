"""Edge-merging members of the containers on the CUDA path (library coalesce), against the golden vectors of the
reference's own method bodies, the oracle on random inputs, and the reference's known answers
(tests/core/test_graph.py:118-128,218-263; tests/core/test_temporal_graph.py:68-79)."""
import os

import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import containers
from pathpyg_b200 import Graph, TemporalGraph

from test_containers import LONG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "container_golden.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def test_known_answers(cuda):
    g = Graph.from_edge_list([("a", "b"), ("b", "c"), ("a", "c")])
    g_u = g.to_undirected()
    assert g_u.data.edge_index.is_undirected and g_u.is_undirected()
    assert g_u.data.edge_index.as_tensor().tolist() == [[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]]
    assert g_u.m == 3 and not g_u.data.edge_index.is_cuda     # results follow the device of the inputs
    multi = Graph.from_edge_list([("a", "b"), ("b", "c"), ("a", "c"), ("a", "b")])
    assert multi.m == 4
    wg = multi.to_weighted_graph()
    assert wg.data.num_edges == 3 and wg.data.num_nodes == 3 and wg["edge_weight", "a", "b"] == 2
    assert g.in_degrees == {"a": 0, "b": 1, "c": 2} and g.out_degrees == {"a": 2, "b": 1, "c": 0}
    assert g.degrees(mode="out", return_tensor=True).equal(torch.tensor([2, 1, 0]))
    tg = TemporalGraph.from_edge_list(LONG)
    s = tg.to_static_graph(weighted=True)
    assert s.n == tg.n
    # a->b twice, a->c once.  The reference's test reads these as edge_weight[2] / edge_weight[0]: positions that
    # come out of an UNSTABLE row sort of 17 edges (EdgeIndex.sort_by -> Tensor.sort(stable=False) -> std::sort);
    # the row sort here is stable, so the merged edges stay (row, col)-ordered and are addressed by their ids instead.
    assert s["edge_weight", "a", "b"].item() == 2.0 and s["edge_weight", "a", "c"].item() == 1.0
    assert s.data.edge_index.as_tensor()[:, :3].tolist() == [[0, 0, 0], [1, 2, 6]]
    assert s.data.edge_weight[:3].tolist() == [2.0, 1.0, 2.0] and s.m == 17


@pytest.mark.parametrize("i", range(4))
@pytest.mark.parametrize("on_device", [False, True])
def test_graph_members_golden(cuda, gold, i, on_device):
    ei, w, n = gold[f"g{i}_edge_index"], gold[f"g{i}_edge_weight"], int(gold[f"g{i}_num_nodes"])
    dev = cuda if on_device else torch.device("cpu")
    g = Graph(pp.Data(edge_index=ei.clone().to(dev), edge_weight=w.clone().to(dev), num_nodes=n))
    u = g.to_undirected()
    assert u.data.edge_index.device.type == dev.type
    assert torch.equal(u.data.edge_index.as_tensor().cpu(), gold[f"g{i}_undirected_edge_index"])
    assert torch.equal(u.data.edge_weight.cpu(), gold[f"g{i}_undirected_edge_weight"])
    assert u.n == n and u.is_undirected()
    wg = Graph(pp.Data(edge_index=ei.clone().to(dev), num_nodes=n)).to_weighted_graph()
    assert torch.equal(wg.data.edge_index.as_tensor().cpu(), gold[f"g{i}_weighted_edge_index"])
    assert torch.equal(wg.data.edge_weight.cpu(), gold[f"g{i}_weighted_edge_weight"])


@pytest.mark.parametrize("i", range(3))
def test_to_static_graph_golden(cuda, gold, i):
    ei, t, window = gold[f"t{i}_edge_index"], gold[f"t{i}_time"], tuple(gold[f"t{i}_window"].tolist())
    n = int(ei.max()) + 1
    tg = TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    for tag, kw in (("plain", {}), ("weighted", {"weighted": True}), ("window", {"weighted": True, "time_window": window})):
        s = tg.to_static_graph(**kw)
        assert s.data.edge_index.is_cuda
        assert torch.equal(s.data.edge_index.as_tensor().cpu(), gold[f"t{i}_{tag}_edge_index"]), tag
        if kw.get("weighted"):
            assert torch.equal(s.data.edge_weight.cpu(), gold[f"t{i}_{tag}_edge_weight"]), tag


@pytest.mark.parametrize("n, e, seed", [(1000, 50_000, 1), (100_000, 2_000_000, 2), (3, 500, 3)])
def test_members_random_vs_oracle(cuda, n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = ei[:, torch.sort(ei[0], stable=True).indices]
    label = torch.randint(0, 1 << 40, (e,), generator=g)          # int64 attribute: selection by edge number is exact
    graph = Graph(pp.Data(edge_index=ei.to(cuda), edge_label=label.to(cuda), num_nodes=n))
    u = graph.to_undirected()
    want_ei, want_label, _ = containers.graph_to_undirected(ei, n, label)
    assert torch.equal(u.data.edge_index.as_tensor().cpu(), want_ei) and torch.equal(u.data.edge_label.cpu(), want_label)
    # symmetric by construction, idempotent
    uu = u.to_undirected()
    assert torch.equal(uu.data.edge_index.as_tensor(), u.data.edge_index.as_tensor())
    wg = graph.to_weighted_graph()
    want_ei, want_w = containers.graph_to_weighted(ei, n)
    assert torch.equal(wg.data.edge_index.as_tensor().cpu(), want_ei) and torch.equal(wg.data.edge_weight.cpu(), want_w)
    assert float(wg.data.edge_weight.sum()) == e


def test_window_views_on_device(cuda):
    tg = TemporalGraph.from_edge_list(LONG, device=cuda)
    assert tg.get_window(1, 10).m == 4 and tg.get_window(10, 14).m == 2 and tg.get_batch(9, 13).m == 4
    assert tg.get_window(1, 10).data.edge_index.is_cuda
    u = tg.to_undirected()
    assert (u.n, u.m) == (9, 40) and u.data.is_sorted_by_time()
    assert tg.to_static_graph(weighted=True, time_window=(9, 12)).data.edge_weight.tolist() == [1.0, 1.0, 1.0]


def test_df_to_graph_undirected(cuda):  # reference tests/io/test_pandas.py:203-210
    import pandas as pd

    from pathpyg_b200.io import df_to_graph

    df = pd.DataFrame({"v": ["a", "b", "c"], "w": ["b", "c", "a"], "edge_weight": ["a", "b", "c"]})
    g = df_to_graph(df, is_undirected=True)
    assert (g.n, g.m) == (3, 3) and g.is_undirected()
    assert g.data.edge_index.as_tensor().tolist() == [[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]]
    assert g.data.edge_weight.tolist() == ["a", "c", "a", "b", "c", "b"]     # text attributes (numpy) follow the edges
    on_gpu = df_to_graph(pd.DataFrame({"v": [0, 1, 2, 0], "w": [1, 2, 0, 1], "x": [1.0, 2.0, 3.0, 4.0]}), multiedges=True,
                         is_undirected=True, device=cuda)
    assert on_gpu.data.edge_index.is_cuda and on_gpu.m == 3
    assert on_gpu.data.edge_x.tolist() == [1.0, 3.0, 1.0, 2.0, 3.0, 2.0]      # merged (0,1) pair keeps the attribute of its first edge
