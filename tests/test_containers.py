"""Graph / TemporalGraph container surface: the known answers of the reference's own tests
(``tests/core/test_graph.py``, ``tests/core/test_temporal_graph.py``) for the members that are index plumbing and
therefore run wherever the tensors live.  The members that merge edges run on the CUDA path only and are covered
in ``tests/test_containers_gpu.py``."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import pathpyg_b200 as pp
from pathpyg_b200 import Graph, IndexMap, TemporalGraph

LONG = [("a", "b", 1), ("b", "c", 5), ("c", "d", 9), ("c", "e", 9), ("c", "f", 11), ("f", "a", 13), ("a", "g", 18),
        ("b", "f", 21), ("a", "g", 26), ("c", "f", 27), ("h", "f", 27), ("g", "h", 28), ("a", "c", 30), ("a", "b", 31),
        ("c", "h", 32), ("f", "h", 33), ("b", "i", 42), ("i", "b", 42), ("c", "i", 47), ("h", "i", 50)]  # tests/core/conftest.py:50-74


@pytest.fixture
def simple_graph():  # tests/core/conftest.py:12-15
    return Graph.from_edge_list([("a", "b"), ("b", "c"), ("a", "c")])


@pytest.fixture
def long_temporal_graph():
    return TemporalGraph.from_edge_list(LONG)


def test_from_edge_list_id_order():  # test_graph.py:59-99
    g = Graph.from_edge_list([("a", "b"), ("c", "a"), ("b", "c")])
    assert [g.mapping.to_idx(v) for v in "abc"] == [0, 1, 2]
    assert torch.equal(g.data.edge_index, pp.EdgeIndex([[0, 1, 2], [1, 2, 0]]))
    g = Graph.from_edge_list([(1, 12), (2, 1)])
    assert [g.mapping.to_idx(v) for v in (1, 2, 12)] == [0, 1, 2]
    g = Graph.from_edge_list([("1", "12"), ("2", "1"), ("21", "3")])
    assert [g.mapping.to_idx(v) for v in ("1", "2", "3", "12", "21")] == [0, 1, 2, 3, 4]


def test_undirected_flag_and_edge_count():  # test_graph.py:102-115, graph.py:646-662
    g = Graph.from_edge_list([("a", "b"), ("b", "a"), ("b", "c"), ("c", "b"), ("c", "a"), ("a", "c")], is_undirected=True)
    assert g.is_undirected() and not g.is_directed()
    assert g.m == 3
    assert str(g).startswith("Undirected graph with 3 nodes and 3 edges")
    loops = Graph(pp.Data(edge_index=pp.EdgeIndex([[0, 0, 1], [0, 1, 0]], sparse_size=(2, 2), is_undirected=True), num_nodes=2))
    assert loops.m == 2 and loops.has_self_loops()


def test_attr_name_lists(simple_graph):  # test_graph.py:131-148
    assert simple_graph.node_attrs() == [] and simple_graph.edge_attrs() == []
    simple_graph.data["node_class"] = torch.tensor([[1], [2], [3]])
    simple_graph.data["edge_weight"] = torch.tensor([[1], [1], [2]])
    assert simple_graph.node_attrs() == ["node_class"] and simple_graph.edge_attrs() == ["edge_weight"]


def test_nodes_edges_neighbours(simple_graph):  # test_graph.py:151-189
    assert simple_graph.nodes == ["a", "b", "c"]
    assert [tuple(e) for e in simple_graph.edges] == [("a", "b"), ("a", "c"), ("b", "c")]
    assert simple_graph.successors("a") == ["b", "c"] and simple_graph.successors("c") == []
    assert simple_graph.predecessors("b") == ["a"] and simple_graph.predecessors("a") == []
    for v, w in (("a", "b"), ("a", "c"), ("b", "c")):
        assert simple_graph.is_edge(v, w) and not simple_graph.is_edge(w, v)
    plain = Graph.from_edge_index(torch.tensor([[0, 0, 1], [1, 2, 2]]))
    assert plain.successors(0) == [1, 2] and plain.predecessors(2) == [0, 1]
    assert plain.get_successors(7).numel() == 0


def test_sparse_adj_matrix(simple_graph):  # test_graph.py:192-215
    adj = simple_graph.sparse_adj_matrix()
    assert adj.shape == (3, 3) and adj.nnz == 3
    simple_graph.data["edge_weight"] = torch.tensor([[1], [1], [2]])
    weighted = simple_graph.sparse_adj_matrix("edge_weight")
    assert isinstance(weighted, sp.coo_matrix) and weighted.shape == (3, 3)
    assert weighted.data.tolist() == [1, 1, 2]
    g = Graph.from_edge_index(torch.tensor([[0], [1]]), num_nodes=5)
    assert g.sparse_adj_matrix().shape == (5, 5) and g.sparse_adj_matrix().nnz == 1
    g.data.edge_attr = torch.tensor([[1]])
    assert g.sparse_adj_matrix("edge_attr").nnz == 1


def test_laplacian(simple_graph):  # test_graph.py:266-276
    lap = simple_graph.laplacian()
    assert isinstance(lap, sp.coo_matrix) and lap.shape == (3, 3) and lap.nnz == 6
    assert lap.data.tolist() == [-1, -1, -1, 2, 1, 0]
    sym = simple_graph.laplacian(normalization="sym").toarray()
    # a -> b, a -> c, b -> c: out-degrees (2, 1, 0); D^-1/2 A D^-1/2 with inf -> 0
    want = np.eye(3)
    want[0, 1] = -1 / np.sqrt(2)
    assert np.allclose(sym, want)
    rw = simple_graph.laplacian(normalization="rw").toarray()
    assert np.allclose(rw, [[1, -0.5, -0.5], [0, 1, -1], [0, 0, 1]])
    simple_graph.data["edge_weight"] = torch.tensor([1.0, 3.0, 2.0])
    assert simple_graph.laplacian(edge_attr="edge_weight").data.tolist() == [-1, -3, -2, 4, 2, 0]


def test_add_without_ids():  # test_graph.py:279-291
    g1 = Graph.from_edge_index(torch.IntTensor([[0, 1, 1], [1, 2, 3]]), num_nodes=4)
    g2 = Graph.from_edge_index(torch.IntTensor([[0, 1, 1], [1, 2, 3]]), num_nodes=4)
    g = g1 + g2
    assert g.n == 4 and g.m == 6
    assert torch.equal(g.data.edge_index, torch.tensor([[0, 0, 1, 1, 1, 1], [1, 1, 2, 3, 2, 3]]))
    g3 = Graph.from_edge_index(torch.IntTensor([[0, 2, 3], [2, 3, 4]]), num_nodes=5)
    g = g1 + g2 + g3
    assert g.n == 5 and g.m == 9
    assert torch.equal(g.data.edge_index, torch.tensor([[0, 0, 0, 1, 1, 1, 1, 2, 3], [1, 1, 2, 2, 3, 2, 3, 3, 4]]))


@pytest.mark.parametrize("ids2, n, want", [
    (["a", "b", "c", "d"], 4, [[0, 0, 1, 1, 1, 1], [1, 1, 2, 3, 2, 3]]),    # test_graph.py:294-301
    (["e", "f", "g", "h"], 8, [[0, 1, 1, 4, 5, 5], [1, 2, 3, 5, 6, 7]]),    # :304-311
    (["a", "b", "g", "h"], 6, [[0, 0, 1, 1, 1, 1], [1, 1, 2, 3, 4, 5]]),    # :314-321
])
def test_add_with_ids(ids2, n, want):
    g1 = Graph.from_edge_index(torch.IntTensor([[0, 1, 1], [1, 2, 3]]), mapping=IndexMap(["a", "b", "c", "d"]))
    g2 = Graph.from_edge_index(torch.IntTensor([[0, 1, 1], [1, 2, 3]]), mapping=IndexMap(ids2))
    g = g1 + g2
    assert g.n == n and g.m == 6
    assert torch.equal(g.data.edge_index, torch.tensor(want))


def test_add_with_attributes():  # test_graph.py:324-357
    g1 = Graph.from_edge_index(torch.IntTensor([[0, 1, 1], [1, 2, 3]]), mapping=IndexMap(["a", "b", "c", "d"]))
    g2 = Graph.from_edge_index(torch.IntTensor([[0, 1, 1], [1, 2, 3]]), mapping=IndexMap(["a", "b", "g", "h"]))
    g1["node_class"], g2["node_class"] = torch.tensor([[1], [2], [3], [4]]), torch.tensor([[5], [6], [7], [8]])
    g1["edge_weight"], g2["edge_weight"] = torch.tensor([[1], [2], [3]]), torch.tensor([[4], [5], [6]])
    g = g1 + g2
    assert torch.equal(g["node_class"], torch.tensor([[6], [8], [3], [4], [7], [8]]))
    assert torch.equal(g["edge_weight"], torch.tensor([[1], [4], [2], [3], [5], [6]]))
    assert torch.equal(g1.__add__(g2, reduce="max")["node_class"], torch.tensor([[5], [6], [3], [4], [7], [8]]))
    assert torch.equal(g1.__add__(g2, reduce="mul")["node_class"], torch.tensor([[5], [12], [3], [4], [7], [8]]))


def test_get_and_set_attributes(simple_graph):  # test_graph.py:371-449
    simple_graph["node_class"] = torch.tensor([[1], [2], [3]])
    assert [simple_graph["node_class", v].item() for v in "abc"] == [1, 2, 3]
    simple_graph["node_class", "a"] = 42
    assert simple_graph["node_class", "a"].item() == 42
    with pytest.raises(KeyError):
        simple_graph["node_class", "d"]
    with pytest.raises(KeyError):
        simple_graph["node_class_1", "a"]
    with pytest.raises(KeyError):
        simple_graph["node_class", "d"] = 42
    with pytest.raises(KeyError):
        simple_graph["node_class_1", "a"] = 42
    simple_graph["edge_weight"] = torch.tensor([[1], [1], [2]])
    assert simple_graph["edge_weight", "a", "b"].item() == 1 and simple_graph["edge_weight", "b", "c"].item() == 2
    simple_graph["edge_weight", "a", "b"] = 42
    assert simple_graph["edge_weight", "a", "b"].item() == 42
    with pytest.raises(KeyError):
        simple_graph["edge_weight", "a", "d"]
    with pytest.raises(KeyError):
        simple_graph["edge_weight_1", "a", "b"] = 42
    with pytest.raises(ValueError):
        simple_graph["node_short"] = torch.tensor([1])
    simple_graph["graph_feature"] = torch.tensor([42])
    assert simple_graph["graph_feature"].item() == 42
    with pytest.raises(KeyError):
        simple_graph["graph_feature", "a"] = 42
    with pytest.raises(KeyError):
        simple_graph["nothing"]


def test_graph_str_lists_attributes(simple_graph):  # graph.py:772-805 (docstring of to_undirected, :222-223)
    assert str(simple_graph) == ("Directed graph with 3 nodes and 3 edges\n"
                                 "{'Edge Attributes': {}, 'Graph Attributes': {'num_nodes': \"<class 'int'>\"}, 'Node Attributes': {}}")


# ---- TemporalGraph (tests/core/test_temporal_graph.py) ----------------------------------------------------------
def test_temporal_sizes_and_span(long_temporal_graph):  # :42-54
    g = long_temporal_graph
    assert (g.n, g.m, g.start_time, g.end_time, g.order) == (9, 20, 1, 50, 1)
    assert g.temporal_edges[:2] == [("a", "b", 1), ("b", "c", 5)]
    assert str(g).startswith("Temporal Graph with 9 nodes, 17 unique edges and 20 events in [1, 50]")


def test_temporal_static_and_undirected(long_temporal_graph):  # :57-85
    g = long_temporal_graph
    s = g.to_static_graph()
    assert (s.n, s.m) == (9, 20)
    w = g.to_static_graph(time_window=(9, 12))
    assert w.m == 3
    u = g.to_undirected()
    assert (u.n, u.m) == (9, 40) and u.data.is_sorted_by_time()
    g.shuffle_time()
    assert (g.n, g.m) == (9, 20) and g.to_static_graph().m == 20
    assert not g.time_is_known_sorted()


def test_temporal_batch_and_window(long_temporal_graph):  # :88-116
    g = long_temporal_graph
    assert (g.get_batch(1, 9).n, g.get_batch(1, 9).m) == (9, 8)
    assert (g.get_batch(9, 13).n, g.get_batch(9, 13).m) == (9, 4)
    assert g.get_window(1, 10).m == 4 and g.get_window(10, 14).m == 2
    g.data.edge_tensor = torch.arange(g.m)
    g.data.edge_array = np.arange(g.m)
    b = g.get_batch(1, 9)
    assert b.data.edge_tensor.tolist() == list(range(1, 9)) and b.data.edge_array.tolist() == list(range(1, 9))
    w = g.get_window(2, 10)
    assert w.data.edge_tensor.tolist() == [1, 2, 3] and w.data.edge_array.tolist() == [1, 2, 3]
    assert w.mapping is g.mapping
    assert g["edge_tensor", "a", "b"].item() == 13 and g["edge_tensor", "a", "b", 1].item() == 0  # last event / exact event
    with pytest.raises(KeyError):
        g["edge_other", "a", "b"]


def test_temporal_ctor_orders_numpy_edge_attributes():
    d = pp.Data(edge_index=torch.tensor([[0, 1, 2], [1, 2, 0]]), time=torch.tensor([5, 1, 3]), num_nodes=3,
                edge_label=np.array(["x", "y", "z"]), edge_w=torch.tensor([1.0, 2.0, 3.0]))
    g = TemporalGraph(d)
    assert g.data.time.tolist() == [1, 3, 5] and g.data.edge_label.tolist() == ["y", "z", "x"]
    assert g.data.edge_w.tolist() == [2.0, 3.0, 1.0]


# ---- callers of the containers ----------------------------------------------------------------------------------
def test_rolling_time_window(long_temporal_graph):  # tests/algorithms/test_rolling_time_window.py:6-20
    # (unweighted snapshots: the weighted ones merge edges on the GPU; no window of this fixture repeats an edge)
    r = pp.algorithms.RollingTimeWindow(long_temporal_graph, 10, 10, False, weighted=False)
    assert [(g.n, g.m) for g in r] == [(5, 4), (7, 3), (8, 6), (8, 3), (9, 4)]
    r = pp.algorithms.RollingTimeWindow(long_temporal_graph, 10, 20, return_window=True, weighted=False)
    assert [w for _, w in r] == [(1, 11), (21, 31), (41, 51)]


def test_path_visit_statistics():  # tests/algorithms/test_centrality.py:22-42 with tests/algorithms/conftest.py:49-55
    paths = pp.PathData(mapping=IndexMap(["A", "B", "C", "D", "E", "F"]))
    for walk in (("C", "B", "D", "F"), ("A", "B", "D"), ("D", "E")):
        paths.append_walk(walk, weight=1.0)
    assert pp.algorithms.path_node_traversals(paths) == {"A": 1, "B": 2, "C": 1, "D": 3, "E": 1, "F": 1}
    prob = pp.algorithms.path_visitation_probabilities(paths)
    assert prob == {"A": 1 / 9, "B": 2 / 9, "C": 1 / 9, "D": 3 / 9, "E": 1 / 9, "F": 1 / 9}
    g = Graph.from_edge_list([("a", "b"), ("b", "c")])
    assert pp.algorithms.map_to_nodes(g, {0: 0.5, 2: 0.3}) == {"a": 0.5, "c": 0.3}
    assert pp.utils.to_numpy(g.data.edge_index).tolist() == [[0, 1], [1, 2]] and pp.utils.to_numpy([1, 2]).tolist() == [1, 2]


def test_add_higher_order_graphs():  # tests/core/test_graph.py:360-368 (layers built by hand: the lift itself needs the GPU)
    base = IndexMap(["A", "B", "C", "D", "E"])
    ns = torch.tensor([[0, 2], [1, 2], [2, 3], [2, 4]])   # order-2 nodes of the walks A-C-D, B-C-E

    def layer(weight):
        d = pp.Data(edge_index=torch.tensor([[0, 1], [2, 3]]), num_nodes=4, node_sequence=ns.clone(),
                    edge_weight=torch.tensor(weight), inverse_idx=torch.tensor([0, 2, 0, 2, 1, 3, 1, 3]))
        return Graph(d, mapping=pp.HigherOrderIndexMap(base, ns))

    g1, g2 = layer([2.0, 2.0]), layer([4.0, 4.0])
    g = g1 + g2
    assert (g.n, g.m, g.order) == (4, 4, 2)
    assert g.nodes == [("A", "C"), ("B", "C"), ("C", "D"), ("C", "E")]
    half = g1.data.inverse_idx.size(0)
    assert (g.mapping.to_ids(g.data.inverse_idx[:half]) == g1.mapping.to_ids(g1.data.inverse_idx)).all()
    assert (g.mapping.to_ids(g.data.inverse_idx[half:]) == g2.mapping.to_ids(g2.data.inverse_idx)).all()
    assert g.data.edge_weight.tolist() == [2.0, 4.0, 2.0, 4.0]
    assert g.successors(("A", "C")) == [("C", "D"), ("C", "D")]


def _both_ways(edges):
    return Graph.from_edge_list(edges + [(w, v) for v, w in edges], is_undirected=True)   # = to_undirected() without the GPU


TWO_PARTS = [("a", "b"), ("b", "c"), ("c", "a"), ("d", "e"), ("e", "f"), ("f", "g"), ("g", "d"), ("d", "f")]


def test_static_hop_distances():  # tests/algorithms/test_shortest_paths.py:8-24
    from pathpyg_b200.algorithms.shortest_paths import avg_path_length, diameter, shortest_paths_dijkstra

    g = _both_ways([("a", "b"), ("b", "c"), ("c", "e"), ("b", "d"), ("d", "e")])
    dist, pred = shortest_paths_dijkstra(g)
    assert dist.tolist() == [[0, 1, 2, 2, 3], [1, 0, 1, 1, 2], [2, 1, 0, 2, 1], [2, 1, 2, 0, 1], [3, 2, 1, 1, 0]]
    for i in range(5):
        for j in range(5):
            if i != j:
                assert dist[i, j] == dist[i, pred[i, j]] + 1
    assert diameter(g) == 3 and avg_path_length(g) == 1.6


def test_connected_components():  # tests/algorithms/test_components.py:9-95
    from pathpyg_b200.algorithms import connected_components, largest_connected_component

    count, labels = connected_components(_both_ways(TWO_PARTS))
    assert count == 2 and labels.tolist() == [0, 0, 0, 1, 1, 1, 1]
    lcc = largest_connected_component(_both_ways(TWO_PARTS))
    assert lcc.n == 4 and set(lcc.mapping.node_ids) == {"d", "e", "f", "g"} and lcc.is_undirected()
    bridged = TWO_PARTS + [("c", "d")]
    count, labels = connected_components(_both_ways(bridged))
    assert count == 1 and labels.tolist() == [0] * 7
    g = Graph.from_edge_list(bridged)
    assert connected_components(g, connection="weak")[0] == 1
    count, labels = connected_components(g, connection="strong")
    assert count == 2 and labels.tolist() == [1, 1, 1, 0, 0, 0, 0]
    assert largest_connected_component(g, connection="weak").n == 7
    strong = largest_connected_component(g, connection="strong")
    assert strong.n == 4 and set(strong.mapping.node_ids) == {"d", "e", "f", "g"}
    count, labels = connected_components(Graph.from_edge_list(TWO_PARTS), connection="weak")
    assert count == 2 and labels.tolist() == [0, 0, 0, 1, 1, 1, 1]


def _reference_betweenness(g):
    """Restatement of the reference's loop (algorithms/centrality.py:100-131) for the check below."""
    from collections import defaultdict

    bw = defaultdict(float)
    for s in range(g.n):
        order, preds, sigma, dist, queue = [], defaultdict(list), defaultdict(int), defaultdict(lambda: -1), [s]
        sigma[s], dist[s] = 1, 0
        while queue:
            v = queue.pop(0)
            order.append(v)
            for w in g.get_successors(v).tolist():
                if dist[w] < 0:
                    queue.append(w)
                    dist[w] = dist[v] + 1
                if dist[w] == dist[v] + 1:
                    sigma[w] += sigma[v]
                    preds[w].append(v)
        delta = defaultdict(float)
        while order:
            w = order.pop()
            for v in preds[w]:
                delta[v] += sigma[v] / sigma[w] * (1 + delta[w])
                bw[w] = bw[w] + delta[w]
    return bw


def test_static_centralities():  # tests/algorithms/test_centrality.py:12-19
    from pathpyg_b200.algorithms import centrality

    triangle = _both_ways([("a", "b"), ("b", "c"), ("a", "c")])
    assert centrality.betweenness_centrality(triangle) == {"a": 0.0, "b": 0.0, "c": 0.0}
    assert centrality.closeness_centrality(triangle) == {"a": 1.0, "b": 1.0, "c": 1.0}       # networkx, re-keyed by node id
    with pytest.raises(NotImplementedError):
        centrality.closeness_centrality(TemporalGraph.from_edge_list(LONG))
    for seed, (n, e) in enumerate([(30, 120), (50, 100), (12, 80)]):     # multigraphs with self-loops
        gen = torch.Generator().manual_seed(seed)
        g = Graph.from_edge_index(torch.randint(0, n, (2, e), generator=gen), num_nodes=n)
        want, got = _reference_betweenness(g), centrality.betweenness_centrality(g)
        assert set(want) == set(got)
        assert all(abs(want[k] - got[k]) <= 1e-9 * max(1.0, abs(want[k])) for k in want)
    line = Graph.from_edge_list([("a", "b"), ("b", "c"), ("c", "d")])
    assert centrality.betweenness_centrality(line, sources=["a"]) == {"b": 2.0, "c": 1.0, "d": 0.0}


def test_weisfeiler_leman():  # tests/algorithms/test_wl.py:7-66
    from pathpyg_b200.algorithms import WeisfeilerLeman_test

    same, c1, c2 = WeisfeilerLeman_test(Graph.from_edge_list([("a", "b"), ("b", "c")]), Graph.from_edge_list([("y", "z"), ("x", "y")]))
    assert same is True and c1 == c2
    same, c1, c2 = WeisfeilerLeman_test(Graph.from_edge_list([("a", "b"), ("b", "c")]), Graph.from_edge_list([("y", "z"), ("x", "z")]))
    assert same is False and c1 != c2
    cube_a = _both_ways([("a", "g"), ("a", "h"), ("a", "i"), ("b", "g"), ("b", "h"), ("b", "j"), ("c", "g"), ("c", "i"), ("c", "j"),
                         ("d", "h"), ("d", "i"), ("d", "j")])
    cube_b = _both_ways([("1", "2"), ("1", "5"), ("1", "4"), ("2", "6"), ("2", "3"), ("3", "7"), ("3", "4"), ("4", "8"), ("5", "6"),
                         ("6", "7"), ("7", "8"), ("8", "5")])
    same, c1, c2 = WeisfeilerLeman_test(cube_a, cube_b)
    assert same is True and c1 == c2
    with pytest.raises(Exception, match="must not overlap"):
        WeisfeilerLeman_test(cube_a, cube_a)


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(), reason="/root/reference not mounted")
def test_weisfeiler_leman_matches_reference_source():
    """The reference's own weisfeiler_leman.py, executed on these containers, returns the same triple."""
    import importlib.util
    import sys
    import types

    from oracle import ref_loader
    from pathpyg_b200.algorithms import WeisfeilerLeman_test

    stubs = {"pathpyG": types.ModuleType("pathpyG"), "pathpyG.core": types.ModuleType("pathpyG.core"),
             "pathpyG.core.graph": types.ModuleType("pathpyG.core.graph")}
    stubs["pathpyG.core.graph"].Graph = Graph
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_wl", ref_loader.REFERENCE_ROOT + "/src/pathpyG/algorithms/weisfeiler_leman.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for seed in range(12):
        gen = torch.Generator().manual_seed(seed)
        e1 = torch.randint(0, 8, (2, 14), generator=gen)
        e2 = torch.randperm(8, generator=gen)[e1] if seed % 2 == 0 else torch.randint(0, 8, (2, 14), generator=gen)
        g1 = Graph.from_edge_index(e1, mapping=IndexMap([f"a{i}" for i in range(8)]), num_nodes=8)
        g2 = Graph.from_edge_index(e2, mapping=IndexMap([f"b{i}" for i in range(8)]), num_nodes=8)
        assert ref.WeisfeilerLeman_test(g1, g2) == WeisfeilerLeman_test(g1, g2)
