"""Drop-in check of the containers: the reference's OWN algorithm modules (executed from /root/reference by
``oracle/ref_loader.reference_module_on``) run on ``pathpyg_b200``'s ``Graph`` / ``TemporalGraph`` / ``PathData`` and
must return what this package's functions of the same name return.  Skipped where /root/reference is not mounted."""
import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import ref_loader
from pathpyg_b200 import Graph, IndexMap
from pathpyg_b200.algorithms import centrality, components, shortest_paths

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


def random_graph(seed: int, n: int, e: int, undirected: bool = False) -> Graph:
    gen = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=gen)
    ids = IndexMap([f"v{i:02d}" for i in range(n)])
    if undirected:
        edges = [(f"v{a:02d}", f"v{b:02d}") for a, b in ei.t().tolist()]
        return Graph.from_edge_list(edges + [(w, v) for v, w in edges], is_undirected=True, mapping=ids)
    return Graph.from_edge_index(ei, mapping=ids, num_nodes=n)


@pytest.mark.parametrize("seed, n, e, undirected", [(0, 30, 25, False), (1, 40, 60, True), (2, 12, 40, False), (3, 50, 30, True)])
def test_components_and_hop_distances(seed, n, e, undirected):
    ref_c = ref_loader.reference_module_on(pp, "algorithms/components.py", "_ref_components")
    ref_s = ref_loader.reference_module_on(pp, "algorithms/shortest_paths.py", "_ref_shortest_paths")
    g = random_graph(seed, n, e, undirected)
    for connection in ("weak", "strong"):
        want, got = ref_c.connected_components(g, connection), components.connected_components(g, connection)
        assert want[0] == got[0] and np.array_equal(want[1], got[1])
        a, b = ref_c.largest_connected_component(g, connection), components.largest_connected_component(g, connection)
        assert a.n == b.n and a.m == b.m and a.mapping == b.mapping
        assert torch.equal(a.data.edge_index.as_tensor(), b.data.edge_index.as_tensor())
    want, got = ref_s.shortest_paths_dijkstra(g), shortest_paths.shortest_paths_dijkstra(g)
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    assert ref_s.diameter(g) == shortest_paths.diameter(g)
    assert ref_s.avg_path_length(g) == shortest_paths.avg_path_length(g)


@pytest.mark.parametrize("seed, n, e", [(0, 30, 120), (1, 50, 100), (2, 12, 80)])
def test_static_betweenness_and_path_statistics(seed, n, e):
    ref = ref_loader.reference_module_on(pp, "algorithms/centrality.py", "_ref_centrality")
    g = random_graph(seed, n, e)
    want, got = ref.betweenness_centrality(g), centrality.betweenness_centrality(g)
    assert set(want) == set(got)
    assert all(abs(want[k] - got[k]) <= 1e-9 * max(1.0, abs(want[k])) for k in want)
    some = g.nodes[: n // 3]
    want, got = ref.betweenness_centrality(g, sources=some), centrality.betweenness_centrality(g, sources=some)
    assert set(want) == set(got) and all(abs(want[k] - got[k]) <= 1e-9 * max(1.0, abs(want[k])) for k in want)
    gen = torch.Generator().manual_seed(seed)
    paths = pp.PathData(mapping=IndexMap([f"v{i:02d}" for i in range(n)]))
    for _ in range(20):
        walk = [f"v{i:02d}" for i in torch.randint(0, n, (int(torch.randint(2, 7, (1,), generator=gen)),), generator=gen).tolist()]
        paths.append_walk(tuple(walk))
    assert ref.path_node_traversals(paths) == centrality.path_node_traversals(paths)
    assert ref.path_visitation_probabilities(paths) == centrality.path_visitation_probabilities(paths)
    assert ref.map_to_nodes(g, {0: 1.5, 3: 2.5}) == centrality.map_to_nodes(g, {0: 1.5, 3: 2.5})


def test_rolling_time_window_unweighted():
    from test_containers import LONG

    ref = ref_loader.reference_module_on(pp, "algorithms/rolling_time_window.py", "_ref_rolling")
    tg = pp.TemporalGraph.from_edge_list(LONG)
    want = list(ref.RollingTimeWindow(tg, 10, 5, return_window=True, weighted=False))
    got = list(pp.algorithms.RollingTimeWindow(tg, 10, 5, return_window=True, weighted=False))
    assert len(want) == len(got) == 10
    for (a, wa), (b, wb) in zip(want, got):
        assert wa == wb and torch.equal(a.data.edge_index.as_tensor(), b.data.edge_index.as_tensor())
