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


def _same_graph(a, b):
    assert a.n == b.n and a.m == b.m and a.mapping == b.mapping
    assert torch.equal(a.data.edge_index.as_tensor(), b.data.edge_index.as_tensor())
    assert sorted(a.edge_attrs()) == sorted(b.edge_attrs()) and sorted(a.node_attrs()) == sorted(b.node_attrs())
    for k in a.edge_attrs() + a.node_attrs():
        x, y = a.data[k], b.data[k]
        assert type(x) is type(y)
        assert torch.equal(x, y) if isinstance(x, torch.Tensor) else np.array_equal(x, y)


@pytest.mark.parametrize("strings, multi", [(True, False), (False, True), (True, True)])
def test_reference_io_module_on_containers(strings, multi, tmp_path):
    """The reference's io/pandas.py building THIS package's Graph / TemporalGraph, against this package's io functions."""
    import pandas as pd

    ref = ref_loader.reference_module_on(pp, "io/pandas.py", "_ref_io_on_containers")
    rng = np.random.default_rng(11)
    v, w = rng.integers(0, 25, 300), rng.integers(0, 25, 300)
    if strings:
        v, w = np.array([f"n{x:02d}" for x in v]), np.array([f"n{x:02d}" for x in w])
    df = pd.DataFrame({"v": v, "w": w, "weight": rng.integers(1, 9, 300).astype(float), "edge_tag": rng.integers(0, 5, 300),
                       "label": [f"l{x}" for x in rng.integers(0, 4, 300)], "vec": [str([int(x), int(x) + 1]) for x in rng.integers(0, 9, 300)]})
    want, got = ref.df_to_graph(df.copy(), multiedges=multi), pp.io.df_to_graph(df.copy(), multiedges=multi)
    _same_graph(want, got)
    assert ref.graph_to_df(want).equals(pp.io.graph_to_df(got))
    assert ref.graph_to_df(want, node_indices=True).equals(pp.io.graph_to_df(got, node_indices=True))
    nodes = pd.DataFrame({"v": list(got.nodes), "score": rng.random(got.n), "node_kind": rng.integers(0, 3, got.n)}).sample(frac=1.0, random_state=1)
    ref.add_node_attributes(nodes.copy(), want)
    pp.io.add_node_attributes(nodes.copy(), got)
    _same_graph(want, got)
    if not multi:
        edges = ref.graph_to_df(want)[["v", "w"]].copy()
        edges["flow"] = rng.random(len(edges))
        ref.add_edge_attributes(edges.copy(), want)
        pp.io.add_edge_attributes(edges.copy(), got)
        _same_graph(want, got)
    ref.write_csv(want, path_or_buf=tmp_path / "a.csv")
    pp.io.write_csv(got, path_or_buf=tmp_path / "b.csv")
    assert (tmp_path / "a.csv").read_text() == (tmp_path / "b.csv").read_text()
    if strings:
        _same_graph(ref.read_csv_graph(str(tmp_path / "a.csv"), multiedges=multi), pp.io.read_csv_graph(str(tmp_path / "b.csv"), multiedges=multi))
    # temporal side: events with ties in time (both constructors use this package's stable time ordering)
    t = rng.integers(0, 40, 300)
    tdf = pd.DataFrame({"v": v, "w": w, "t": t, "weight": rng.integers(1, 9, 300).astype(float)})
    twant, tgot = ref.df_to_temporal_graph(tdf.copy(), multiedges=multi), pp.io.df_to_temporal_graph(tdf.copy(), multiedges=multi)
    _same_graph(twant, tgot)
    assert torch.equal(twant.data.time, tgot.data.time)
    assert ref.temporal_graph_to_df(twant).equals(pp.io.temporal_graph_to_df(tgot))
