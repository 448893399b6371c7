"""Static-graph readers / writers and the attribute loaders of ``pathpyg_b200.io`` -- the known answers of the
reference's ``tests/io/test_pandas.py`` (:31-57, :174-297, :351-453, :495-528), and the reference's own module
executed live when /root/reference is mounted."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import ref_loader
from pathpyg_b200 import Graph, TemporalGraph
from pathpyg_b200.io import (add_edge_attributes, add_node_attributes, df_to_graph, graph_to_df, read_csv_graph,
                             temporal_graph_to_df, write_csv)
from pathpyg_b200.io.pandas import _integer_re, _iterable_re, _number_re


@pytest.fixture
def simple_graph():
    return Graph.from_edge_list([("a", "b"), ("b", "c"), ("a", "c")])


@pytest.fixture
def simple_temporal_graph():
    return TemporalGraph.from_edge_list([("a", "b", 1), ("b", "c", 5), ("c", "d", 9), ("c", "e", 9)])


def test_column_classifiers():  # :31-57
    assert _iterable_re.match("[1, 2, 3]") and _iterable_re.match("(1, 2, 3)") and _iterable_re.match("[[1, 2], [3, 4]]")
    for bad in ("{1, 2, 3}", "1, 2, 3", "1, 2, 3]", "(1, 2, 3"):
        assert not _iterable_re.match(bad)
    assert _number_re.match("1") and _number_re.match("1.0") and _number_re.match("1.0e10")
    for bad in ("1,000", "one", "1.0.0"):
        assert not _number_re.match(bad)
    assert _integer_re.match("1") and _integer_re.match("1000")
    for bad in ("1.0", "1.0e10", "1,000", "one", "1.0.0"):
        assert not _integer_re.match(bad)


def test_df_to_graph():  # :174-210 (the undirected case merges edges on the GPU: tests/test_containers_gpu.py)
    g = df_to_graph(pd.DataFrame({"v": ["a", "b", "c"], "w": ["b", "c", "a"]}))
    assert (g.n, g.m) == (3, 3)
    g = df_to_graph(pd.DataFrame({"v": ["a", "b", "c"], "w": ["b", "c", "a"], "edge_weight": [2.0, 1.0, 42.0]}))
    assert "edge_weight" in g.edge_attrs() and torch.equal(g.data.edge_weight, torch.tensor([2.0, 1.0, 42.0]).double())
    g = df_to_graph(pd.DataFrame([["a", "b", 2.0], ["b", "c", 1.0], ["c", "a", 42.0]]))
    assert "edge_attr_0" in g.edge_attrs() and g.data.edge_attr_0.tolist() == [2.0, 1.0, 42.0]
    multi = pd.DataFrame({"v": ["a", "b", "c", "a"], "w": ["b", "c", "a", "b"], "edge_weight": [2.0, 1.0, 42.0, 3.0]})
    assert df_to_graph(multi.copy(), multiedges=False).m == 3 and df_to_graph(multi.copy(), multiedges=True).m == 4
    g = df_to_graph(multi.copy(), multiedges=True)   # attributes follow the constructor's stable row sort
    assert g.data.edge_index.as_tensor().tolist() == [[0, 0, 1, 2], [1, 1, 2, 0]] and g.data.edge_weight.tolist() == [2.0, 3.0, 1.0, 42.0]
    g = df_to_graph(pd.DataFrame({"v": [3, 1], "w": [1, 2], "label": ["x", "y"]}), num_nodes=5)
    assert g.n == 5 and g.mapping.node_ids.tolist() == [1, 2, 3] and g.data.edge_label.tolist() == ["y", "x"]


def test_add_node_attributes(simple_graph):  # :213-246
    add_node_attributes(pd.DataFrame({"v": ["b", "a", "c"], "x": [2, 1, 3], "node_y": [0.2, 0.1, 0.3]}), simple_graph)
    assert torch.equal(simple_graph.data["node_x"], torch.tensor([1, 2, 3]))
    assert torch.allclose(simple_graph.data["node_y"], torch.tensor([0.1, 0.2, 0.3], dtype=torch.double))
    add_node_attributes(pd.DataFrame({"index": [1, 0, 2], "z": [20, 10, 30]}), simple_graph)
    assert torch.equal(simple_graph.data["node_z"], torch.tensor([10, 20, 30]))
    with pytest.raises(ValueError, match="multiple attribute values for single node"):
        add_node_attributes(pd.DataFrame({"v": ["a", "a", "b", "c"], "x": [1, 2, 3, 4]}), simple_graph)
    with pytest.raises(ValueError, match="Mismatch between nodes"):
        add_node_attributes(pd.DataFrame({"v": ["a", "b", "d"], "x": [1, 2, 3]}), simple_graph)
    with pytest.raises(ValueError, match="must either have `index` or `v` column"):
        add_node_attributes(pd.DataFrame({"foo": [1, 2, 3], "bar": [4, 5, 6]}), simple_graph)


def test_add_edge_attributes(simple_graph, simple_temporal_graph):  # :249-297
    add_edge_attributes(pd.DataFrame({"v": ["a", "b", "a"], "w": ["b", "c", "c"], "weight": [1, 3, 2]}), simple_graph)
    assert simple_graph.data["edge_weight"].tolist() == [1, 2, 3]
    add_edge_attributes(pd.DataFrame({"v": ["a", "b", "a"], "w": ["b", "c", "c"], "edge_score": [5, 6, 7]}), simple_graph)
    assert simple_graph.data["edge_score"].tolist() == [5, 7, 6]
    with pytest.raises(ValueError, match="Please ensure all nodes in the DataFrame are present in the graph."):
        add_edge_attributes(pd.DataFrame({"v": ["a", "x", "a"], "w": ["b", "c", "c"], "weight": [1.0, 2.0, 3.0]}), simple_graph)
    with pytest.raises(ValueError, match="does not exist in the graph"):
        add_edge_attributes(pd.DataFrame({"v": ["a", "b", "a"], "w": ["a", "c", "c"], "weight": [1.0, 2.0, 3.0]}), simple_graph)
    with pytest.raises(ValueError, match="must have columns `v` and `w`"):
        add_edge_attributes(pd.DataFrame({"v": ["a"], "weight": [1.0]}), simple_graph)
    tg = simple_temporal_graph
    add_edge_attributes(pd.DataFrame({"v": ["a", "b", "c", "c"], "w": ["b", "c", "e", "d"], "t": [1, 5, 9, 9],
                                      "weight": [1, 2, 4, 3]}), tg, time_attr="t")
    assert tg.data["edge_weight"].tolist() == [1, 2, 3, 4]
    with pytest.raises(ValueError, match="Please ensure the DataFrame matches the number of edges in the graph"):
        add_edge_attributes(pd.DataFrame({"v": ["a"], "w": ["b"], "t": [99], "weight": [1.0]}), tg, time_attr="t")
    with pytest.raises(ValueError, match="does not exist at time"):
        add_edge_attributes(pd.DataFrame({"v": ["a", "b", "c", "c"], "w": ["b", "c", "d", "e"], "t": [1, 5, 9, 10],
                                          "weight": [1.0, 2.0, 3.0, 4.0]}), tg, time_attr="t")
    with pytest.raises(ValueError, match="must have column when"):
        add_edge_attributes(pd.DataFrame({"v": ["a", "b", "c", "c"], "w": ["b", "c", "d", "e"]}), tg, time_attr="when")


def test_graph_to_df(simple_graph):  # :351-379
    df = graph_to_df(simple_graph)
    assert set(df.columns) == {"v", "w"} and len(df) == 3 and set(df["v"]) == {"a", "b"} and set(df["w"]) == {"b", "c"}
    simple_graph.data.edge_weight = torch.tensor([1.0, 2.0, 3.0])
    simple_graph.data.edge_label = torch.tensor([0, 1, 2])
    df = graph_to_df(simple_graph)
    assert list(df["edge_weight"]) == [1.0, 2.0, 3.0] and list(df["edge_label"]) == [0, 1, 2]
    df = graph_to_df(simple_graph, node_indices=True)
    assert set(df["v"]) == {0, 1} and set(df["w"]) == {1, 2}


def test_read_csv_graph(tmp_path):  # :415-453
    path = tmp_path / "graph.csv"
    pd.DataFrame({"v": ["a", "b", "a"], "w": ["b", "c", "c"]}).to_csv(path, index=False)
    g = read_csv_graph(str(path))
    assert isinstance(g, Graph) and (g.n, g.m) == (3, 3) and set(g.nodes) == {"a", "b", "c"}
    pd.DataFrame({"v": ["a", "b"], "w": ["b", "c"], "edge_weight": [1.0, 2.0]}).to_csv(path, index=False)
    assert torch.allclose(read_csv_graph(str(path)).data.edge_weight, torch.tensor([1.0, 2.0], dtype=torch.double))
    pd.DataFrame([["a", "b"], ["b", "c"], ["a", "c"]]).to_csv(path, index=False, header=False)
    g = read_csv_graph(str(path), header=False)
    assert (g.n, g.m) == (3, 3) and set(g.nodes) == {"a", "b", "c"}
    pd.DataFrame({"v": ["a", "a", "b"], "w": ["b", "b", "c"]}).to_csv(path, index=False)
    assert read_csv_graph(str(path), multiedges=False).m == 2 and read_csv_graph(str(path), multiedges=True).m == 3


def test_write_csv(tmp_path, simple_graph, simple_temporal_graph):  # :495-528
    simple_graph.data.edge_weight = torch.tensor([1.0, 2.0, 3.0])
    path = tmp_path / "graph.csv"
    write_csv(simple_graph, path_or_buf=path)
    df = pd.read_csv(path)
    assert set(df.columns) == {"v", "w", "edge_weight"} and len(df) == 3 and list(df["edge_weight"]) == [1.0, 2.0, 3.0]
    assert set(df["v"]) == {"a", "b"} and set(df["w"]) == {"b", "c"}
    simple_temporal_graph.data.edge_weight = torch.tensor([1.0, 2.0, 3.0, 4.0])
    write_csv(simple_temporal_graph, path_or_buf=path)
    df = pd.read_csv(path)
    assert set(df.columns) == {"v", "w", "t", "edge_weight"} and len(df) == 4 and set(df["t"]) == {1, 5, 9}
    assert list(df["edge_weight"]) == [1.0, 2.0, 3.0, 4.0]
    assert list(temporal_graph_to_df(simple_temporal_graph, node_indices=True)["v"]) == [0, 1, 2, 2]
    del simple_graph.data.__dict__["edge_weight"]
    write_csv(simple_graph, node_indices=True, path_or_buf=path)
    df = pd.read_csv(path)
    assert set(df.columns) == {"v", "w"} and set(df["v"]) == {0, 1} and set(df["w"]) == {1, 2}


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("strings, multi", [(True, False), (False, True), (True, True)])
def test_df_to_graph_matches_reference_module(strings, multi):
    """The reference's own io/pandas.py (executed by oracle/ref_loader.io_module) on a random edge table."""
    io = ref_loader.io_module()
    rng = np.random.default_rng(7)
    v, w = rng.integers(0, 40, 600), rng.integers(0, 40, 600)
    if strings:
        v, w = np.array([f"n{x:02d}" for x in v]), np.array([f"n{x:02d}" for x in w])
    df = pd.DataFrame({"v": v, "w": w, "weight": rng.integers(1, 9, 600).astype(float), "edge_tag": rng.integers(0, 5, 600)})
    want = io.df_to_graph(df.copy(), multiedges=multi)
    got = df_to_graph(df.copy(), multiedges=multi)
    assert np.array_equal(np.asarray(got.mapping.node_ids), np.asarray(want.mapping.node_ids))
    assert torch.equal(got.data.edge_index.as_tensor(), want.data.edge_index)
    assert torch.equal(got.data.edge_weight, want.data.edge_weight)
    assert got.n == want.data.num_nodes
