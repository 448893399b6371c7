"""Ingest step (pathpyg_b200.io) and its oracle against the reference's known answers (tests/io/test_pandas.py) and
the golden vectors made by the reference's own io/pandas.py.  Host logic runs everywhere; the device-side time
ordering is checked in the gpu-marked tests."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

import pathpyg_b200 as pp
from oracle import ingest
from pathpyg_b200.io.pandas import _parse_df_column, _parse_timestamp


@pytest.fixture(scope="module")
def ing_golden():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def event_frame(g, i):
    return pd.DataFrame({"v": g[f"ev{i}_v"], "w": g[f"ev{i}_w"], "t": g[f"ev{i}_t"], "weight": g[f"ev{i}_weight"]})


def check_events(g, i, node_ids, edge_index, time, weight, num_nodes):
    assert np.array_equal(np.asarray(node_ids).astype(g[f"ev{i}_out_node_ids"].dtype), g[f"ev{i}_out_node_ids"])
    assert np.array_equal(edge_index, g[f"ev{i}_out_edge_index"])
    assert np.array_equal(time, g[f"ev{i}_out_time"])
    assert np.array_equal(weight, g[f"ev{i}_out_weight"])
    assert num_nodes == int(g[f"ev{i}_out_num_nodes"])


# ------------------------------------------------------------------------------------------ known answers (host)
def test_parse_timestamp_known_answers():  # reference tests/io/test_pandas.py:60-111
    df = pd.DataFrame({"t": ["2023-01-01 12:00:00", "2023-01-01 13:00:00"]})
    _parse_timestamp(df)
    assert np.issubdtype(df["t"].dtype, np.integer) and df["t"].iloc[1] > df["t"].iloc[0]
    df = pd.DataFrame({"t": ["01/01/2023 12:00", "01/01/2023 13:00"]})
    _parse_timestamp(df, timestamp_format="%d/%m/%Y %H:%M")
    assert df["t"].iloc[1] > df["t"].iloc[0]
    df = pd.DataFrame({"t": [1000, 2000, 3000]})
    _parse_timestamp(df)
    assert np.all(df["t"] == np.array([1000, 2000, 3000]))
    df = pd.DataFrame({"t": pd.to_datetime(["2023-01-01", "2023-01-02"])})
    _parse_timestamp(df)
    assert np.issubdtype(df["t"].dtype, np.integer) and df["t"].iloc[1] > df["t"].iloc[0]
    df = pd.DataFrame({"t": ["2023-01-01 12:00:00", "2023-01-01 13:00:00"]})
    _parse_timestamp(df, time_rescale=10**6 if pd.__version__ >= "3.0.0" else 10**9)
    assert np.all(df["t"].diff().dropna() == 3600)
    with pytest.raises(ValueError, match="Column `t` must be of type"):
        _parse_timestamp(pd.DataFrame({"t": [None, None]}))


def test_parse_df_column_known_answers():  # reference tests/io/test_pandas.py:114-171
    data = pp.Data(edge_index=torch.tensor([[0, 1, 2], [1, 2, 0]]))
    back = np.array([2, 1, 0])
    _parse_df_column(pd.DataFrame({"attr": [1, 2, 3]}), data, "attr")
    assert torch.equal(data["attr"], torch.tensor([1, 2, 3]))
    _parse_df_column(pd.DataFrame({"attr": ["1", "2", "3"]}), data, "attr", prefix="node_", idx=back)
    assert torch.equal(data["node_attr"], torch.tensor([3, 2, 1]))
    _parse_df_column(pd.DataFrame({"attr": ["1.1", "2.2", "3.3"]}), data, "attr", prefix="edge_", idx=back)
    assert torch.allclose(data["edge_attr"], torch.tensor([3.3, 2.2, 1.1], dtype=torch.double))
    _parse_df_column(pd.DataFrame({"attr": ["[1, 2]", "[3, 4]", "[5, 6]"]}), data, "attr", prefix="edge_")
    assert torch.equal(data["edge_attr"], torch.tensor([[1, 2], [3, 4], [5, 6]]))
    _parse_df_column(pd.DataFrame({"attr": [(1, 2), (3, 4), (5, 6)]}), data, "attr", prefix="node_", idx=back)
    assert torch.equal(data["node_attr"], torch.tensor([[5, 6], [3, 4], [1, 2]]))
    _parse_df_column(pd.DataFrame({"attr": ["foo", "bar", "baz"]}), data, "attr", prefix="edge_", idx=back)
    assert np.array_equal(data["edge_attr"], np.array(["baz", "bar", "foo"]))


def test_df_to_temporal_graph_known_answers():  # reference tests/io/test_pandas.py:300-349
    g = pp.io.df_to_temporal_graph(pd.DataFrame({"v": ["a", "b", "c"], "w": ["b", "c", "a"], "t": [1, 2, 3]}))
    assert (g.n, g.m) == (3, 3) and torch.equal(g.data.time, torch.tensor([1, 2, 3]))
    g = pp.io.df_to_temporal_graph(pd.DataFrame({"v": ["a", "b"], "w": ["b", "c"], "t": [20, 10], "weight": [2.0, 1.0]}))
    assert torch.allclose(g.data.edge_weight, torch.tensor([1.0, 2.0], dtype=torch.double))
    dup = {"v": ["a", "a", "b"], "w": ["b", "b", "c"], "t": [1, 1, 2]}
    assert pp.io.df_to_temporal_graph(pd.DataFrame(dup), multiedges=False).m == 2
    assert pp.io.df_to_temporal_graph(pd.DataFrame(dup), multiedges=True).m == 3
    g = pp.io.df_to_temporal_graph(pd.DataFrame([["a", "b", 1], ["b", "c", 2], ["c", "a", 3], ["a", "b", 4]]))
    assert (g.n, g.m) == (3, 4)
    g = pp.io.df_to_temporal_graph(pd.DataFrame({"v": ["a", "b"], "w": ["b", "c"], "t": [1000, 2000]}), time_rescale=1000)
    assert torch.equal(g.data.time, torch.tensor([1, 2]))
    g = pp.io.df_to_temporal_graph(pd.DataFrame({"v": ["a", "b"], "w": ["b", "c"], "t": [1, 2], "foo": [10, 20], "edge_bar": [0.1, 0.2]}))
    assert torch.equal(g.data.edge_foo, torch.tensor([10, 20]))
    assert torch.allclose(g.data.edge_bar, torch.tensor([0.1, 0.2], dtype=torch.double))
    df = pp.io.temporal_graph_to_df(g)
    assert df["v"].tolist() == ["a", "b"] and df["w"].tolist() == ["b", "c"] and df["t"].tolist() == [1, 2]


# ------------------------------------------------------------------------------------------ golden vectors (host)
@pytest.mark.parametrize("i", range(4))
def test_events_golden_oracle_and_host(ing_golden, i):
    g = ing_golden
    kw = dict(multiedges=bool(g[f"ev{i}_multiedges"]), time_rescale=int(g[f"ev{i}_rescale"]))
    o = ingest.df_to_temporal_graph(event_frame(g, i), **kw)
    check_events(g, i, o["node_ids"], o["edge_index"].numpy(), o["time"].numpy(), o["edge_weight"].numpy(), o["num_nodes"])
    tg = pp.io.df_to_temporal_graph(event_frame(g, i), **kw)
    check_events(g, i, tg.mapping.node_ids, tg.data.edge_index.as_tensor().numpy(), tg.data.time.numpy(),
                 tg.data.edge_weight.numpy(), tg.n)


@pytest.mark.parametrize("i", range(3))
def test_paths_golden_oracle_and_host(ing_golden, i, tmp_path):
    g = ing_golden
    path = tmp_path / "walks.ngram"
    path.write_text("\n".join(g[f"pa{i}_lines"].tolist()) + "\n")
    weighted = bool(g[f"pa{i}_weighted"])
    ids, walks = ingest.read_csv_path_data(str(path), weight=weighted)
    pdata = pp.io.read_csv_path_data(str(path), weight=weighted)
    for node_ids, d in ((ids, walks), (pdata.mapping.node_ids, pdata.data)):
        assert np.array_equal(np.asarray(node_ids).astype(str), g[f"pa{i}_out_node_ids"])
        assert np.array_equal(d.edge_index.numpy(), g[f"pa{i}_out_edge_index"])
        assert np.array_equal(d.node_sequence.numpy(), g[f"pa{i}_out_node_sequence"])
        assert np.array_equal(d.dag_weight.numpy(), g[f"pa{i}_out_dag_weight"])
        assert np.array_equal(d.dag_num_nodes.numpy(), g[f"pa{i}_out_dag_num_nodes"])
        assert np.array_equal(d.dag_num_edges.numpy(), g[f"pa{i}_out_dag_num_edges"])
    assert pdata.get_walk(0) == tuple(g[f"pa{i}_lines"][0].split(",")[: len(g[f"pa{i}_lines"][0].split(",")) - int(weighted)])


def test_index_map_bulk_lookup_matches_dictionary():
    ids = np.array([f"id{i}" for i in np.random.default_rng(0).permutation(500)])
    m = pp.IndexMap(ids)
    q = ids[np.random.default_rng(1).integers(0, 500, (2, 300))]
    want = torch.tensor([[m.id_to_idx[v] for v in row] for row in q.tolist()])
    assert torch.equal(m.to_idxs(q), want)
    with pytest.raises(KeyError):
        m.to_idxs(np.array(["missing"] * 100))
    pd_ = pp.PathData(pp.IndexMap(list("abcde")))
    pd_.append_walks([("a", "c", "d"), ("b", "c", "e")], [1.0, 2.0])           # reference tests/core/test_path_data.py:57-74
    assert pd_.data.node_sequence.squeeze().tolist() == [0, 2, 3, 1, 2, 4]
    assert pd_.data.edge_index.tolist() == [[0, 1, 3, 4], [1, 2, 4, 5]]


# ------------------------------------------------------------------------------------------ device path
@pytest.mark.gpu
@pytest.mark.parametrize("i", range(4))
def test_events_golden_device(cuda, ing_golden, i):
    g = ing_golden
    tg = pp.io.df_to_temporal_graph(event_frame(g, i), multiedges=bool(g[f"ev{i}_multiedges"]),
                                    time_rescale=int(g[f"ev{i}_rescale"]), device=cuda)
    assert tg.data.edge_index.is_cuda and tg.data.time.is_cuda
    check_events(g, i, tg.mapping.node_ids, tg.data.edge_index.as_tensor().cpu().numpy(), tg.data.time.cpu().numpy(),
                 tg.data.edge_weight.cpu().numpy(), tg.n)


@pytest.mark.gpu
def test_stable_argsort_device(cuda):
    from pathpyg_b200 import ops

    gen = torch.Generator().manual_seed(5)
    for t in (torch.randint(-50, 50, (100_000,), generator=gen), torch.randint(0, 1 << 40, (300_000,), generator=gen),
              torch.randint(-(1 << 62), 1 << 62, (50_000,), generator=gen) * 2,
              torch.randn(200_000, generator=gen, dtype=torch.float64).round(decimals=2),
              torch.tensor([0.0, -0.0, 1.5, -1.5, 0.0, float("inf"), -float("inf")], dtype=torch.float64)):
        want = torch.sort(t, stable=True).indices
        assert torch.equal(ops.stable_argsort(t.to(cuda)).cpu(), want)
    g = pp.TemporalGraph.from_tensors(torch.randint(0, 9, (2, 1000), generator=gen).to(cuda),
                                      torch.rand(1000, generator=gen, dtype=torch.float64).to(cuda), 9)
    assert bool((g.data.time[1:] >= g.data.time[:-1]).all())


@pytest.mark.gpu
def test_csv_round_trip_to_lift(cuda, tmp_path):
    """csv -> TemporalGraph on the GPU -> MultiOrderModel: the ingest feeds the lift without a host detour."""
    path = tmp_path / "events.csv"
    pd.DataFrame({"v": list("abccbd"), "w": list("bcdeda"), "t": [1, 5, 9, 9, 7, 12]}).to_csv(path, index=False)
    g = pp.io.read_csv_temporal_graph(str(path), device=cuda)
    m = pp.MultiOrderModel.from_temporal_graph(g, delta=4, max_order=2)
    host = pp.MultiOrderModel.from_temporal_graph(pp.io.read_csv_temporal_graph(str(path)), delta=4, max_order=2)
    for k in (1, 2):
        assert torch.equal(m.layers[k].data.edge_index.as_tensor().cpu(), host.layers[k].data.edge_index.as_tensor())
        assert torch.equal(m.layers[k].data.edge_weight.cpu(), host.layers[k].data.edge_weight)
