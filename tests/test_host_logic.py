"""Host-side containers and the no-fallback rule.  CPU only."""
import os

import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from pathpyg_b200 import ops


def test_path_data_matches_reference_fixture():  # tests/core/test_path_data.py:18-74
    p = pp.PathData(pp.IndexMap(["a", "c", "b", "d", "e"]))
    p.append_walk(("a", "c", "d"), weight=1.0)
    p.append_walk(("a", "c"), weight=1.0)
    p.append_walk(("b", "c", "d"), weight=1.5)
    p.append_walk(("b", "c", "e"), weight=1.0)
    assert p.num_paths == 4
    assert [p.get_walk(i) for i in range(4)] == [("a", "c", "d"), ("a", "c"), ("b", "c", "d"), ("b", "c", "e")]
    assert torch.equal(p.data.dag_weight, torch.tensor([1.0, 1.0, 1.5, 1.0]))
    assert torch.equal(p.data.dag_num_nodes, torch.tensor([3, 2, 3, 3]))
    assert torch.equal(p.data.dag_num_edges, torch.tensor([2, 1, 2, 2]))
    q = pp.PathData(pp.IndexMap(["a", "c", "b", "d", "e"]))
    q.append_walks([("a", "c", "d"), ("a", "c"), ("b", "c", "d"), ("b", "c", "e")], weights=[1.0, 1.0, 1.5, 1.0])
    for key in ("edge_index", "node_sequence", "dag_weight", "dag_num_edges", "dag_num_nodes"):
        assert torch.equal(p.data[key], q.data[key]), key
    assert p.data.num_nodes == q.data.num_nodes == 11
    assert p.data.num_edges == 7
    assert p.map_node_seq([0, 1, 2]) == ["a", "c", "b"]
    assert "4 paths" in str(p)


def test_temporal_graph_sorts_by_time_and_keeps_attrs():
    g = pp.TemporalGraph.from_edge_list([("c", "d", 9), ("a", "b", 1), ("c", "e", 9), ("b", "c", 5)])
    assert g.n == 5 and g.m == 4
    assert g.data.time.tolist() == [1, 5, 9, 9]
    assert g.data.edge_index.as_tensor().tolist() == [[0, 1, 2, 2], [1, 2, 3, 4]]
    assert g.tedge_to_index[(0, 1, 1)] == 0
    assert g.temporal_edges[1] == ("b", "c", 5)
    assert g.data.is_sorted_by_time()


def test_index_map_and_lazy_higher_order_map():
    base = pp.IndexMap(["a", "b", "c"])
    assert base.to_idx("b") == 1 and base.to_id(2) == "c"
    assert base.to_idxs([["a", "b"], ["c", "a"]]).tolist() == [[0, 1], [2, 0]]
    with pytest.raises(ValueError):
        base.add_id("a")
    ns = torch.tensor([[0, 1], [1, 2], [2, 0]])
    lazy = pp.HigherOrderIndexMap(base, ns)
    eager = pp.IndexMap([tuple(base.to_ids(v)) for v in ns])  # the reference's construction (multi_order_model.py:119)
    assert lazy.num_ids() == 3
    assert lazy == eager
    assert lazy.to_idx(("b", "c")) == 1 and lazy.to_id(2) == ("c", "a")
    assert np.array_equal(lazy.to_ids([0, 2]), eager.to_ids([0, 2]))


def test_graph_container_sorts_by_row_and_validates():
    g = pp.Graph.from_edge_list([("a", "b"), ("b", "c"), ("a", "c"), ("a", "b")])  # tests/core/conftest.py:18-21
    assert g.data.edge_index.as_tensor().tolist() == [[0, 0, 0, 1], [1, 2, 1, 2]]
    assert g.n == 3 and g.m == 4 and g.order == 1
    assert g.row_ptr.tolist() == [0, 3, 4, 4]
    assert g.get_predecessors(2).tolist() == [0, 1]
    assert g.is_edge("a", "c") and not g.is_edge("c", "a")
    with pytest.raises(ValueError):
        pp.Graph(pp.Data(edge_index=torch.tensor([[0, 5], [1, 1]]), num_nodes=3))


def test_edge_index_wrapper_equality_and_device_move():
    ei = pp.EdgeIndex([[0, 1], [1, 2]], sparse_size=(3, 3))
    assert torch.equal(ei, pp.EdgeIndex([[0, 1], [1, 2]]))
    assert ei.as_tensor().tolist() == [[0, 1], [1, 2]]
    assert isinstance(ei.to("cpu"), pp.EdgeIndex)
    assert type(ei[0]) is torch.Tensor


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a machine without a GPU")
def test_no_cpu_fallback():
    ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 0, 3, 0]])
    with pytest.raises(RuntimeError, match="CUDA"):
        pp.algorithms.lift_order_edge_index(ei, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.lift_order_edge_index(ei, 4)
    g = pp.TemporalGraph.from_edge_list([("a", "b", 1), ("b", "c", 5)])
    with pytest.raises(RuntimeError, match="CUDA"):
        pp.MultiOrderModel.from_temporal_graph(g, delta=5, max_order=2)
    with pytest.raises(ValueError):
        pp.algorithms.aggregate_node_attributes(ei, torch.ones(4), "unknown")


def test_product_never_imports_the_oracle():
    import pathlib
    root = pathlib.Path(pp.__file__).parent
    for path in root.rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_bench_prints_only_the_json_line_on_stdout(tmp_path):
    """bench.py's contract is ONE JSON line on stdout: whatever a library writes to file descriptor 1 during the run
    (NCCL announces its version there) must end up on stderr."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "bench._only_json_on_stdout(lambda: (os.write(1, b'library chatter\\n'), print('{\"metric\": 1}')))\n" % root)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-500:]
    assert res.stdout == '{"metric": 1}\n'
    assert "library chatter" in res.stderr


def test_bench_reference_arm_and_gpu_arm_describe_the_same_config():
    """The driver compares the two arms' `config`: both come from bench.static_config."""
    import bench

    for name in bench.WORKLOADS:
        for world in (1, 2, 8):
            cfg = bench.static_config(name, world)
            assert cfg["name"] == name and cfg["workload"] == bench.WORKLOADS[name]["label"]
            assert "model" not in cfg
