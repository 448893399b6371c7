"""CUDA path vs oracle / golden vectors for a1-a4, through the public API (which goes through the C ABI).
Bit-exact (torch.equal) for every index tensor; weights are integer-valued fp32, hence exact too."""
import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import lift
from pathpyg_b200 import ops
from pathpyg_b200.algorithms import (aggregate_edge_index, aggregate_node_attributes, lift_order_edge_index,
                                     lift_order_edge_index_weighted, lift_order_temporal)

pytestmark = pytest.mark.gpu


class TG:  # what lift_order_temporal reads from a TemporalGraph
    def __init__(self, ei, t, n):
        self.data = pp.Data(edge_index=ei, time=t, num_nodes=n)


def sorted_multigraph(seed, n, e):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    return ei[:, torch.sort(ei[0], stable=True).indices].contiguous()


# ------------------------------------------------------------------ a3
def test_aggregate_node_attributes_known_answer(cuda):  # reference tests/algorithms/test_lift_order.py:12-31
    ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 0, 3, 0]], device=cuda)
    a = torch.tensor([1, 2, 3, 4], device=cuda)
    want = {"src": [1, 2, 3, 3, 4], "dst": [2, 3, 1, 4, 1], "max": [2, 3, 3, 4, 4], "mul": [2, 6, 3, 12, 4], "add": [3, 5, 4, 7, 5]}
    for rule, v in want.items():
        out = aggregate_node_attributes(ei, a, rule)
        assert out.dtype == torch.int64 and out.is_cuda and out.tolist() == v
    with pytest.raises(ValueError):
        aggregate_node_attributes(ei, a, "unknown")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.int64, torch.int32])
def test_aggregate_node_attributes_dtypes(cuda, dtype):
    g = torch.Generator().manual_seed(3)
    ei = torch.randint(0, 500, (2, 7000), generator=g)
    a = torch.randint(-20, 20, (500,), generator=g).to(dtype)
    for rule in ("src", "dst", "max", "mul", "add"):
        assert torch.equal(aggregate_node_attributes(ei.to(cuda), a.to(cuda), rule).cpu(), lift.aggregate_node_attributes(ei, a, rule))


# ------------------------------------------------------------------ a2
def test_lift_order_edge_index_known_answer(cuda):  # reference tests/algorithms/test_lift_order.py:34-57
    ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 0, 3, 0]], device=cuda)
    assert lift_order_edge_index(ei, 4).tolist() == [[0, 1, 1, 2, 3, 4], [1, 2, 3, 0, 4, 0]]
    assert lift_order_edge_index(ei).tolist() == [[0, 1, 1, 2, 3, 4], [1, 2, 3, 0, 4, 0]]  # num_nodes inferred
    ho, w = lift_order_edge_index_weighted(ei, torch.tensor([1, 2, 3, 4, 5], device=cuda), 4)
    assert ho.tolist() == [[0, 1, 1, 2, 3, 4], [1, 2, 3, 0, 4, 0]] and w.tolist() == [1, 2, 2, 3, 4, 5]


@pytest.mark.parametrize("case", [f"lift{i}" for i in range(4)])
def test_lift_golden(cuda, golden, case):
    ei = torch.from_numpy(golden[f"{case}_edge_index"]).to(cuda)
    n = int(golden[f"{case}_num_nodes"])
    out = lift_order_edge_index(ei, n)
    assert out.dtype == torch.int64 and out.is_contiguous()
    assert np.array_equal(out.cpu().numpy(), golden[f"{case}_out"])
    w = torch.from_numpy(golden[f"{case}_weight"]).to(cuda)
    for rule in ("src", "dst", "max", "mul", "add"):
        ho, hw = lift_order_edge_index_weighted(ei, w, n, rule)
        assert np.array_equal(hw.cpu().numpy(), golden[f"{case}_w_{rule}"])


@pytest.mark.parametrize("seed,n,e", [(0, 1, 1), (1, 3, 2), (2, 10, 2047), (3, 10, 2049), (4, 1000, 4096), (5, 100, 100000),
                                      (6, 100000, 1000000), (7, 50000, 300000)])
def test_lift_random_vs_oracle(cuda, seed, n, e):
    ei = sorted_multigraph(seed, n, e)
    assert torch.equal(lift_order_edge_index(ei.to(cuda), n).cpu(), lift.lift_order_edge_index(ei, n))


def test_lift_edge_cases(cuda):
    empty = torch.empty((2, 0), dtype=torch.long, device=cuda)
    assert lift_order_edge_index(empty, 5).shape == (2, 0)
    assert lift_order_edge_index(empty).shape == (2, 0)
    # no continuation at all: every edge ends in a sink
    ei = torch.tensor([[0, 0, 1], [2, 3, 3]], device=cuda)
    assert lift_order_edge_index(ei, 4).shape == (2, 0)
    # self loops and multi-edges are ordinary edges
    ei = torch.tensor([[0, 0, 0, 1, 1], [0, 0, 1, 0, 1]])
    assert torch.equal(lift_order_edge_index(ei.to(cuda), 2).cpu(), lift.lift_order_edge_index(ei, 2))
    # a star: one hub with huge fan-out after long runs of zero-count edges (skewed expand)
    hub_in = torch.stack([torch.arange(1, 3001), torch.zeros(3000, dtype=torch.long)])
    hub_out = torch.stack([torch.zeros(5000, dtype=torch.long), torch.arange(3001, 8001)])
    ei = torch.cat([hub_out, hub_in], dim=1)  # row 0 first => sorted by row
    assert torch.equal(lift_order_edge_index(ei.to(cuda), 8001).cpu(), lift.lift_order_edge_index(ei, 8001))
    with pytest.raises(ValueError):  # id outside [0, num_nodes)
        lift_order_edge_index(torch.tensor([[0, 1], [1, 7]], device=cuda), 3)


def test_lift_host_tensors_round_trip(cuda):
    ei = sorted_multigraph(11, 40, 500)
    out = lift_order_edge_index(ei, 40)  # host in -> host out, computed on the GPU
    assert not out.is_cuda and torch.equal(out, lift.lift_order_edge_index(ei, 40))


# ------------------------------------------------------------------ a1
def test_temporal_known_answer(cuda):  # reference tests/algorithms/test_temporal.py:11-17
    g = pp.TemporalGraph.from_edge_list([("a", "b", 1), ("b", "c", 5), ("c", "d", 9), ("c", "e", 9)]).to(cuda)
    out = lift_order_temporal(g, delta=5)
    assert out.is_cuda and out.tolist() == [[0, 1, 1], [1, 2, 3]]
    with pytest.raises(RuntimeError):  # the reference's torch.cat([]) (temporal.py:53)
        lift_order_temporal(g, delta=1)


@pytest.mark.parametrize("case", [f"temp{i}" for i in range(6)])
def test_temporal_golden(cuda, golden, case):
    ei = torch.from_numpy(golden[f"{case}_edge_index"]).to(cuda)
    t = torch.from_numpy(golden[f"{case}_time"]).to(cuda)
    out = lift_order_temporal(TG(ei, t, int(golden[f"{case}_num_nodes"])), golden[f"{case}_delta"].item())
    assert out.is_contiguous() and np.array_equal(out.cpu().numpy(), golden[f"{case}_out"])


@pytest.mark.parametrize("seed,n,m,horizon,delta", [(0, 5, 40, 10, 2), (1, 100, 5000, 50, 3), (2, 1000, 20000, 100000, 900),
                                                   (3, 100000, 1000000, 1000, 200), (4, 2, 3000, 30, 5), (5, 50, 4097, 64, 64)])
def test_temporal_random_vs_oracle(cuda, seed, n, m, horizon, delta):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=g)).values
    want = lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), delta)
    out = lift_order_temporal(TG(ei.to(cuda), t.to(cuda), n), delta)
    assert np.array_equal(out.cpu().numpy(), want)


def test_temporal_float_promotions(cuda):
    """int64 time + float delta compares in float32, float64 time adds double(float32(delta)) (temporal.py:30,43)."""
    g = torch.Generator().manual_seed(9)
    ei = torch.randint(0, 30, (2, 1200), generator=g)
    t = torch.sort(torch.randint(1_700_000_000, 1_700_000_400, (1200,), generator=g)).values  # epoch seconds: float32 loses the low bits
    for delta in (2.5, 64.0, 130.7):
        want = lift.lift_order_temporal(ei, t, delta)
        assert torch.equal(lift_order_temporal(TG(ei.to(cuda), t.to(cuda), 30), delta).cpu(), want)
    tf = (t - 1_700_000_000).double() * 0.1
    for delta in (0.3, 1.1):
        want = lift.lift_order_temporal(ei, tf, delta)
        assert torch.equal(lift_order_temporal(TG(ei.to(cuda), tf.to(cuda), 30), delta).cpu(), want)


@pytest.mark.parametrize("seed,n,m,horizon,delta,as_float", [(0, 12, 400, 40, 3, False), (1, 300, 20_000, 500, 9, False),
                                                             (2, 20, 600, 30, 2.5, True)])
def test_temporal_lift_of_shuffled_time_stamps(cuda, seed, n, m, horizon, delta, as_float):
    """The reference's lift does not need a time-ordered event list (temporal.py:33-53: masks per distinct time
    stamp); its DBGNN tutorial lifts right after TemporalGraph.shuffle_time().  Pairs come out by ascending source
    time stamp, source position, target position."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.randint(0, horizon, (m,), generator=g)        # NOT sorted
    if as_float:
        t = t.double() / 2
    want = lift.lift_order_temporal(ei, t, delta)
    got = lift_order_temporal(TG(ei.to(cuda), t.to(cuda), n), delta)
    assert torch.equal(got.cpu(), want)


def test_model_from_a_graph_with_shuffled_time_stamps(cuda):
    """TemporalGraph.shuffle_time() then from_temporal_graph (the reference's DBGNN tutorial): the event graph is taken
    over the SHUFFLED positions of g.data, node sequences and weights from the re-sorted events
    (multi_order_model.py:148-170).  Distinct time stamps, so that the re-sorting has no ties to break.  Orders 1
    and 2 only: beyond, the reference hands this event graph -- no longer sorted by its source row -- to
    lift_order_edge_index, whose precondition (lift_order.py:56, "sorted edge index") it violates; what comes out
    there is not a defined result."""
    from oracle import mom

    g = torch.Generator().manual_seed(4)
    n, m = 15, 500
    ei = torch.randint(0, n, (2, m), generator=g)
    tg = pp.TemporalGraph.from_tensors(ei.to(cuda), torch.arange(m, device=cuda), n)
    tg.shuffle_time()
    shuffled = tg.data.time.cpu()
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=25, max_order=2)
    order = torch.sort(shuffled, stable=True).indices
    ref = mom.from_temporal_graph(ei[:, order], shuffled[order], n, delta=25, max_order=2,
                                  event_graph=lift.lift_order_temporal(ei, shuffled, 25))
    for k in (1, 2):
        assert torch.equal(model.layers[k].data.edge_index.as_tensor().cpu(), ref[k].edge_index), k
        assert torch.equal(model.layers[k].data.edge_weight.cpu(), ref[k].edge_weight), k
        assert torch.equal(model.layers[k].data.node_sequence.cpu(), ref[k].node_sequence), k


def test_pair_attributes_rejects_ids_outside_the_attribute(cuda):
    ei = torch.tensor([[0, 1, 5], [1, 2, 0]], device=cuda)
    with pytest.raises(IndexError):
        aggregate_node_attributes(ei, torch.ones(4, device=cuda), "src")
    with pytest.raises(IndexError):
        aggregate_node_attributes(torch.tensor([[0, -1], [1, 0]], device=cuda), torch.ones(4, device=cuda), "add")
    with pytest.raises(IndexError):  # weights shorter than the edge list they belong to
        lift_order_edge_index_weighted(torch.tensor([[0, 1], [1, 0]], device=cuda), torch.ones(1, device=cuda), 2)


def test_temporal_rejects_bad_ids(cuda):
    ei = torch.tensor([[0, 9], [1, 2]], device=cuda)
    with pytest.raises(ValueError):
        lift_order_temporal(TG(ei, torch.tensor([1, 2], device=cuda), 3), 5)


# ------------------------------------------------------------------ a4
def test_aggregate_known_answer(cuda):  # reference tests/algorithms/test_lift_order.py:60-79 (int64 weights kept)
    g = aggregate_edge_index(torch.tensor([[0, 2, 2, 1], [1, 1, 3, 0]], device=cuda),
                             torch.tensor([[1, 2], [2, 3], [1, 2], [4, 5]], device=cuda), torch.tensor([1, 2, 3, 4], device=cuda))
    assert g.data.edge_index.as_tensor().tolist() == [[0, 0, 1], [1, 2, 0]]
    assert g.data.edge_weight.tolist() == [3, 3, 4] and g.data.edge_weight.dtype == torch.int64
    assert g.data.node_sequence.tolist() == [[1, 2], [2, 3], [4, 5]]
    assert g.data.inverse_idx.tolist() == [0, 1, 0, 2]
    assert g.n == 3 and g.order == 2


@pytest.mark.parametrize("case", [f"agg{i}" for i in range(4)])
def test_aggregate_golden(cuda, golden, case):
    ei = torch.from_numpy(golden[f"{case}_edge_index"]).to(cuda)
    ns = torch.from_numpy(golden[f"{case}_node_sequence"]).to(cuda)
    w = torch.from_numpy(golden[f"{case}_weight"]).to(cuda)
    g = aggregate_edge_index(ei, ns, w)
    assert np.array_equal(g.data.edge_index.as_tensor().cpu().numpy(), golden[f"{case}_out_edge_index"])
    assert np.array_equal(g.data.edge_weight.cpu().numpy(), golden[f"{case}_out_weight"])
    assert np.array_equal(g.data.node_sequence.cpu().numpy(), golden[f"{case}_out_node_sequence"])
    assert np.array_equal(g.data.inverse_idx.cpu().numpy(), golden[f"{case}_out_inverse"])
    assert g.data.num_nodes == int(golden[f"{case}_out_num_nodes"])
    g = aggregate_edge_index(ei, ns)  # default unit weights
    assert np.array_equal(g.data.edge_weight.cpu().numpy(), golden[f"{case}_out_weight_unit"])


@pytest.mark.parametrize("k,rows,vals", [(1, 1, 1), (2, 5000, 50), (3, 100000, 40), (4, 30000, 70000), (7, 20000, 600), (9, 3000, 1 << 40)])
def test_unique_rows_vs_oracle(cuda, k, rows, vals):
    g = torch.Generator().manual_seed(k)
    ns = torch.randint(0, vals, (rows, k), generator=g)
    ns[rows // 2:] = ns[: rows - rows // 2]  # force duplicates also for wide value ranges
    u, inv = ops.unique_rows(ns.to(cuda))
    wu, winv = lift.unique_rows_closed_form(ns.numpy())
    assert np.array_equal(u.cpu().numpy(), wu) and np.array_equal(inv.cpu().numpy(), winv)


def test_unique_rows_negative_and_sorted_inputs(cuda):
    ns = torch.tensor([[3, -2], [-5, 7], [3, -2], [0, 0], [-5, 6]], device=cuda)
    u, inv = ops.unique_rows(ns)
    tu, tinv = torch.unique(ns.cpu(), dim=0, return_inverse=True)
    assert torch.equal(u.cpu(), tu) and torch.equal(inv.cpu(), tinv)
    asc = torch.arange(1000, device=cuda).unsqueeze(1)  # layer-1 case: already the sorted distinct rows
    u, inv = ops.unique_rows(asc)
    assert torch.equal(u, asc) and torch.equal(inv, torch.arange(1000, device=cuda))


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.int64])
def test_coalesce_reductions(cuda, reduce, dtype):
    from oracle import pyg
    g = torch.Generator().manual_seed(5)
    n, e = 40, 6000
    ei = torch.randint(0, n, (2, e), generator=g)
    w = torch.randint(-6, 9, (e,), generator=g).to(dtype)
    out_ei, out_w = ops.coalesce(ei.to(cuda), None, n, w.to(cuda), reduce)
    want_ei, want_w = pyg.coalesce(ei, w, n, reduce)
    assert torch.equal(out_ei.cpu(), want_ei)
    if dtype.is_floating_point and reduce == "mean":
        assert torch.allclose(out_w.cpu(), want_w, rtol=1e-6)
    else:
        assert torch.equal(out_w.cpu(), want_w)


def test_aggregate_rejects_ids_beyond_num_nodes(cuda):
    # first order: node_sequence values are used as ids; 7 >= #distinct values (EdgeIndex.validate in the reference)
    with pytest.raises(ValueError):
        aggregate_edge_index(torch.tensor([[0, 1], [1, 2]], device=cuda), torch.tensor([[0], [3], [7]], device=cuda))


def test_aggregate_large_vs_oracle(cuda):
    g = torch.Generator().manual_seed(21)
    m, n = 1_000_000, 100_000
    ei = torch.randint(0, n, (2, m), generator=g)
    ns = ei.t().contiguous()                       # order-2 node sequences of a 1M-edge stream
    ho = torch.randint(0, m, (2, 2_000_000), generator=g)
    layer = aggregate_edge_index(ho.to(cuda), ns.to(cuda))
    wu, winv = lift.unique_rows_closed_form(ns.numpy())
    assert np.array_equal(layer.data.node_sequence.cpu().numpy(), wu)
    assert np.array_equal(layer.data.inverse_idx.cpu().numpy(), winv)
    mapped = torch.from_numpy(winv)[ho]
    key = mapped[0] * wu.shape[0] + mapped[1]
    uk, counts = torch.unique(key, return_counts=True)
    assert torch.equal(layer.data.edge_index.as_tensor().cpu(), torch.stack([uk // wu.shape[0], uk % wu.shape[0]]))
    assert torch.equal(layer.data.edge_weight.cpu(), counts.float())
