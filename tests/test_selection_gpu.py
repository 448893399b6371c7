"""Model-selection statistics on the GPU (csrc/selection.cu behind MultiOrderModel / Graph) against the oracle,
the reference's known answers and the golden vectors produced by the reference's own method bodies."""
import os

import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import mom
from oracle import selection as sel
from pathpyg_b200 import ops
from pathpyg_b200.core.index_map import IndexMap

pytestmark = pytest.mark.gpu
LLH_RTOL = 1e-5  # fp32 terms; the reference sums them in fp32, the kernel in fp64 (np.isclose default of the reference tests)


@pytest.fixture(scope="module")
def sel_golden():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selection_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def build(cuda, seqs, weights=None, K=2, ids="abcde"):
    pd = pp.PathData(IndexMap(list(ids)), device=cuda)
    for s, w in zip(seqs, weights or [1.0] * len(seqs)):
        pd.append_walk(s, weight=w)
    return pp.MultiOrderModel.from_path_data(pd, max_order=K), pd


def test_dof_known_answers(cuda):  # reference tests/core/test_multi_order_model.py:45-62
    m, _ = build(cuda, [("a", "b", "c", "d")], K=4, ids="abcd")
    assert [m.get_mon_dof(assumption="paths", max_order=k) for k in range(5)] == [3] * 5
    m, _ = build(cuda, [("a", "c", "d"), ("b", "c", "e")])
    assert [m.get_mon_dof(assumption="paths", max_order=k) for k in range(3)] == [4, 5, 7]
    assert m.get_mon_dof(assumption="ngrams", max_order=2) == 4 + 5 * 4 + 25 * 4
    with pytest.raises(ValueError):
        m.get_mon_dof(max_order=3)
    with pytest.raises(ValueError):
        m.get_mon_dof(assumption="walks")


def test_log_likelihood_known_answers(cuda):  # reference tests/core/test_multi_order_model.py:102-143
    m, pd = build(cuda, [("a", "c", "d"), ("b", "c", "e")])
    want = [np.log(1 / 6) * 4 + np.log(2 / 6) * 2, np.log(1 / 6) * 2 + 2 * np.log(1 / 2), np.log(1 / 6) * 2]
    assert np.allclose([m.get_mon_log_likelihood(pd.data, max_order=k) for k in range(3)], want)
    m, pd = build(cuda, [("a", "c", "d"), ("b", "c", "e"), ("a", "c", "e"), ("b", "c", "d")])
    want = [np.log(2 / 12) * 8 + np.log(4 / 12) * 4, np.log(2 / 12) * 4 + 4 * np.log(1 / 2), np.log(1 / 6) * 4 + 4 * np.log(1 / 2)]
    assert np.allclose([m.get_mon_log_likelihood(pd.data, max_order=k) for k in range(3)], want)
    m, pd = build(cuda, [("a",), ("a", "b"), ("a", "b", "c")])
    want = [np.log(3 / 6) * 3 + np.log(2 / 6) * 2 + np.log(1 / 6), np.log(3 / 6) * 3, np.log(3 / 6) * 3]
    assert np.allclose([m.get_mon_log_likelihood(pd.data, max_order=k) for k in range(3)], want)


def test_estimate_order_known_answers(cuda):  # reference tests/core/test_multi_order_model.py:146-162,193-224
    m, pd = build(cuda, [("a", "c", "d"), ("b", "c", "e")], [3, 3])
    assert m.estimate_order(pd, max_order=2, significance_threshold=0.01) == 1
    m, pd = build(cuda, [("a", "c", "d"), ("b", "c", "e")], [4, 4])
    assert m.estimate_order(pd, max_order=2, significance_threshold=0.01) == 2
    m, pd = build(cuda, [("d", "b", "c"), ("a", "b", "c"), ("a", "b", "e"), ("d", "b", "e"), ("a",)], [1, 20, 1, 20, 1], K=3)
    assert m.estimate_order(pd, max_order=3) == 2
    with pytest.raises(ValueError):
        m.estimate_order(pd, max_order=4)
    with pytest.raises(ValueError):
        m.estimate_order(pd, max_order=1)
    with pytest.raises(ValueError):
        m.likelihood_ratio_test(pd.data, max_order_null=2, max_order=2)


def test_host_model_statistics(cuda):
    """A model built from host tensors lives on the host; its statistics stage to the GPU and agree."""
    pd = pp.PathData(IndexMap(list("abcde")))
    pd.append_walk(("a", "c", "d"))
    pd.append_walk(("b", "c", "e"))
    m = pp.MultiOrderModel.from_path_data(pd, max_order=2)
    assert not m.layers[1].data.edge_index.is_cuda
    assert [m.get_mon_dof(max_order=k) for k in range(3)] == [4, 5, 7]
    assert np.isclose(m.get_mon_log_likelihood(pd.data, max_order=1), np.log(1 / 6) * 2 + 2 * np.log(1 / 2))
    tp = m.layers[1].transition_probabilities(edge_attr="edge_weight")
    assert not tp.is_cuda and tp.tolist() == [1.0, 1.0, 0.5, 0.5]
    assert m.layers[1].degrees(mode="in") == {"a": 0, "b": 0, "c": 2, "d": 1, "e": 1}


@pytest.mark.parametrize("i", range(4))
def test_selection_golden(cuda, sel_golden, i):
    g = sel_golden
    n, K = int(g[f"sel{i}_num_nodes"]), int(g[f"sel{i}_max_order"])
    pd = pp.PathData(IndexMap(list(range(n))), device=cuda)
    pd.append_index_walks(torch.from_numpy(g[f"sel{i}_flat"]).to(cuda), torch.from_numpy(g[f"sel{i}_lengths"]).to(cuda),
                          torch.from_numpy(g[f"sel{i}_weights"]).to(cuda))
    m = pp.MultiOrderModel.from_path_data(pd, max_order=K)
    assert [m.get_mon_dof(max_order=k) for k in range(K + 1)] == g[f"sel{i}_dof_paths"].tolist()
    assert [float(m.get_mon_dof(max_order=k, assumption="ngrams")) for k in range(K + 1)] == g[f"sel{i}_dof_ngrams"].tolist()
    assert np.allclose([m.get_mon_log_likelihood(pd.data, max_order=k) for k in range(K + 1)], g[f"sel{i}_llh"], rtol=LLH_RTOL)
    assert np.allclose([m.get_intermediate_order_log_likelihood(pd.data, k) for k in range(1, K)], g[f"sel{i}_llh_mid"], rtol=LLH_RTOL)
    for k in range(1, K + 1):
        layer = m.layers[k]
        # weights are small integers: segment sums and the ratios are exact in fp32
        assert torch.equal(layer.transition_probabilities(edge_attr="edge_weight").cpu(), torch.from_numpy(g[f"sel{i}_tp{k}"]))
        assert torch.equal(layer.transition_probabilities().cpu(), torch.from_numpy(g[f"sel{i}_tp_unit{k}"]))
        assert torch.equal(layer.degrees("in", "edge_weight", True).cpu(), torch.from_numpy(g[f"sel{i}_indeg{k}"]))
        assert torch.equal(layer.degrees("out", None, True).cpu(), torch.from_numpy(g[f"sel{i}_outdeg_unit{k}"]))
        reject, p = m.likelihood_ratio_test(pd.data, max_order_null=k - 1, max_order=k)
        assert bool(reject) == bool(g[f"sel{i}_lrt_reject"][k - 1]) and np.isclose(p, g[f"sel{i}_lrt_p"][k - 1], atol=1e-6)


def test_statistics_vs_oracle_large(cuda):
    """200k walks over 2k nodes, orders 1-3: exact DoF, log-likelihoods within fp32 tolerance, one hub node with a
    50k-edge row (the long-segment path of segment_sum)."""
    gen = torch.Generator().manual_seed(7)
    n, p = 2000, 200_000
    lengths = torch.randint(1, 8, (p,), generator=gen)
    flat = torch.randint(0, n, (int(lengths.sum()),), generator=gen)
    flat[:n] = torch.arange(n)
    weights = torch.randint(1, 4, (p,), generator=gen).float()
    pd = pp.PathData(IndexMap(list(range(n))), device=cuda)
    pd.append_index_walks(flat.to(cuda), lengths.to(cuda), weights.to(cuda))
    m = pp.MultiOrderModel.from_path_data(pd, max_order=3)
    seqs, o = [], 0
    for length in lengths.tolist():
        seqs.append(flat[o:o + length].tolist())
        o += length
    walks = mom.append_walks(seqs, weights.tolist())
    layers = mom.from_path_data(walks, max_order=3)
    for k in range(3):  # DoF of order 3 needs the 4th lift in the oracle (minutes); 0-2 cover the recurrence
        assert m.get_mon_dof(max_order=k) == sel.get_mon_dof(layers, k)
    for k in range(4):
        a, b = m.get_mon_log_likelihood(pd.data, max_order=k), sel.get_mon_log_likelihood(layers, walks, k)
        assert abs(a - b) <= LLH_RTOL * abs(b), (k, a, b)
    for k in (1, 2, 3):
        assert torch.equal(m.layers[k].transition_probabilities("edge_weight").cpu(), sel.transition_probabilities(layers[k], True))


def test_segment_sum_long_and_short(cuda):
    gen = torch.Generator().manual_seed(3)
    sizes = torch.cat([torch.randint(0, 40, (500,), generator=gen), torch.tensor([100_000, 0, 65, 64, 1])])
    ids = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    w = torch.randint(1, 8, (ids.numel(),), generator=gen).float()
    ptr = ops.sorted_ids_ptr(ids.to(cuda), sizes.numel())
    assert torch.equal(ptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(sizes, 0)]))
    want = torch.zeros(sizes.numel()).scatter_add_(0, ids, w)
    assert torch.equal(ops.segment_sum(ptr, w.to(cuda)).cpu(), want)
    assert torch.equal(ops.segment_sum(ptr, None).cpu(), sizes.float())
    perm = torch.randperm(ids.numel(), generator=gen)
    got = ops.segment_sum(ptr, w[torch.argsort(perm)].to(cuda), perm.int().to(cuda))
    assert torch.equal(got.cpu(), want)
    # non-integer weights: the long segment is reduced by an fp64 tree, short ones in slot order
    wf = torch.rand(ids.numel(), generator=gen)
    got = ops.segment_sum(ptr, wf.to(cuda)).cpu()
    ref = torch.zeros(sizes.numel(), dtype=torch.float64).scatter_add_(0, ids, wf.double())
    assert torch.allclose(got.double(), ref, rtol=1e-5)


def test_walk_counts_and_log_sum_errors(cuda):
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 2, 0]], device=cuda)
    walks, sources = ops.walk_counts(ei, 3, 3)
    assert walks == [4, 5, 7] and sources == [3, 3, 3]
    with pytest.raises(ValueError):
        ops.walk_counts(torch.tensor([[0], [5]], device=cuda), 3, 1)
    with pytest.raises(IndexError):
        ops.weighted_log_sum(torch.ones(2, device=cuda), torch.ones(2, device=cuda), torch.tensor([0, 2], device=cuda))
    assert ops.weighted_log_sum(torch.empty(0, device=cuda), torch.ones(1, device=cuda)) == 0.0
