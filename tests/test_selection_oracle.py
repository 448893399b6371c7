"""The oracle's model-selection statistics against the reference's known answers and the golden vectors that the
reference's own method bodies produced (tests/golden/make_selection_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch
from scipy.stats import chi2

from oracle import mom, ref_loader
from oracle import selection as sel

LLH_RTOL = 1e-5  # fp32 sums (np.isclose default of the reference's own tests)


@pytest.fixture(scope="module")
def sel_golden():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selection_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def model(seqs, weights=None, K=2):
    walks = mom.append_walks(seqs, weights or [1.0] * len(seqs))
    return mom.from_path_data(walks, max_order=K), walks


def walks_of(g, i):
    flat, lengths, weights = g[f"sel{i}_flat"], g[f"sel{i}_lengths"], g[f"sel{i}_weights"]
    seqs, o = [], 0
    for length in lengths.tolist():
        seqs.append(flat[o:o + length].tolist())
        o += length
    return seqs, weights.tolist()


def test_dof_known_answers():  # tests/core/test_multi_order_model.py:45-62
    layers, _ = model([(0, 1, 2, 3)], K=4)
    assert [sel.get_mon_dof(layers, k) for k in range(5)] == [3] * 5
    layers, _ = model([(0, 2, 3), (1, 2, 4)])
    assert [sel.get_mon_dof(layers, k) for k in range(3)] == [4, 5, 7]
    with pytest.raises(ValueError):
        sel.get_mon_dof(layers, 3)
    with pytest.raises(ValueError):
        sel.get_mon_dof(layers, 1, "walks")


def test_log_likelihood_known_answers():  # tests/core/test_multi_order_model.py:102-143
    layers, w = model([(0, 2, 3), (1, 2, 4)])
    want = [np.log(1 / 6) * 4 + np.log(2 / 6) * 2, np.log(1 / 6) * 2 + 2 * np.log(1 / 2), np.log(1 / 6) * 2]
    assert np.allclose([sel.get_mon_log_likelihood(layers, w, k) for k in range(3)], want)
    layers, w = model([(0, 2, 3), (1, 2, 4), (0, 2, 4), (1, 2, 3)])
    want = [np.log(2 / 12) * 8 + np.log(4 / 12) * 4, np.log(2 / 12) * 4 + 4 * np.log(1 / 2), np.log(1 / 6) * 4 + 4 * np.log(1 / 2)]
    assert np.allclose([sel.get_mon_log_likelihood(layers, w, k) for k in range(3)], want)
    layers, w = model([(0,), (0, 1), (0, 1, 2)])
    want = [np.log(3 / 6) * 3 + np.log(2 / 6) * 2 + np.log(1 / 6), np.log(3 / 6) * 3, np.log(3 / 6) * 3]
    assert np.allclose([sel.get_mon_log_likelihood(layers, w, k) for k in range(3)], want)


def test_likelihood_ratio_known_answer():  # tests/core/test_multi_order_model.py:65-99
    llh = [np.log(1 / 6) * 4 + np.log(2 / 6) * 2, np.log(1 / 6) * 2 + 2 * np.log(1 / 2), np.log(1 / 6) * 2]
    dof = [4, 5, 7]
    layers, w = model([(0, 2, 3), (1, 2, 4)])
    for a, b in ((0, 1), (1, 2)):
        p = 1 - chi2.cdf(-2 * (llh[a] - llh[b]), dof[b] - dof[a])
        reject, p_code = sel.likelihood_ratio_test(layers, w, a, b, significance_threshold=0.1)
        assert reject == (p < 0.1) and np.isclose(p_code, p)


def test_estimate_order_known_answers():  # tests/core/test_multi_order_model.py:146-162,193-224
    layers, w = model([(0, 2, 3), (1, 2, 4)], [3, 3])
    assert sel.estimate_order(layers, w, 2) == 1
    layers, w = model([(0, 2, 3), (1, 2, 4)], [4, 4])
    assert sel.estimate_order(layers, w, 2) == 2
    layers, w = model([(3, 1, 2), (0, 1, 2), (0, 1, 4), (3, 1, 4), (0,)], [1, 20, 1, 20, 1], K=3)
    assert sel.estimate_order(layers, w, 3) == 2


@pytest.mark.parametrize("i", range(4))
def test_selection_golden(sel_golden, i):
    g = sel_golden
    seqs, weights = walks_of(g, i)
    K = int(g[f"sel{i}_max_order"])
    walks = mom.append_walks(seqs, weights)
    layers = mom.from_path_data(walks, max_order=K)
    assert [sel.get_mon_dof(layers, k) for k in range(K + 1)] == g[f"sel{i}_dof_paths"].tolist()
    assert [float(sel.get_mon_dof(layers, k, "ngrams")) for k in range(K + 1)] == g[f"sel{i}_dof_ngrams"].tolist()
    assert np.allclose([sel.get_mon_log_likelihood(layers, walks, k) for k in range(K + 1)], g[f"sel{i}_llh"], rtol=LLH_RTOL)
    assert np.allclose([sel.get_intermediate_order_log_likelihood(layers, walks, k) for k in range(1, K)], g[f"sel{i}_llh_mid"], rtol=LLH_RTOL)
    for k in range(1, K + 1):
        assert torch.equal(sel.transition_probabilities(layers[k], True), torch.from_numpy(g[f"sel{i}_tp{k}"]))
        assert torch.equal(sel.transition_probabilities(layers[k]), torch.from_numpy(g[f"sel{i}_tp_unit{k}"]))
        assert torch.equal(sel.degrees(layers[k], "in", True), torch.from_numpy(g[f"sel{i}_indeg{k}"]))
        assert torch.equal(sel.degrees(layers[k], "out"), torch.from_numpy(g[f"sel{i}_outdeg_unit{k}"]))
        reject, p = sel.likelihood_ratio_test(layers, walks, k - 1, k)
        assert bool(reject) == bool(g[f"sel{i}_lrt_reject"][k - 1]) and np.isclose(p, g[f"sel{i}_lrt_p"][k - 1], atol=1e-9)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("seed", range(3))
def test_oracle_vs_reference_methods(seed):
    """The reference's own method bodies, executed from /root/reference, on fresh random walks."""
    gen = torch.Generator().manual_seed(300 + seed)
    n, p = 20 + 10 * seed, 300
    lengths = torch.randint(1, 6, (p,), generator=gen)
    flat = torch.randint(0, n, (int(lengths.sum()),), generator=gen)
    flat[:n] = torch.arange(n)
    weights = torch.randint(1, 4, (p,), generator=gen).float()
    seqs, o = [], 0
    for length in lengths.tolist():
        seqs.append(flat[o:o + length].tolist())
        o += length
    walks = mom.append_walks(seqs, weights.tolist())
    layers = mom.from_path_data(walks, max_order=3)
    Model, _ = ref_loader.selection_methods()
    ref, dag = Model(layers), ref_loader.ref_walks_data(walks)
    for k in range(4):
        assert ref.get_mon_dof(k) == sel.get_mon_dof(layers, k)
        assert ref.get_mon_log_likelihood(dag, k) == sel.get_mon_log_likelihood(layers, walks, k)
    for k in (1, 2, 3):
        a, b = ref.likelihood_ratio_test(dag, k - 1, k), sel.likelihood_ratio_test(layers, walks, k - 1, k)
        assert bool(a[0]) == bool(b[0]) and a[1] == b[1]
