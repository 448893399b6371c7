"""Shortest time-respecting paths and temporal closeness on the GPU (csrc/paths.cu) against the oracle (the
reference's scipy-dijkstra formulation) and the reference's known answers."""
import numpy as np
import pytest
import torch

import pathpyg_b200 as pp
from oracle import paths
from pathpyg_b200 import ops

pytestmark = pytest.mark.gpu

LONG = [("a", "b", 1), ("b", "c", 5), ("c", "d", 9), ("c", "e", 9), ("c", "f", 11), ("f", "a", 13), ("a", "g", 18),
        ("b", "f", 21), ("a", "g", 26), ("c", "f", 27), ("h", "f", 27), ("g", "h", 28), ("a", "c", 30), ("a", "b", 31),
        ("c", "h", 32), ("f", "h", 33), ("b", "i", 42), ("i", "b", 42), ("c", "i", 47), ("h", "i", 50)]  # tests/core/conftest.py:50-73
INF = float("inf")


def test_known_answer(cuda):  # reference tests/algorithms/test_temporal.py:20-93
    g = pp.TemporalGraph.from_edge_list(LONG)
    dist, pred = pp.algorithms.temporal_shortest_paths(g, delta=10)
    true_dist = np.array([[0, 1, 1, 3, 3, 3, 1, 2, INF], [3, 0, 1, 2, 2, 1, 4, 5, 1], [2, INF, 0, 1, 1, 1, 3, 1, 1],
                          [INF, INF, INF, 0, INF, INF, INF, INF, INF], [INF, INF, INF, INF, 0, INF, INF, INF, INF],
                          [1, INF, INF, INF, INF, 0, 2, 1, INF], [INF, INF, INF, INF, INF, INF, 0, 1, INF],
                          [INF, INF, INF, INF, INF, 1, INF, 0, 1], [INF, 1, INF, INF, INF, INF, INF, INF, 0]])
    true_pred = np.array([[0, 0, 0, 2, 2, 2, 0, 2, -1], [5, 1, 1, 2, 2, 1, 0, 6, 1], [5, -1, 2, 2, 2, 2, 0, 2, 2],
                          [-1, -1, -1, 3, -1, -1, -1, -1, -1], [-1, -1, -1, -1, 4, -1, -1, -1, -1],
                          [5, -1, -1, -1, -1, 5, 0, 5, -1], [-1, -1, -1, -1, -1, -1, 6, 6, -1],
                          [-1, -1, -1, -1, -1, 7, -1, 7, 7], [-1, 8, -1, -1, -1, -1, -1, -1, 8]])
    assert dist.shape == pred.shape == (g.n, g.n)
    assert np.array_equal(dist, true_dist) and np.array_equal(pred, true_pred)


def test_closeness_known_answer(cuda):  # reference tests/algorithms/test_centrality.py:58-70
    g = pp.TemporalGraph.from_edge_list(LONG).to(cuda)
    c = pp.algorithms.temporal_closeness_centrality(g, delta=5)
    assert c == {"a": 12.0, "b": 16.0, "c": 16.0, "d": 14.666666666666666, "e": 14.666666666666666, "f": 24.0,
                 "g": 14.666666666666666, "h": 28.0, "i": 24.0}


@pytest.mark.parametrize("seed,n,m,horizon,delta", [(0, 12, 150, 60, 5), (1, 40, 1200, 300, 12), (2, 70, 3000, 100, 3),
                                                    (3, 33, 500, 500, 40), (4, 200, 8000, 2000, 25)])
def test_vs_oracle_random(cuda, seed, n, m, horizon, delta):
    gen = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=gen)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=gen)).values
    want_dist, want_pred = paths.temporal_shortest_paths(ei, t, n, delta)
    g = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    dist, pred = pp.algorithms.temporal_shortest_paths(g, delta)
    assert np.array_equal(dist, want_dist)
    # the predecessor scipy reports is the source of the LAST event that ends a shortest path; the kernel applies that
    # rule explicitly, so the matrices agree entry by entry, and every entry is a valid predecessor
    assert np.array_equal(pred, want_pred)
    if n <= 40:
        assert paths.is_valid_pred(ei, t, delta, dist, pred)
    got = pp.algorithms.temporal_closeness_centrality(g, delta)
    want = paths.temporal_closeness_centrality(want_dist)
    assert [got[i] for i in range(n)] == want.tolist()


def test_source_chunks_and_edge_cases(cuda):
    gen = torch.Generator().manual_seed(9)
    n, m = 100, 2000
    ei = torch.randint(0, n, (2, m), generator=gen).to(cuda)
    t = torch.sort(torch.randint(0, 400, (m,), generator=gen)).values.to(cuda)
    eg = ops.lift_order_temporal(ei, t, 10, n)
    whole = ops.temporal_paths(ei, eg, n)
    chunked = ops.temporal_paths(ei, eg, n, max_workspace_bytes=1)  # 32 sources per call
    assert torch.equal(whole[0], chunked[0]) and torch.equal(whole[1], chunked[1])
    # no event pair at all: only direct edges connect
    dist, pred = ops.temporal_paths(torch.tensor([[0, 1], [1, 2]], device=cuda), None, 3)
    assert dist.cpu().tolist() == [[0, 1, INF], [INF, 0, 1], [INF, INF, 0]]
    assert pred.cpu().tolist() == [[0, 0, -1], [-1, 1, 1], [-1, -1, 2]]
    with pytest.raises(ValueError):
        ops.temporal_paths(torch.tensor([[0], [7]], device=cuda), None, 3)


def test_betweenness_known_answer(cuda):  # reference tests/algorithms/test_centrality.py:45-55
    g = pp.TemporalGraph.from_edge_list(LONG)
    bw = pp.algorithms.temporal_betweenness_centrality(g, delta=5)
    assert [bw[x] for x in "abcdefghi"] == [2.0, 2.0, 4.5, 0, 0, 2.0, 0.5, 0, 0]
    assert bw["not a node"] == 0.0                                   # defaultdict like the reference's return value


@pytest.mark.parametrize("seed,n,m,horizon,delta", [(0, 10, 120, 50, 4), (1, 25, 600, 200, 9), (2, 40, 1500, 60, 2),
                                                    (3, 16, 400, 400, 50), (4, 60, 2500, 1000, 30)])
def test_betweenness_vs_oracle_random(cuda, seed, n, m, horizon, delta):
    gen = torch.Generator().manual_seed(100 + seed)
    ei = torch.randint(0, n, (2, m), generator=gen)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=gen)).values
    want = paths.temporal_betweenness_centrality(ei, t, n, delta)     # the reference's Python loops, restated
    g = pp.TemporalGraph.from_tensors(ei.to(cuda), t.to(cuda), n)
    got = pp.algorithms.temporal_betweenness_centrality(g, delta)
    got = np.array([got[i] for i in range(n)])
    # float64 sums of fractions in a different (fixed) order than the reference's stack order: 1e-12 relative
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    eg = ops.lift_order_temporal(ei.to(cuda), t.to(cuda), delta, n)
    small = ops.temporal_betweenness(ei.to(cuda), t.to(cuda), eg, n, max_workspace_bytes=1)   # one source per launch
    assert torch.equal(small.cpu(), torch.from_numpy(got))           # batching does not change a single bit
