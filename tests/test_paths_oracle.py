"""The oracle's shortest time-respecting paths (the reference's scipy formulation) against the reference's known
answers.  CPU only."""
import numpy as np
import torch

from oracle import paths

LONG = [("a", "b", 1), ("b", "c", 5), ("c", "d", 9), ("c", "e", 9), ("c", "f", 11), ("f", "a", 13), ("a", "g", 18),
        ("b", "f", 21), ("a", "g", 26), ("c", "f", 27), ("h", "f", 27), ("g", "h", 28), ("a", "c", 30), ("a", "b", 31),
        ("c", "h", 32), ("f", "h", 33), ("b", "i", 42), ("i", "b", 42), ("c", "i", 47), ("h", "i", 50)]  # tests/core/conftest.py:50-73
INF = float("inf")


def tensors():
    ids = sorted({x for e in LONG for x in e[:2]})
    lut = {v: i for i, v in enumerate(ids)}
    return torch.tensor([[lut[e[0]] for e in LONG], [lut[e[1]] for e in LONG]]), torch.tensor([e[2] for e in LONG]), len(ids)


def test_temporal_shortest_paths_known_answer():  # reference tests/algorithms/test_temporal.py:20-93
    ei, t, n = tensors()
    dist, pred = paths.temporal_shortest_paths(ei, t, n, 10)
    true_dist = np.array([[0, 1, 1, 3, 3, 3, 1, 2, INF], [3, 0, 1, 2, 2, 1, 4, 5, 1], [2, INF, 0, 1, 1, 1, 3, 1, 1],
                          [INF, INF, INF, 0, INF, INF, INF, INF, INF], [INF, INF, INF, INF, 0, INF, INF, INF, INF],
                          [1, INF, INF, INF, INF, 0, 2, 1, INF], [INF, INF, INF, INF, INF, INF, 0, 1, INF],
                          [INF, INF, INF, INF, INF, 1, INF, 0, 1], [INF, 1, INF, INF, INF, INF, INF, INF, 0]])
    true_pred = np.array([[0, 0, 0, 2, 2, 2, 0, 2, -1], [5, 1, 1, 2, 2, 1, 0, 6, 1], [5, -1, 2, 2, 2, 2, 0, 2, 2],
                          [-1, -1, -1, 3, -1, -1, -1, -1, -1], [-1, -1, -1, -1, 4, -1, -1, -1, -1],
                          [5, -1, -1, -1, -1, 5, 0, 5, -1], [-1, -1, -1, -1, -1, -1, 6, 6, -1],
                          [-1, -1, -1, -1, -1, 7, -1, 7, 7], [-1, 8, -1, -1, -1, -1, -1, -1, 8]])
    assert np.array_equal(dist, true_dist) and np.array_equal(pred, true_pred)
    assert paths.is_valid_pred(ei, t, 10, dist, pred)
    bad = pred.copy()
    bad[0, 7] = 3
    assert not paths.is_valid_pred(ei, t, 10, dist, bad)


def test_temporal_closeness_known_answer():  # reference tests/algorithms/test_centrality.py:58-70
    ei, t, n = tensors()
    dist, _ = paths.temporal_shortest_paths(ei, t, n, 5)
    want = [12.0, 16.0, 16.0, 14.666666666666666, 14.666666666666666, 24.0, 14.666666666666666, 28.0, 24.0]
    assert paths.temporal_closeness_centrality(dist).tolist() == want


def test_temporal_betweenness_known_answer():  # reference tests/algorithms/test_centrality.py:45-55
    ei, t, n = tensors()
    bw = paths.temporal_betweenness_centrality(ei, t, n, 5)
    assert bw.tolist() == [2.0, 2.0, 4.5, 0.0, 0.0, 2.0, 0.5, 0.0, 0.0]
