"""oracle/containers.py against the golden vectors the reference's own method bodies produced
(tests/golden/make_container_golden.py) and, when /root/reference is mounted, against those bodies live."""
import os

import numpy as np
import pytest
import torch

from oracle import containers, ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "container_golden.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("i", range(4))
def test_graph_members_match_reference(gold, i):
    ei, w, n = gold[f"g{i}_edge_index"], gold[f"g{i}_edge_weight"], int(gold[f"g{i}_num_nodes"])
    u_ei, u_w, _ = containers.graph_to_undirected(ei, n, w)
    assert torch.equal(u_ei, gold[f"g{i}_undirected_edge_index"]) and torch.equal(u_w, gold[f"g{i}_undirected_edge_weight"])
    w_ei, w_w = containers.graph_to_weighted(ei, n)
    assert torch.equal(w_ei, gold[f"g{i}_weighted_edge_index"]) and torch.equal(w_w, gold[f"g{i}_weighted_edge_weight"])
    assert float(w_w.sum()) == ei.size(1)


@pytest.mark.parametrize("i", range(3))
def test_to_static_graph_matches_reference(gold, i):
    ei, t, window = gold[f"t{i}_edge_index"], gold[f"t{i}_time"], tuple(gold[f"t{i}_window"].tolist())
    for tag, kw in (("plain", {}), ("weighted", {"weighted": True}), ("window", {"weighted": True, "time_window": window})):
        got_ei, got_w, _ = containers.temporal_to_static(ei, t, kw.get("weighted", False), kw.get("time_window"))
        assert torch.equal(got_ei, gold[f"t{i}_{tag}_edge_index"]), tag
        if got_w is not None:
            assert torch.equal(got_w, gold[f"t{i}_{tag}_edge_weight"]), tag


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_live_reference_bodies_agree():
    make_self, ref = ref_loader.container_methods()
    g = torch.Generator().manual_seed(5)
    ei = torch.randint(0, 30, (2, 400), generator=g)
    ei = ei[:, torch.sort(ei[0], stable=True).indices]
    w = torch.randint(1, 5, (400,), generator=g).float()
    u = ref["to_undirected"](make_self(ei, 30, edge_weight=w))
    u_ei, u_w, _ = containers.graph_to_undirected(ei, 30, w)
    assert torch.equal(u.data.edge_index, u_ei) and torch.equal(u.data.edge_weight, u_w)
    assert u.data.edge_index.is_undirected if hasattr(u.data.edge_index, "is_undirected") else True
