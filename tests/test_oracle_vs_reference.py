"""The oracle against the reference's own source, executed from /root/reference on random inputs.
Skipped where the reference tree is absent (the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import lift, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("seed", range(5))
def test_lift_order_edge_index(seed):
    g = torch.Generator().manual_seed(seed)
    n, e = 60 + 40 * seed, 500 + 700 * seed
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = ei[:, torch.sort(ei[0], stable=True).indices]
    w = torch.randint(1, 9, (e,), generator=g).float()
    L = ref_loader.lift_order_module()
    assert torch.equal(L.lift_order_edge_index(ei, n), lift.lift_order_edge_index(ei, n))
    for rule in ("src", "dst", "max", "mul", "add"):
        a = L.lift_order_edge_index_weighted(ei, w, n, rule)
        b = lift.lift_order_edge_index_weighted(ei, w, n, rule)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("seed", range(5))
def test_aggregate_edge_index(seed):
    g = torch.Generator().manual_seed(100 + seed)
    k = 1 + seed % 4
    rows, e = 300, 4000
    ns = torch.randint(0, 7, (rows, k), generator=g) if k > 1 else torch.randperm(rows, generator=g).unsqueeze(1)
    ei = torch.randint(0, rows, (2, e), generator=g)
    w = torch.randint(1, 4, (e,), generator=g).float()
    a = ref_loader.lift_order_module().aggregate_edge_index(ei.clone(), ns.clone(), w.clone())
    b = lift.aggregate_edge_index(ei, ns, w)
    assert torch.equal(a.data.edge_index, b.edge_index)
    assert torch.equal(a.data.edge_weight, b.edge_weight)
    assert torch.equal(a.data.node_sequence, b.node_sequence)
    assert torch.equal(a.data.inverse_idx, b.inverse_idx)


@pytest.mark.parametrize("seed,delta", [(0, 1), (1, 5), (2, 20), (3, 2.5), (4, 1.5)])
def test_lift_order_temporal(seed, delta):
    g = torch.Generator().manual_seed(200 + seed)
    n, m = 40, 1500
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, 200, (m,), generator=g)).values
    a = ref_loader.ref_lift_order_temporal(ei, t, delta)
    assert torch.equal(a, lift.lift_order_temporal(ei, t, delta))
    if isinstance(delta, int):
        assert np.array_equal(a.numpy(), lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), delta))
    tf = t.double() * 0.25
    a = ref_loader.ref_lift_order_temporal(ei, tf, float(delta))
    assert np.array_equal(a.numpy(), lift.lift_order_temporal_closed_form(ei.numpy(), tf.numpy(), float(delta)))
