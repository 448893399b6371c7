"""Generate the committed golden vectors by running the REFERENCE'S OWN source in this container.

    python tests/golden/make_golden.py        # needs /root/reference (read-only mount)

``oracle/ref_loader.py`` executes ``src/pathpyG/algorithms/lift_order.py`` and ``lift_order_temporal``
from ``/root/reference`` unmodified (PyG utilities replaced by ``oracle/pyg.py``).  The inputs are
seeded; inputs and outputs are stored together in ``tests/golden/lift_golden.npz`` so that the GPU box
(which has no /root/reference) can check both the oracle and the CUDA path against them.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lift_golden.npz")


def sorted_multigraph(gen, num_nodes, num_edges):
    ei = torch.randint(0, num_nodes, (2, num_edges), generator=gen)
    order = torch.sort(ei[0], stable=True).indices
    return ei[:, order].contiguous()


def temporal_stream(gen, num_nodes, num_edges, horizon, float_time=False):
    ei = torch.randint(0, num_nodes, (2, num_edges), generator=gen)
    t = torch.sort(torch.randint(0, horizon, (num_edges,), generator=gen)).values
    if float_time:
        t = t.double() + torch.rand(num_edges, generator=gen, dtype=torch.float64).sort().values * 0.5
        t = t.sort().values
    return ei, t


def main() -> None:
    assert ref_loader.available(), "reference tree not mounted"
    L = ref_loader.lift_order_module()
    gen = torch.Generator().manual_seed(20261017)
    out: dict[str, np.ndarray] = {}

    # ---- lift_order_edge_index (+ weighted, all five rules)
    for i, (n, e) in enumerate([(7, 20), (50, 400), (300, 5000), (4000, 30000)]):
        ei = sorted_multigraph(gen, n, e)
        w = torch.randint(1, 6, (e,), generator=gen).float()
        out[f"lift{i}_edge_index"] = ei.numpy()
        out[f"lift{i}_num_nodes"] = np.int64(n)
        out[f"lift{i}_weight"] = w.numpy()
        out[f"lift{i}_out"] = L.lift_order_edge_index(ei, n).numpy()
        for rule in ("src", "dst", "max", "mul", "add"):
            ho, hw = L.lift_order_edge_index_weighted(ei, w, n, rule)
            assert torch.equal(ho, torch.from_numpy(out[f"lift{i}_out"]))
            out[f"lift{i}_w_{rule}"] = hw.numpy()

    # ---- aggregate_edge_index at k = 1, 2, 3 (weights are small integers => sums are exact)
    for i, (k, rows, vals, e) in enumerate([(1, 40, 40, 300), (2, 500, 12, 3000), (3, 2000, 9, 20000), (5, 1500, 4, 8000)]):
        if k == 1:
            ns = torch.arange(rows).unsqueeze(1)
        else:
            ns = torch.randint(0, vals, (rows, k), generator=gen)
        ei = torch.randint(0, rows, (2, e), generator=gen)
        w = torch.randint(1, 5, (e,), generator=gen).float()
        g = L.aggregate_edge_index(ei.clone(), ns.clone(), w.clone())
        out[f"agg{i}_edge_index"] = ei.numpy()
        out[f"agg{i}_node_sequence"] = ns.numpy()
        out[f"agg{i}_weight"] = w.numpy()
        out[f"agg{i}_out_edge_index"] = g.data.edge_index.numpy()
        out[f"agg{i}_out_weight"] = g.data.edge_weight.numpy()
        out[f"agg{i}_out_node_sequence"] = g.data.node_sequence.numpy()
        out[f"agg{i}_out_inverse"] = g.data.inverse_idx.numpy()
        out[f"agg{i}_out_num_nodes"] = np.int64(g.data.num_nodes)
        g = L.aggregate_edge_index(ei.clone(), ns.clone(), None)
        out[f"agg{i}_out_weight_unit"] = g.data.edge_weight.numpy()

    # ---- lift_order_temporal: int time / int delta, int time / float delta, float time / float delta
    cases = [(30, 200, 40, 3, False), (200, 3000, 500, 7, False), (100, 2500, 60, 2, False),
             (60, 1500, 300, 4.5, False), (60, 1500, 300, 2.75, True), (25, 400, 2000, 1, False)]
    for i, (n, m, horizon, delta, float_time) in enumerate(cases):
        ei, t = temporal_stream(gen, n, m, horizon, float_time)
        ho = ref_loader.ref_lift_order_temporal(ei, t, delta)
        out[f"temp{i}_edge_index"] = ei.numpy()
        out[f"temp{i}_time"] = t.numpy()
        out[f"temp{i}_delta"] = np.float64(delta) if isinstance(delta, float) else np.int64(delta)
        out[f"temp{i}_num_nodes"] = np.int64(n)
        out[f"temp{i}_out"] = ho.numpy()

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
