"""Golden vectors of the edge-merging container members, produced by the REFERENCE'S OWN method bodies
(``Graph.to_undirected`` / ``Graph.to_weighted_graph`` core/graph.py:211-270, ``TemporalGraph.to_static_graph``
core/temporal_graph.py:191-220), compiled from /root/reference by ``oracle/ref_loader.container_methods``.

    python tests/golden/make_container_golden.py        # needs /root/reference (read-only mount)

Inputs are seeded random multigraphs / event streams; inputs and outputs are stored together in
``tests/golden/container_golden.npz``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "container_golden.npz")
GRAPH_CASES = [(5, 12), (40, 300), (300, 5000), (8, 200)]          # (nodes, edges): multi-edges and self-loops included
EVENT_CASES = [(6, 40, 10), (50, 2000, 300), (200, 20000, 1000)]   # (nodes, events, horizon)


def main() -> None:
    assert ref_loader.available(), "reference tree not mounted"
    make_self, ref = ref_loader.container_methods()
    g = torch.Generator().manual_seed(20261021)
    out: dict[str, np.ndarray] = {}
    for i, (n, e) in enumerate(GRAPH_CASES):
        ei = torch.randint(0, n, (2, e), generator=g)
        ei = ei[:, torch.sort(ei[0], stable=True).indices]      # a Graph holds its edges sorted by row
        w = torch.randint(1, 9, (e,), generator=g).float()
        u = ref["to_undirected"](make_self(ei, n, edge_weight=w))
        wg = ref["to_weighted_graph"](make_self(ei, n))
        out[f"g{i}_edge_index"], out[f"g{i}_edge_weight"], out[f"g{i}_num_nodes"] = ei.numpy(), w.numpy(), np.int64(n)
        out[f"g{i}_undirected_edge_index"] = u.data.edge_index.numpy()
        out[f"g{i}_undirected_edge_weight"] = u.data.edge_weight.numpy()
        out[f"g{i}_weighted_edge_index"] = wg.data.edge_index.numpy()
        out[f"g{i}_weighted_edge_weight"] = wg.data.edge_weight.numpy()
    for i, (n, m, horizon) in enumerate(EVENT_CASES):
        ei = torch.randint(0, n, (2, m), generator=g)
        t = torch.sort(torch.randint(0, horizon, (m,), generator=g)).values
        window = (horizon // 4, horizon // 2)
        out[f"t{i}_edge_index"], out[f"t{i}_time"], out[f"t{i}_window"] = ei.numpy(), t.numpy(), np.array(window)
        for tag, kw in (("plain", {}), ("weighted", {"weighted": True}), ("window", {"weighted": True, "time_window": window})):
            s = ref["to_static_graph"](make_self(ei, n, time=t), **kw)
            out[f"t{i}_{tag}_edge_index"] = s.data.edge_index.numpy()
            if kw.get("weighted"):
                out[f"t{i}_{tag}_edge_weight"] = s.data.edge_weight.numpy()
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
