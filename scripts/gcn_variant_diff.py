"""Where do two tcgen05 layer kernels differ?  (development aid)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathpyg_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
F = H = 64
g = torch.Generator().manual_seed(5)
n, e = 148 * 128 * 3 + 77, 160_000
ei = torch.randint(0, n, (2, e), generator=g)
key = torch.unique(ei[0] * n + ei[1])
ei = torch.stack([key // n, key % n])
x = torch.randn(n, F, generator=g).to(dev)
W = (torch.randn(H, F, generator=g) / 8).to(dev)
graph = ops.gcn_prepare(ei.to(dev), None, n)
outs = {}
for v in ("staged", "single"):
    os.environ["PPG_GCN_TC"] = v
    outs[v] = ops.gcn_layer_tc(graph, x, W, None, _lib.ACT_NONE).cpu()
bad = (outs["staged"] != outs["single"]).any(1).nonzero().flatten()
print("rows", n, "wrong rows", bad.numel())
if bad.numel():
    tile = bad // 128
    print("t index (tile // 148) histogram:", torch.bincount(tile // 148).tolist())
    print("row-in-tile % 4 histogram:", torch.bincount(bad % 4, minlength=4).tolist())
    print("first wrong rows:", bad[:20].tolist())
    cp = graph.colptr.cpu()
    deg = (cp[1:] - cp[:-1])
    print("in-degree of wrong rows (first 20):", deg[bad[:20]].tolist())
    print("wrong rows per tile (first 12 tiles with errors):", torch.unique(tile, return_counts=True)[1][:12].tolist(), torch.unique(tile)[:12].tolist())
    # does a wrong row equal the correct value of another row of the same tile?
    r = int(bad[0])
    t0 = (r // 128) * 128
    d = (outs["single"][t0:t0 + 128] - outs["staged"][r]).abs().max(1).values
    print("row", r, "closest correct row in its tile:", int(d.argmin()) + t0, float(d.min()))
    print("E0 of its tile:", int(cp[t0]), "E0 % 4:", int(cp[t0]) % 4)
