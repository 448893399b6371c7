"""One from_temporal_graph build at a BASELINE configuration, for an ncu launch list (development aid)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathpyg_b200 as pp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=1_000_000)
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--delta", type=int, default=200)
ap.add_argument("--order", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ei = torch.randint(0, a.n, (2, a.m), generator=g).to(dev)
t = torch.sort(torch.randint(0, a.T, (a.m,), generator=g)).values.to(dev)
tg = pp.TemporalGraph.from_tensors(ei, t, a.n)
for _ in range(a.reps):
    torch.cuda.synchronize()
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=a.delta, max_order=a.order)
    torch.cuda.synchronize()
    print("BUILD", [(k, v.n, v.m) for k, v in model.layers.items()], flush=True)
