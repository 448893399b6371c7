"""Host-side cost of one from_temporal_graph build + DBGNN forward at cfg2 (cProfile; development aid)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathpyg_b200 as pp  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
n, m = 100_000, 1_000_000
ei = torch.randint(0, n, (2, m), generator=g).to(dev)
t = torch.sort(torch.randint(0, 1000, (m,), generator=g)).values.to(dev)
tg = pp.TemporalGraph.from_tensors(ei, t, n)
for _ in range(5):
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=200, max_order=2)
torch.cuda.synchronize()
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=200, max_order=2)
torch.cuda.synchronize()
print(f"wall per build: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=200, max_order=2)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
