"""One warm-up build, then one build of MultiOrderModel.from_temporal_graph inside a cudaProfilerStart/Stop range:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/chain_profile.py cfg3
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import pathpyg_b200 as pp  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cfg = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
ei, t = bench.make_stream(cfg, seed=0)
tg = pp.TemporalGraph.from_tensors(ei.to(dev), t.to(dev), cfg["n"])
for i in range(2):
    if i == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=cfg["delta"], max_order=cfg["order"])
    torch.cuda.synchronize()
    if i == 1:
        torch.cuda.cudart().cudaProfilerStop()
    del model
