"""Device time of the fused GCN layer (segment-reduce gather + tcgen05 transform + ELU) on the order-2 layer of a
BASELINE configuration; PPG_GCN_TC=staged|single|ws selects the kernel (default staged).
Checks the two against each other when run with --check (needs both results, so it calls itself)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import pathpyg_b200 as pp  # noqa: E402
from pathpyg_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--layer", type=int, default=None, help="order of the layer (default: the highest one)")
    ap.add_argument("--save", default=None)
    ap.add_argument("--compare", default=None)
    a = ap.parse_args()
    cfg = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    ei, t = bench.make_stream(cfg, seed=0)
    tg = pp.TemporalGraph.from_tensors(ei.to(dev), t.to(dev), cfg["n"])
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=cfg["delta"], max_order=cfg["order"])
    layer = model.layers[a.layer or cfg["order"]]
    H = cfg["hidden"]
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(layer.n, H, generator=gen).to(dev)
    w = (torch.randn(H, H, generator=gen) / 8).to(dev)
    b = torch.randn(H, generator=gen).to(dev)
    graph = ops.gcn_prepare(layer.data.edge_index.as_tensor(), layer.data.edge_weight, layer.n)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        out = ops.gcn_layer_tc(graph, x, w, b, _lib.ACT_ELU)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = ops.gcn_layer_tc(graph, x, w, b, _lib.ACT_ELU)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    esl = layer.m + layer.n
    alg = 20 * esl + 4 * H * esl + 12 * H * layer.n
    fused_min = 8 * layer.m + 8 * layer.n + 8 * H * layer.n
    med = times[len(times) // 2]
    which = {"single": "single-role", "ws": "warp-specialised"}.get(os.environ.get("PPG_GCN_TC", "staged"), "staged")
    if os.environ.get("PPG_GCN_TC_WS") == "1":
        which = "warp-specialised"
    print(f"gcn_tc ({which}) n={layer.n} e={layer.m} F=H={H}: median {med * 1e3:.1f} us, min {times[0] * 1e3:.1f} us; "
          f"SURVEY bytes {alg / med / 1e6:.0f} GB/s, fused-min bytes {fused_min / med / 1e6:.0f} GB/s")
    if a.save:
        torch.save(out.cpu(), a.save)
    if a.compare:
        other = torch.load(a.compare)
        err = ((out.cpu() - other).abs() / other.abs().clamp(min=1.0)).max().item()
        print(f"max rel deviation between the two kernels: {err:.3e}; bit-identical: {torch.equal(out.cpu(), other)}")


if __name__ == "__main__":
    main()
