"""Per-stage CUDA-event timings of the hot path at a BASELINE configuration (development aid)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathpyg_b200 as pp  # noqa: E402
from pathpyg_b200 import ops  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    out = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--delta", type=int, default=200)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--hidden", type=int, default=64)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    ei = torch.randint(0, a.n, (2, a.m), generator=g).to(dev)
    t = torch.sort(torch.randint(0, a.T, (a.m,), generator=g)).values.to(dev)
    tg = pp.TemporalGraph.from_tensors(ei, t, a.n)

    ms, ev = timed(lambda: ops.lift_order_temporal(ei, t, a.delta, a.n))
    E2 = ev.size(1)
    print(f"lift_order_temporal      m={a.m} -> E2={E2}: {ms:.3f} ms  ({(24*a.m+16*E2)/ms/1e6:.1f} GB/s alg)")
    ns2 = ei.t().contiguous()
    ms, (u, inv) = timed(lambda: ops.unique_rows(ns2))
    print(f"unique_rows [m,2]        -> n2={u.size(0)}: {ms:.3f} ms")
    ms, (cei, cw) = timed(lambda: ops.coalesce(ev, inv, u.size(0), None))
    print(f"coalesce E2              -> {cei.size(1)}: {ms:.3f} ms")
    ms, l3 = timed(lambda: ops.lift_order_edge_index(ev, a.m))
    print(f"lift_order_edge_index    E2={E2} -> E3={l3.size(1)}: {ms:.3f} ms  ({(16*E2+16*l3.size(1))/ms/1e6:.1f} GB/s alg)")
    ms, model = timed(lambda: pp.MultiOrderModel.from_temporal_graph(tg, delta=a.delta, max_order=a.order), reps=3)
    lifted = E2 if a.order == 2 else None
    print(f"from_temporal_graph K={a.order}: {ms:.3f} ms; layers {[(k, v.n, v.m) for k, v in model.layers.items()]}")

    H = a.hidden
    n2 = model.layers[a.order].n
    model.layers[1].data.x = torch.randn(a.n, H, device=dev)
    data = model.to_dbgnn_data(max_order=a.order, x_h=torch.randn(n2, H, device=dev))
    net = pp.nn.DBGNN(num_classes=16, num_features=(H, H), hidden_dims=[H, H, H]).to(dev).eval()
    with torch.no_grad():
        ms, out = timed(lambda: net(data))
    print(f"DBGNN({H}) forward: {ms:.3f} ms -> {(a.n + n2)/ms/1e3:.2f} M nodes/s")
    with torch.no_grad():
        gr = ops.gcn_prepare(data.edge_index_higher_order, data.edge_weights_higher_order, n2)
        ms, _ = timed(lambda: ops.gcn_prepare(data.edge_index_higher_order, data.edge_weights_higher_order, n2))
        print(f"  gcn_prepare HO: {ms:.3f} ms")
        ms, agg = timed(lambda: ops.spmm_csc(gr, data.x_h))
        print(f"  spmm HO: {ms:.3f} ms")
        ms, _ = timed(lambda: ops.linear(agg, net.higher_order_layers[0].lin.weight, net.higher_order_layers[0].bias, 1))
        print(f"  linear HO: {ms:.3f} ms")


if __name__ == "__main__":
    main()
