"""Phase timeline of the fused tcgen05 GCN layer on the order-2 graph of cfg2 (needs the instrumented library:
`make -C pathpyg_b200/csrc trace`, then PATHPYG_B200_LIB=pathpyg_b200/_C/libpathpyg_b200_trace.so; development aid)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathpyg_b200 as pp  # noqa: E402
from pathpyg_b200 import _lib, ops  # noqa: E402

n, m, T, delta, H = 100_000, 1_000_000, 1000, 200, 64
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.ppg_debug_set_sort_trace.argtypes = [ctypes.c_void_p]
g = torch.Generator().manual_seed(0)
ei = torch.randint(0, n, (2, m), generator=g).to(dev)
t = torch.sort(torch.randint(0, T, (m,), generator=g)).values.to(dev)
model = pp.MultiOrderModel.from_temporal_graph(pp.TemporalGraph.from_tensors(ei, t, n), delta=delta, max_order=2)
layer = model.layers[2]
n2 = layer.n
grouped = ops.gcn_prepare(layer.data.edge_index, layer.data.edge_weight, n2)
x = torch.randn(n2, H, device=dev)
w = torch.randn(H, H, device=dev) / 8
b = torch.zeros(H, device=dev)
tiles = -(-n2 // 128)
trace = torch.zeros(tiles * 8, dtype=torch.int64, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for it in range(3):
    flush.fill_(1)
    torch.cuda.synchronize()
    assert lib.ppg_debug_set_sort_trace(ctypes.c_void_p(trace.data_ptr())) == 0
    ops.gcn_layer_tc(grouped, x, w, b, 1)
    torch.cuda.synchronize()
    assert lib.ppg_debug_set_sort_trace(ctypes.c_void_p(0)) == 0
tr = trace.view(tiles, 8).cpu().double()
t0 = tr[:, 0].min()
variant = os.environ.get("PPG_GCN_TC", "staged")
if variant in ("single", "ws"):
    names = ["tile start", "rows gathered", "product in TMEM", "tile written"]
else:   # staged kernel: time stamps of consumer thread 0 (slots 0-3) and of lane 0 of the first producer warp (4, 5)
    names = ["tile start (pointers visible)", "rows reduced", "operands handed over", "previous tile written"]
print(f"[{variant}] n2={n2} e2={layer.m} tiles={tiles}: per-phase mean duration (us); absolute time of phase end (us since first tile start): min / mean / max")
for i in range(len(names)):
    d = (tr[:, i] - tr[:, i - 1]) / 1e3 if i else tr[:, 0] * 0
    a = (tr[:, i] - t0) / 1e3
    print(f"  {names[i]:>36}: dur {d.mean():7.2f} (p10 {d.quantile(0.1):6.2f}, p90 {d.quantile(0.9):6.2f})   abs {a.min():8.2f} / {a.mean():8.2f} / {a.max():8.2f}")
last = len(names) - 1
ctas = 148 if variant not in ("single",) else 296
print("  span of kernel (us):", float((tr[:, last].max() - t0) / 1e3), " tiles per CTA:", tiles / ctas)
if variant not in ("single", "ws"):
    ok = (tr[:, 4] > 0) & (tr[:, 5] > 0)
    lead = (tr[ok, 0] - tr[ok, 4]) / 1e3      # consumer starts the tile this long after the producer acquired its first stage
    done = (tr[ok, 1] - tr[ok, 5]) / 1e3      # consumer finishes reducing this long after the last chunk was REQUESTED
    span = (tr[ok, 5] - tr[ok, 4]) / 1e3      # producer: first stage acquired -> last chunk requested
    for nm, x in (("producer ahead of consumer at tile start", lead), ("last request -> rows reduced", done), ("producer: first stage -> last request", span)):
        print(f"  {nm:>44}: mean {x.mean():7.2f}  p10 {x.quantile(0.1):7.2f}  p50 {x.quantile(0.5):7.2f}  p90 {x.quantile(0.9):7.2f}")
