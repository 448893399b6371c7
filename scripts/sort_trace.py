"""Phase timeline of one onesweep digit pass (needs a library built with -DPPG_SORT_TRACE -rdc=true; development aid)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathpyg_b200 import _lib, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
bits = 8  # one pass
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.ppg_debug_set_sort_trace.argtypes = [ctypes.c_void_p]
items = int(os.environ.get("PPG_SORT_ITEMS", 8 if n <= (512 << 10) else 16))
tiles = -(-n // (256 * items))
trace = torch.zeros(tiles * 8, dtype=torch.int64, device=dev)
g = torch.Generator().manual_seed(0)
base = torch.randint(0, 1 << 40, (n,), generator=g).to(dev)
for it in range(3):
    keys = base.clone()
    torch.cuda.synchronize()
    assert lib.ppg_debug_set_sort_trace(ctypes.c_void_p(trace.data_ptr())) == 0
    ops.sort_pairs_u64(keys, bits)
    torch.cuda.synchronize()
t = trace.view(tiles, 8).cpu().double()
t0 = t[:, 0].min()
names = ["start", "loads issued", "ranked", "look-back done", "scans done", "reordered", "written"]
print(f"n={n} tiles={tiles}: per-phase mean duration (us), and absolute time of phase end (us since first tile start): min / mean / max")
for i in range(7):
    d = (t[:, i] - t[:, i - 1]) / 1e3 if i else t[:, 0] * 0
    a = (t[:, i] - t0) / 1e3
    print(f"  {names[i]:>15}: dur {d.mean():7.2f}   abs {a.min():7.2f} / {a.mean():7.2f} / {a.max():7.2f}")
order = torch.argsort(t[:, 0])
print("  first/last tile to start:", int(order[0]), int(order[-1]), " span of kernel (us):", float((t[:, 6].max() - t0) / 1e3))
