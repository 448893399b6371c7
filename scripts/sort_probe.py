"""Per-pass CUDA-event timing of the onesweep radix sort at several sizes (development aid)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathpyg_b200 import ops  # noqa: E402


def hbm_peak_gbs() -> float:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        return 6455.9


def main():
    dev = torch.device("cuda", 0)
    peak = hbm_peak_gbs()
    sizes = [int(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1000000,1800000,4000000,20000000,64000000".split(","))]
    bits = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    for n in sizes:
        base = torch.randint(0, 1 << bits, (n,), generator=g)
        if os.environ.get("PROBE_SORTED"):  # already ordered input: the high digits are uniform across a warp
            base = torch.sort(base).values
        base = base.to(dev)
        per = []
        for it in range(6):
            keys = base.clone()
            flush.fill_(1)
            _, ms = ops.sort_pairs_u64(keys, bits, time_passes=True)
            if it >= 2:
                per.append(ms)
        assert bool((keys[1:] >= keys[:-1]).all())
        avg = [sum(p[i] for p in per) / len(per) for i in range(len(per[0]))]
        mean = sum(avg) / len(avg)
        print(f"n={n:>9} bits={bits} passes={len(avg)} per-pass us: {[round(a * 1e3, 1) for a in avg]}  mean {mean * 1e3:.1f} us "
              f"-> {24 * n / mean / 1e6:.0f} GB/s ({24 * n / mean / 1e6 / peak:.3f} of measured HBM peak)")


if __name__ == "__main__":
    main()
