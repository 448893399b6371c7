"""Whole-sort CUDA-event time of ppg_sort_pairs_u64 at several sizes (development aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from pathpyg_b200 import ops
    dev = torch.device("cuda", 0)
    sizes = [int(s) for s in sys.argv[1].split(",")]
    bits = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    for n in sizes:
        base = torch.randint(0, 1 << bits, (n,), generator=g).to(dev)
        best = 1e9
        for it in range(8):
            keys = base.clone()
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.sort_pairs_u64(keys, bits)
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, a.elapsed_time(b))
        assert bool((keys[1:] >= keys[:-1]).all())
        print(f"n={n:>9} bits={bits}: {best * 1e3:.1f} us")


if __name__ == "__main__":
    main()
