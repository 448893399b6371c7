"""A/B of the two layer chains of MultiOrderModel.from_temporal_graph on one GPU: generation-order tiles (csrc/chain.cu)
against one radix sort per order (PPG_CHAIN=0).  Device time (CUDA events), best and median of `reps` builds."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import pathpyg_b200 as pp  # noqa: E402
from pathpyg_b200 import _lib  # noqa: E402


def main():
    names = sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    for name in names:
        cfg = bench.WORKLOADS[name]
        ei, t = bench.make_stream(cfg, seed=0)
        tg = pp.TemporalGraph.from_tensors(ei.to(dev), t.to(dev), cfg["n"])
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        for mode in ("1", "0"):
            os.environ["PPG_CHAIN"] = mode
            reps = 12 if cfg["m"] <= 10_000_000 else 4
            ms, launches = [], 0
            for i in range(reps + 2):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0 = lib.ppg_launch_count()
                a.record()
                model = pp.MultiOrderModel.from_temporal_graph(tg, delta=cfg["delta"], max_order=cfg["order"])
                b.record()
                torch.cuda.synchronize()
                launches = lib.ppg_launch_count() - c0
                if i >= 2:
                    ms.append(a.elapsed_time(b))
                sizes = {k: (g.n, g.m) for k, g in model.layers.items()}
                del model
            print(f"{name} chain={'tiles' if mode == '1' else 'sort '}: best {min(ms):9.3f} ms  median {statistics.median(ms):9.3f} ms  "
                  f"launches {launches}  peak {torch.cuda.max_memory_allocated(dev) / 2**30:.1f} GiB  layers {sizes}", flush=True)
            torch.cuda.reset_peak_memory_stats(dev)
        del tg
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
