"""Phase table of one distributed lift step (cfg5 by default) under torchrun: CUDA-event time stamps of every phase of
`parallel.distributed_temporal_layers` on every rank (PPG_DIST_TRACE=1), printed for rank 0 and as max over ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_trace.py [--events ...]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PPG_DIST_TRACE"] = "1"
import bench  # noqa: E402
from pathpyg_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--events", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--last-share", type=float, default=None, help="size of the last rank's range relative to the others")
    a = ap.parse_args()
    cfg = dict(bench.WORKLOADS[a.workload])
    if a.events:
        cfg["n"] = max(1, cfg["n"] * a.events // cfg["m"])
        cfg["m"] = a.events
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29512")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ei, t = bench.make_stream(cfg, seed=0)
    lo, hi = parallel.partition_stream(cfg["m"], rank, world, last_share=a.last_share if a.last_share is not None else 1.0 + bench.ghost_overhead(cfg, world))
    ei_l, t_l = ei[:, lo:hi].contiguous().to(dev), t[lo:hi].contiguous().to(dev)
    del ei, t
    K = cfg["order"]
    best = None
    for i in range(a.steps):
        dist.barrier()
        torch.cuda.synchronize()
        layers = parallel.distributed_temporal_layers(ei_l, t_l, cfg["n"], cfg["delta"], K)
        del layers
        tr = parallel.last_trace
        total = torch.tensor([tr[-1][1]], device=dev)
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
        if best is None or float(total) < best[0]:
            best = (float(total), tr)
    total, tr = best
    durations = torch.tensor([b[1] - a_[1] for a_, b in zip(tr[:-1], tr[1:])], device=dev)
    worst = durations.clone()
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    every = [torch.empty_like(durations) for _ in range(world)]
    dist.all_gather(every, durations)
    if rank == 0:
        print(f"# distributed lift {a.workload}: {cfg['m']} events, {world} rank(s); best step {total:.2f} ms (max over ranks)")
        print(f"# {'phase':24s} {'rank0 ms':>10s} {'max ms':>10s}")
        agg = {}
        for (label, _), d0, dmax in zip(tr[:-1], durations.tolist(), worst.tolist()):
            print(f"  {label:24s} {d0:10.3f} {dmax:10.3f}")
            key = label.split("[")[0]
            agg[key] = agg.get(key, 0.0) + dmax
        print("# by phase kind (sum of per-phase max over ranks):", json.dumps({k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}))
        print(f"# peak memory rank 0: {torch.cuda.max_memory_allocated(dev) / 2**30:.1f} GiB")
        kinds = sorted({label.split("[")[0] for label, _ in tr[:-1]})
        print("# per rank, ms by phase kind (a rank that waits at a synchronisation shows it under route_count / merge_sync):")
        for r, d in enumerate(every):
            row = {k: 0.0 for k in kinds}
            for (label, _), v in zip(tr[:-1], d.tolist()):
                row[label.split("[")[0]] += v
            print(f"#   rank {r}: " + "  ".join(f"{k} {v:.1f}" for k, v in row.items()))
        print("# expansion of every level per rank: lift_next ms")
        for r, d in enumerate(every):
            print(f"#   rank {r}: " + "  ".join(f"{label} {v:.2f}" for (label, _), v in zip(tr[:-1], d.tolist()) if label.startswith("lift_next")))
    sizes = [None] * world
    dist.all_gather_object(sizes, parallel.last_sizes)
    if rank == 0:
        print("# (level, sources, sources expanded, pairs) per rank:")
        for r, sz in enumerate(sizes):
            print(f"#   rank {r}: {sz}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
