"""Distributed temporal lift (cfg5 style: one time-sorted stream split across the ranks, ghost-zone exchange, k-gram
keys resolved at their owner with all-to-all-v) timed under torchrun; rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_lift_bench.py --events 20000000 --order 3
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathpyg_b200 as pp  # noqa: E402
from pathpyg_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=int, default=20_000_000)
    ap.add_argument("--nodes", type=int, default=400_000)
    ap.add_argument("--horizon", type=int, default=1_000)
    ap.add_argument("--delta", type=int, default=20)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--profile", action="store_true", help="rank 0 prints a torch.profiler table of one step")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator().manual_seed(0)                       # every rank draws the same stream, keeps its range
    ei = torch.randint(0, a.nodes, (2, a.events), generator=gen)
    t = torch.sort(torch.randint(0, a.horizon, (a.events,), generator=gen)).values
    lo, hi = parallel.partition_stream(a.events, rank, world)
    ei_l, t_l = ei[:, lo:hi].to(dev).contiguous(), t[lo:hi].to(dev).contiguous()

    def step():
        return parallel.distributed_temporal_layers(ei_l, t_l, a.nodes, a.delta, a.order)

    layers = step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    times = []
    for _ in range(a.steps):
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        layers = step()
        e.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms))
    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        dist.barrier()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize(dev)
        if rank == 0:
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70), file=sys.stderr)
    sizes = {}
    for k, layer in layers.items():
        cnt = torch.tensor([layer.node_sequence.size(0), layer.edge_index.size(1)], device=dev)
        dist.all_reduce(cnt)
        sizes[k] = cnt.tolist()
    single = None
    if rank == 0 and a.events <= 30_000_000:                          # the same stream on one GPU, for the ratio
        tg = pp.TemporalGraph.from_tensors(ei.to(dev), t.to(dev), a.nodes)
        pp.MultiOrderModel.from_temporal_graph(tg, delta=a.delta, max_order=a.order)
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        pp.MultiOrderModel.from_temporal_graph(tg, delta=a.delta, max_order=a.order)
        e.record()
        torch.cuda.synchronize(dev)
        single = s.elapsed_time(e)
    if rank == 0:
        best = min(times)
        print(json.dumps({"what": "distributed_temporal_layers", "n_gpus": world, "events": a.events, "nodes": a.nodes, "delta": a.delta,
                          "max_order": a.order, "layers_nodes_edges": sizes, "ms_best": best, "ms_all": times,
                          "events_per_s": a.events / best * 1e3, "single_gpu_ms_same_stream": single}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
