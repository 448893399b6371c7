"""Benchmark of the lift -> DBGNN hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3] [--impl reference]

Workload (N = 1 headline) = BASELINE.json configs[1]: synthetic temporal ER graph, 100k nodes, 1M
time-stamped edges (T = 1000 distinct time stamps, delta = 200, seed 0 + rank), order-2 lift via
``MultiOrderModel.from_temporal_graph`` followed by a DBGNN(hidden 64-64-64, 16 classes, dense
64-wide features) forward.  One step = one pass of that path over one such graph.

Primary metric : k-order lift edges/s  (lifted edges = event-graph columns E_2 (+ E_3.. for K > 2))
Secondary      : DBGNN forward nodes/s (nodes = N + n_K), reported under "dbgnn" in the same line.

Timing: CUDA events on the stream the kernels are launched on, around every step; L2 (126 MB) is
flushed between steps by writing a 512 MB buffer (not timed); max over ranks of the summed step times.
Multi-GPU: one process per GPU, every rank lifts its own independent graph of the same size (weak
scaling, no data-path collective: SURVEY.md 8e "independent units"); NCCL carries only the barrier
and the max-reduction of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nodes, edges, distinct time stamps, delta, max_order, hidden, classes)
    "cfg2": dict(n=100_000, m=1_000_000, T=1000, delta=200, order=2, hidden=64, classes=16,
                 label="synthetic temporal ER N=100k, 1M timestamped edges, order-2 lift + DBGNN(64) forward"),
    "cfg3": dict(n=100_000, m=10_000_000, T=250, delta=5, order=3, hidden=64, classes=16,
                 label="synthetic 10M timestamped edges, delta=5 causal-path extraction + order-3 lift"),
}


def make_stream(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, cfg["n"], (2, cfg["m"]), generator=g)
    t = torch.sort(torch.randint(0, cfg["T"], (cfg["m"],), generator=g)).values
    return ei, t


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons during the timed region (NVML, 20 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # NVML missing: report nulls rather than fail the bench
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def lifted_edges(model, cfg, e2):
    return e2  # K = 2: the event graph; higher orders add their line-graph columns (see run_gpu)


def alg_bytes_lift(cfg, m, e2, layers):
    """a1 24m + 16 E2; a4 at order k: 8(k+1) E_{k-1} + 20 E_k + 8k n_k + 20 E^_k."""
    n1, eh1 = layers[1]
    total = 24 * m + 16 * e2
    total += 8 * 2 * cfg["n"] + 20 * m + 8 * 1 * n1 + 20 * eh1         # layer 1 (rows = arange(N))
    if 2 in layers:
        n2, eh2 = layers[2]
        total += 8 * 3 * m + 20 * e2 + 8 * 2 * n2 + 20 * eh2            # layer 2
    return total


def alg_bytes_dbgnn(n, e, n2, e2, H, classes):
    """SURVEY.md 8d per GCN layer: 20 e_sl + 4H e_sl + 12 H n; bipartite: 16 nK + 4H nK + 8H (nK + N) + 4H N."""
    def gcn(nn, ee):
        esl = ee + nn
        return 20 * esl + 4 * H * esl + 12 * H * nn
    return 2 * gcn(n, e) + 2 * gcn(n2, e2) + 16 * n2 + 4 * H * n2 + 8 * H * (n2 + n) + 4 * H * n + 4 * n * (H + classes)


def run_gpu(args, cfg, rank, world, local_rank):
    import pathpyg_b200 as pp
    from pathpyg_b200 import _lib

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    ei_h, t_h = make_stream(cfg, seed=rank)
    ei_pin, t_pin = ei_h.pin_memory(), t_h.pin_memory()
    ei, t = ei_h.to(dev), t_h.to(dev)
    H, K = cfg["hidden"], cfg["order"]
    gen = torch.Generator().manual_seed(1000 + rank)
    x = torch.randn(cfg["n"], H, generator=gen).to(dev)
    net = pp.nn.DBGNN(num_classes=cfg["classes"], num_features=(H, H), hidden_dims=[H, H, H]).to(dev).eval()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    state = {}

    def dbgnn_step(model):
        model.layers[1].data.x = x
        nK = model.layers[K].n
        if state.get("x_h") is None or state["x_h"].size(0) != nK:
            state["x_h"] = torch.randn(nK, H, generator=torch.Generator().manual_seed(7)).to(dev)
        data = model.to_dbgnn_data(max_order=K, x_h=state["x_h"])
        with torch.no_grad():
            return net(data)

    resident = pp.TemporalGraph.from_tensors(ei, t, cfg["n"])  # the input container, resident in HBM before the timed region

    def one_step(timed):
        flush.fill_(1)  # evict L2 between steps
        e0, e1, e2_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0 = lib.ppg_launch_count()
        e0.record(stream)
        model = pp.MultiOrderModel.from_temporal_graph(resident, delta=cfg["delta"], max_order=K)
        e1.record(stream)
        c1 = lib.ppg_launch_count()
        out = dbgnn_step(model)
        e2_.record(stream)
        c2 = lib.ppg_launch_count()
        if timed is not None:
            timed.append((e0, e1, e2_, c1 - c0, c2 - c1))
        return model, out

    for _ in range(args.warmup):
        model, out = one_step(None)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    layers = {k: (g.n, g.m) for k, g in model.layers.items()}
    # lifted edges per step: event-graph columns, plus line-graph columns of every further order
    e2 = int(pp.algorithms.lift_order_temporal(pp.TemporalGraph.from_tensors(ei, t, cfg["n"]), cfg["delta"]).size(1))
    lifted = e2
    if K > 2:
        idx = pp.algorithms.lift_order_temporal(pp.TemporalGraph.from_tensors(ei, t, cfg["n"]), cfg["delta"])
        num = cfg["m"]
        for _ in range(3, K + 1):
            nxt = pp.algorithms.lift_order_edge_index(idx, num)
            num, idx = idx.size(1), nxt
            lifted += int(idx.size(1))
        del idx
    nodes = cfg["n"] + layers[K][0]

    timed = []
    torch.cuda.synchronize(dev)
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            one_step(timed)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - wall0
    if world > 1:
        torch.distributed.barrier()
    lift_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in timed)
    dbgnn_ms = sum(b.elapsed_time(c) for _, b, c, _, _ in timed)
    launches = timed[-1][3] + timed[-1][4]

    # ---- e2e: the public call with HOST (pinned) inputs; H2D of the inputs and D2H of the result inside the timed region
    out_pin = torch.empty((cfg["n"], cfg["classes"]), dtype=torch.float32).pin_memory()   # host buffer the activations land in

    host_graph = pp.TemporalGraph.from_tensors(ei_pin, t_pin, cfg["n"])   # the caller's graph: pinned HOST tensors

    def e2e_step():
        flush.fill_(1)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(stream)
        # public call on the host graph: uploads the edge index and the time stamps (24 B per event) itself
        model_ = pp.MultiOrderModel.from_temporal_graph(host_graph, delta=cfg["delta"], max_order=K, device=dev)
        sizes = torch.tensor([g.m for g in model_.layers.values()], device=dev).cpu()  # result read-back of the lift
        b.record(stream)
        out_h = out_pin.copy_(dbgnn_step(model_), non_blocking=True)  # result read-back of the forward
        c.record(stream)
        return a, b, c, out_h, sizes

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize(dev)
    rec = [e2e_step() for _ in range(e2e_steps)]
    torch.cuda.synchronize(dev)
    e2e_lift_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in rec) / e2e_steps
    e2e_dbgnn_ms = sum(b.elapsed_time(c) for _, b, c, _, _ in rec) / e2e_steps
    h2d = ei_pin.numel() * 8 + t_pin.numel() * 8
    d2h_lift, d2h_dbgnn = 8 * len(layers), cfg["n"] * cfg["classes"] * 4

    # ---- dominant kernel, timed live with CUDA events on the launch stream: one onesweep digit pass
    roofline = measure_sort_pass(pp, dev, e2, layers[K][0])
    at_scale = measure_sort_pass(pp, dev, 64_000_000, 1 << 20) if rank == 0 else None

    times = torch.tensor([lift_ms, dbgnn_ms, e2e_lift_ms, e2e_dbgnn_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    lift_ms, dbgnn_ms, e2e_lift_ms, e2e_dbgnn_ms = times.tolist()

    if rank != 0:
        return
    peaks = load_peaks()
    steps = args.steps
    lift_s = lift_ms / 1e3
    value = world * lifted * steps / lift_s
    lift_bytes = alg_bytes_lift(cfg, cfg["m"], e2, layers)
    db_bytes = alg_bytes_dbgnn(cfg["n"], layers[1][1], layers[K][0], layers[K][1], H, cfg["classes"])
    line = {
        "metric": "k-order lift edges/s",
        "value": value,
        "unit": "lifted edges/s",
        "n_gpus": world,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": (lift_ms + dbgnn_ms) / steps,
        "ms_per_step_lift": lift_ms / steps,
        "ms_per_step_dbgnn": dbgnn_ms / steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64 indices (int32/u64 keys inside) + fp32 weights/activations",
        "data": "synthetic",
        "config": {"workload": cfg["label"], "name": args.workload, "nodes": cfg["n"], "edges": cfg["m"], "timestamps": cfg["T"],
                   "delta": cfg["delta"], "max_order": K, "lifted_edges_per_step": lifted, "layers": {str(k): v for k, v in layers.items()},
                   "dbgnn": {"hidden_dims": [H, H, H], "classes": cfg["classes"], "features": "dense randn fp32"},
                   "l2": "flushed between steps (512 MB write)", "parallelism": f"{world} independent graphs (one per GPU)"},
        "dbgnn": {"metric": "DBGNN forward nodes/s", "value": world * nodes * steps / (dbgnn_ms / 1e3), "unit": "nodes/s",
                  "nodes_per_step": nodes, "ms_per_step": dbgnn_ms / steps,
                  "roofline": stage_roofline(db_bytes, dbgnn_ms / steps, peaks),
                  "e2e": {"value": world * nodes / (e2e_dbgnn_ms / 1e3), "unit": "nodes/s", "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": d2h_dbgnn}},
        "lift_stage_roofline": stage_roofline(lift_bytes, lift_ms / steps, peaks),
        "roofline": roofline_entry(roofline, peaks, traffic_key="pairs_1.8M"),
        "roofline_at_scale": roofline_entry(at_scale, peaks, traffic_key="pairs_64M"),
        "e2e": {"value": world * lifted / (e2e_lift_ms / 1e3), "unit": "lifted edges/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h_lift, "ms_per_step": e2e_lift_ms},
        "gpu_launches": launches,
        "gpu_launches_lift": timed[-1][3],
        "gpu_launches_dbgnn": timed[-1][4],
        "wall_s_timed_region": wall,
        "clocks": clocks.summary(),
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, budget_s=args.cpu_budget)
    print(json.dumps(line))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


def stage_roofline(alg_bytes, ms, peaks):
    achieved = alg_bytes / (ms / 1e3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "algorithmic_bytes": alg_bytes, "peak_source": peaks["source"]}


def measure_sort_pass(pp, dev, num_pairs, num_nodes):
    """Dominant kernel of the lift: the onesweep digit pass over (u64 key, u32 payload) pairs.  Its launches
    are timed live with CUDA events on the launching stream (ppg_sort_pairs_u64 records an event around every
    pass) on keys shaped like the order-K coalesce: num_pairs keys of 2*ceil(log2 num_nodes) bits; L2 is
    flushed before every sort, so the first pass reads cold keys."""
    from pathpyg_b200 import ops
    g = torch.Generator().manual_seed(5)
    bits = 2 * max(num_nodes - 1, 1).bit_length()
    base = (torch.randint(0, num_nodes, (num_pairs,), generator=g) << (bits // 2)) | torch.randint(0, num_nodes, (num_pairs,), generator=g)
    base = base.to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    per_pass = []
    for it in range(8):
        keys = base.clone()
        flush.fill_(1)
        _, ms = ops.sort_pairs_u64(keys, bits, time_passes=True)
        if it >= 3:
            per_pass.extend(ms)
    assert bool((keys[1:] >= keys[:-1]).all()), "sort probe produced unsorted keys"
    return {"kernel": "onesweep_pass_kernel<u64 key, u32 payload>", "pairs": num_pairs, "passes": len(ms),
            "launch_ms": sum(per_pass) / len(per_pass), "bytes_per_launch": 2 * 12 * num_pairs}


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the probe kernel, from the committed
    `ncu --set full` capture of the same probe (profiles/r01_sort_traffic.json); None if not captured."""
    path = os.path.join(ROOT, "profiles", "r01_sort_traffic.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        return json.load(f).get(key, {}).get("dram_bytes_per_launch")


def roofline_entry(r, peaks, traffic_key=None):
    achieved = r["bytes_per_launch"] / (r["launch_ms"] / 1e3) / 1e9
    return {"bound": "hbm", "kernel": r["kernel"], "pairs": r["pairs"], "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(traffic_key) if traffic_key else None,
            "algorithmic_bytes_per_launch": r["bytes_per_launch"],
            "launch_ms": r["launch_ms"], "launches_averaged": r["passes"] * 5,
            "how": "CUDA events around every digit pass on the launch stream; bytes = read + write of (8 B key + 4 B payload) per pair",
            "peak_source": peaks["source"]}


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(cfg, budget_s, threads):
    """The reference's CPU implementation of the path (oracle port, same operation order) on a bounded
    sample of the workload: ``lift_order_temporal``'s loop over the first S distinct time stamps as sources
    against the FULL stream (so the per-time-stamp work is exactly the full run's), then
    ``aggregate_edge_index`` of the order-2 layer for the event edges produced."""
    from oracle import lift

    torch.set_num_threads(threads)
    ei, t = make_stream(cfg, seed=0)
    stamps = torch.unique(t)
    delta = torch.tensor(cfg["delta"])
    pos = torch.arange(ei.size(1))
    t0 = time.perf_counter()
    pieces, used = [], 0
    for ts in stamps:  # oracle/lift.py::lift_order_temporal body, stopped after the time budget
        heads = pos[t == ts]
        tails = pos[(t > ts) & (t <= ts + delta)]
        if heads.numel() and tails.numel():
            pairs = torch.cartesian_prod(heads, tails)
            pieces.append(pairs[ei[1, pairs[:, 0]] == ei[0, pairs[:, 1]]])
        used += 1
        if time.perf_counter() - t0 > budget_s:
            break
    ho = torch.cat(pieces, dim=0).t().contiguous()
    t_lift = time.perf_counter() - t0
    t1 = time.perf_counter()
    ns = ei.t().contiguous()
    lift.aggregate_edge_index(ho, ns, None)
    t_agg = time.perf_counter() - t1
    edges = int(ho.size(1))
    return {"edges": edges, "seconds": t_lift + t_agg, "stamps": used, "of_stamps": int(stamps.numel()),
            "t_lift": t_lift, "t_aggregate": t_agg}


def cpu_baseline(cfg, budget_s):
    threads = os.cpu_count() or 1
    s = cpu_reference_sample(cfg, budget_s, threads)
    return {"value": s["edges"] / s["seconds"], "unit": "lifted edges/s", "cores": threads, "kind": "port",
            "sample": f"oracle port of lift_order_temporal over the first {s['stamps']} of {s['of_stamps']} source time stamps against the "
                      f"full {cfg['m']}-edge stream ({s['t_lift']:.1f} s) + aggregate_edge_index of the order-2 layer on the "
                      f"{s['edges']} event edges produced ({s['t_aggregate']:.1f} s)"}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step_budget = max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_sample(cfg, per_step_budget, threads)
    edges, seconds, last = 0, 0.0, None
    for _ in range(args.steps):
        last = cpu_reference_sample(cfg, per_step_budget, threads)
        edges += last["edges"]
        seconds += last["seconds"]
    value = edges / seconds
    sample = (f"per step: oracle port (reference operation order, torch CPU, {threads} threads) of lift_order_temporal over the first "
              f"{last['stamps']} of {last['of_stamps']} source time stamps against the full stream + aggregate_edge_index of the order-2 layer")
    print(json.dumps({
        "impl": "reference", "metric": "k-order lift edges/s", "value": value, "unit": "lifted edges/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64 + fp32 (torch CPU)", "data": "synthetic",
        "config": {"workload": cfg["label"], "name": args.workload, "nodes": cfg["n"], "edges": cfg["m"], "timestamps": cfg["T"],
                   "delta": cfg["delta"], "max_order": 2},
        "cpu_baseline": {"value": value, "unit": "lifted edges/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "lifted edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps on the GPU arm
    cfg = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if world > 1:
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu(args, cfg, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
