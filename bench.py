"""Benchmark of the lift -> DBGNN hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5] [--impl reference] [--no-extras]

Workloads (BASELINE.json `configs`; synthetic parametrisations of SURVEY.md 8d, seed 0):
  cfg2  temporal ER, N=100k nodes, 1M events (T=1000 stamps, delta=200), order-2 lift + DBGNN(64) forward     [N=1 headline]
  cfg3  10M events (T=250, delta=5), causal-path extraction + order-3 lift
  cfg4  order-2 DBGNN(32) TRAINING on 5M walks sharded by walk id, one NCCL all-reduce of the gradients per step
  cfg5  ONE 50M-event stream (N=1M nodes, T=2500, delta=100), orders 1-5, split over the ranks by time range;
        per order one all-to-all of 16-byte records (cross-partition lifted edges) + 4-byte ids back         [N>1 headline]

Default workload: cfg2 on one GPU; cfg5 (strong scaling: the SAME stream on more GPUs) under torchrun with
WORLD_SIZE > 1 -- there the line also carries the time of the same stream on ONE GPU measured in the same run
(`strong_scaling`) and `parity_ok`: 64-bit order-sensitive digests of every layer, summed over the ranks, equal the
digests of the single-GPU layers.  The other workloads are reported under `extra` (skip with --no-extras).

Primary metric : k-order lift edges/s  (lifted edges = line-graph columns E_2 + ... + E_K)
Secondary      : DBGNN forward nodes/s (nodes = N + n_K), under "dbgnn".

Timing: CUDA events on the launching stream around every step, barrier + synchronize on both sides of the timed
region, max over ranks.  cfg2: L2 (126 MB) flushed between steps by a 512 MB write (not timed); cfg3/cfg4/cfg5:
inputs and intermediates are many times larger than L2.
"""
from __future__ import annotations

import argparse
import ctypes
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
    "cfg2": dict(kind="lift", n=100_000, m=1_000_000, T=1000, delta=200, order=2, hidden=64, classes=16, dbgnn=True,
                 label="synthetic temporal ER N=100k, 1M timestamped edges, order-2 lift + DBGNN(64) forward"),
    "cfg3": dict(kind="lift", n=100_000, m=10_000_000, T=250, delta=5, order=3, hidden=64, classes=16, dbgnn=False,
                 label="synthetic 10M timestamped edges, delta=5 causal-path extraction + order-3 lift"),
    "cfg4": dict(kind="train", n=100_000, walks=5_000_000, min_len=3, max_len=11, order=2, hidden=32, classes=16,
                 label="order-2 DBGNN training, 32-dim, 5M paths sharded by walk id, NCCL grad all-reduce"),
    "cfg5": dict(kind="dist", n=1_000_000, m=50_000_000, T=2500, delta=100, order=5,
                 label="MultiOrderModel orders 1-5 lift on 50M-edge temporal stream, cross-partition edge all-to-all"),
}


def static_config(name: str, world: int) -> dict:
    """What both arms (this one and --impl reference) print as `config`: the workload only, nothing measured."""
    cfg = WORKLOADS[name]
    out = {"workload": cfg["label"], "name": name, "nodes": cfg["n"], "max_order": cfg["order"], "seed": 0}
    if cfg["kind"] == "train":
        out.update(walks=cfg["walks"], walk_length=f"randint({cfg['min_len']}, {cfg['max_len'] + 1})",
                   dbgnn={"hidden_dims": [cfg["hidden"]] * 3, "classes": cfg["classes"], "features": "dense randn fp32", "optimizer": "Adam"},
                   parallelism=f"dp{world}: walks sharded by walk id, one flat gradient all-reduce per step")
        return out
    out.update(edges=cfg["m"], timestamps=cfg["T"], delta=cfg["delta"])
    if cfg["kind"] == "dist":
        out["parallelism"] = (f"{world} rank(s): time-range partition of ONE stream, ghost zone (K-1)*delta, per order 16-byte "
                              "records to the owner of their row + 4-byte ids back")
    else:
        if cfg["dbgnn"]:
            out["dbgnn"] = {"hidden_dims": [cfg["hidden"]] * 3, "classes": cfg["classes"], "features": "dense randn fp32"}
        out["parallelism"] = f"{world} independent graphs (one per GPU)"
    return out


def ghost_overhead(cfg, world):
    """Estimated extra work of a rank that has a ghost zone (all but the last): at order j it also lifts the paths that
    start up to (K - j) delta after its own range; orders weighted by their line-graph sizes E_j ~ c^(j-2),
    c = (m / n)(delta / T).  The last rank gets 1 + g times the others' share of the stream."""
    K, c = cfg["order"], cfg["m"] / cfg["n"] * cfg["delta"] / cfg["T"]
    weights = [c ** (j - 2) for j in range(2, K + 1)]
    extra = sum(w * (K - j) for w, j in zip(weights, range(2, K + 1))) / sum(weights)
    return extra * cfg["delta"] * world / cfg["T"]


def make_stream(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, cfg["n"], (2, cfg["m"]), generator=g)
    t = torch.sort(torch.randint(0, cfg["T"], (cfg["m"],), generator=g)).values
    return ei, t


def make_walks(cfg, seed):
    """(flat node ids, lengths): random walks over N nodes; the first walks enumerate all nodes so that every shard of
    a rank covers the node set (lift_order.py:133-143 sizes layer 1 by the distinct nodes present)."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(cfg["min_len"], cfg["max_len"] + 1, (cfg["walks"],), generator=g)
    flat = torch.randint(0, cfg["n"], (int(lengths.sum()),), generator=g)
    return flat, lengths


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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


def stage_roofline(alg_bytes, ms, peaks, **more):
    achieved = alg_bytes / (ms / 1e3) / 1e9
    out = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
           "algorithmic_bytes": alg_bytes, "peak_source": peaks["source"]}
    out.update(more)
    return out


def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def max_over_ranks(values, dev, world):
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return t.tolist()


# ------------------------------------------------------------------------------------------------ algorithmic bytes
def alg_bytes_lift(n, m, line_sizes, layers):
    """SURVEY.md 8d.  a1: 24 m + 16 E_2; a2 at order k >= 3: 16 E_{k-1} + 16 E_k; a4 at order k:
    8(k+1) E_{k-1} (k-gram rows + inverse) + 20 E_k (edges + weights) + 8k n_k + 20 E^_k."""
    K = max(layers)
    n1, eh1 = layers[1]
    total = 8 * 2 * n + 20 * m + 8 * n1 + 20 * eh1                      # layer 1 (rows = arange(N))
    if K >= 2:
        total += 24 * m + 16 * line_sizes[2]                             # a1
    prev_line = m
    for k in range(2, K + 1):
        nk, ehk = layers[k]
        if k >= 3:
            total += 16 * line_sizes[k - 1] + 16 * line_sizes[k]        # a2
        total += 8 * (k + 1) * prev_line + 20 * line_sizes[k] + 8 * k * nk + 20 * ehk
        prev_line = line_sizes[k]
    return total


def alg_bytes_dbgnn(n, e, n2, e2, H, classes):
    """SURVEY.md 8d per GCN layer: 20 e_sl + 4H e_sl + 12 H n; bipartite: 16 nK + 4H nK + 8H (nK + N) + 4H N."""
    def gcn(nn, ee):
        esl = ee + nn
        return 20 * esl + 4 * H * esl + 12 * H * nn
    return 2 * gcn(n, e) + 2 * gcn(n2, e2) + 16 * n2 + 4 * H * n2 + 8 * H * (n2 + n) + 4 * H * n + 4 * n * (H + classes)


def fused_min_bytes_dbgnn(n, e, n2, e2, H, classes):
    """Compulsory traffic of the FUSED layers (gather + transform + activation in one kernel, every feature row read
    from HBM once thanks to the 126 MB L2, CSC structure 8 B per edge slot + 8 B per node): per GCN layer
    8 e + 8 n + 4H n (read) + 4H n (write); bipartite 4 nK + 4H (nK + N) + 4H N; classifier 4 N (H + classes)."""
    def gcn(nn, ee):
        return 8 * ee + 8 * nn + 8 * H * nn
    return 2 * gcn(n, e) + 2 * gcn(n2, e2) + 4 * n2 + 4 * H * (n2 + n) + 4 * H * n + 4 * n * (H + classes)


# ------------------------------------------------------------------------------------------------ in-step kernel timing
KERNEL_KINDS = {0: "onesweep_pass_kernel (radix digit pass)", 1: "chain_tile_kernel (expand + rank one level of the layer chain)",
                2: "merge_tile_kernel (owner-side merge of the senders' sorted runs)", 3: "gcn_tc_kernel (fused tcgen05 GCN layer: staged producer / consumer kernel on sparse layers, single-role on dense ones)"}


def profiled_kernels(lib, fn):
    """Run `fn` once with the library's CUDA-event hook switched on; returns [(ms, items, bytes per item, kind)] of every
    hot-kernel launch inside it (the launches of a REAL step, on the launching stream)."""
    cap = 512
    ms, items, bpi, kind = (ctypes.c_float * cap)(), (ctypes.c_int64 * cap)(), (ctypes.c_int * cap)(), (ctypes.c_int * cap)()
    count = ctypes.c_int(0)
    lib.ppg_profile_begin()
    try:
        fn()
    finally:
        lib.ppg_profile_end(ms, items, bpi, kind, cap, ctypes.byref(count))
    return [(ms[i], items[i], bpi[i], kind[i]) for i in range(count.value)]


def kernel_roofline(launches, peaks, runs, step_ms=None, gcn_bytes=None):
    """Dominant kernel = the kind with the largest total device time inside the profiled steps; reported for its largest
    launch shape (average over the launches of that shape).  ``runs``: number of profiled steps; ``gcn_bytes``: nodes of a
    layer -> compulsory bytes of the fused layer (the library does not know the edge count)."""
    if not launches:
        return None
    per_kind = {}
    for ms, items, bpi, kind in launches:
        nbytes = items * bpi if kind != 3 else (gcn_bytes or {}).get(items, 0)
        d = per_kind.setdefault(kind, {"launches": 0, "ms_total": 0.0, "bytes": 0})
        d["launches"] += 1
        d["ms_total"] += ms
        d["bytes"] += nbytes
    top_kind = max(per_kind, key=lambda k: per_kind[k]["ms_total"])
    mine = [l for l in launches if l[3] == top_kind]
    top_items = max(l[1] for l in mine)
    top = [l for l in mine if l[1] == top_items]
    ms = sum(l[0] for l in top) / len(top)
    per_launch = top[0][1] * max(l[2] for l in top) if top_kind != 3 else (gcn_bytes or {}).get(top_items, 0)
    achieved = per_launch / (ms / 1e3) / 1e9
    out = {"bound": "hbm", "kernel": KERNEL_KINDS[top_kind], "items": top_items, "achieved": achieved, "peak": peaks["hbm_gbs"],
           "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(top_kind, top_items),
           "algorithmic_bytes_per_launch": per_launch, "launch_ms": ms, "launches_averaged": len(top),
           "by_kernel": {KERNEL_KINDS[k].split(" ")[0]: {"launches_per_step": d["launches"] / runs, "ms_per_step": d["ms_total"] / runs,
                                                         "achieved": d["bytes"] / (d["ms_total"] / 1e3) / 1e9 if d["ms_total"] else None}
                         for k, d in sorted(per_kind.items())},
           "how": "CUDA events around every launch of the hot kernels INSIDE real steps (library profile hook, launching stream); the "
                  "dominant kernel is the one with the largest total time; bytes = what the launch must read and write once",
           "peak_source": peaks["source"]}
    if step_ms:
        out["share_of_step"] = per_kind[top_kind]["ms_total"] / runs / step_ms
    return out


def ncu_traffic(kind, items):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/kernel_traffic.json: "<kind>:<items>" -> bytes); None if that launch shape was not captured."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f).get(f"{kind}:{items}", {}).get("dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------------------------------------ cfg2 / cfg3
def run_lift(args, name, rank, world, local_rank, as_extra=False):
    import pathpyg_b200 as pp
    from pathpyg_b200 import _lib

    cfg = WORKLOADS[name]
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    ei_h, t_h = make_stream(cfg, seed=rank)
    ei_pin, t_pin = ei_h.pin_memory(), t_h.pin_memory()
    ei, t = ei_h.to(dev), t_h.to(dev)
    H, K, with_dbgnn = cfg["hidden"], cfg["order"], cfg["dbgnn"]
    steps = args.steps if not as_extra else max(3, min(args.steps, 10))
    warmup = args.warmup if not as_extra else 3
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if name == "cfg2" else None
    stream = torch.cuda.current_stream(dev)
    x_pin = torch.randn(cfg["n"], H, generator=torch.Generator().manual_seed(1000 + rank)).pin_memory() if with_dbgnn else None
    x = x_pin.to(dev) if with_dbgnn else None
    net = pp.nn.DBGNN(num_classes=cfg["classes"], num_features=(H, H), hidden_dims=[H, H, H]).to(dev).eval() if with_dbgnn else None
    state = {}

    def dbgnn_step(model, x_dev):
        model.layers[1].data.x = x_dev
        nK = model.layers[K].n
        if state.get("x_h") is None or state["x_h"].size(0) != nK:
            # higher-order features live on the device: their row count is an OUTPUT of the lift (the reference's
            # eye(n_K) is infeasible at this size)
            state["x_h"] = torch.randn(nK, H, generator=torch.Generator().manual_seed(7)).to(dev)
        data = model.to_dbgnn_data(max_order=K, x_h=state["x_h"])
        with torch.no_grad():
            return net(data)

    resident = pp.TemporalGraph.from_tensors(ei, t, cfg["n"])  # the input container, resident in HBM before the timed region

    def one_step(timed):
        if flush is not None:
            flush.fill_(1)  # evict L2 between steps
        e0, e1, e2_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0 = lib.ppg_launch_count()
        e0.record(stream)
        model = pp.MultiOrderModel.from_temporal_graph(resident, delta=cfg["delta"], max_order=K)
        e1.record(stream)
        c1 = lib.ppg_launch_count()
        out = dbgnn_step(model, x) if with_dbgnn else None
        e2_.record(stream)
        c2 = lib.ppg_launch_count()
        if timed is not None:
            timed.append((e0, e1, e2_, c1 - c0, c2 - c1))
        return model, out

    for _ in range(warmup):
        model, out = one_step(None)
    torch.cuda.synchronize(dev)
    layers = {k: (g.n, g.m) for k, g in model.layers.items()}
    # lifted edges per step = line-graph columns of every order = total weight of every layer (unit event weights)
    line_sizes = {k: int(round(float(model.layers[k].data.edge_weight.double().sum()))) for k in range(2, K + 1)}
    lifted = sum(line_sizes.values())
    nodes = cfg["n"] + layers[K][0]
    del model, out

    timed = []
    barrier(world)
    torch.cuda.synchronize(dev)
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        for _ in range(steps):
            one_step(timed)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - wall0
    barrier(world)
    lift_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in timed)
    dbgnn_ms = sum(b.elapsed_time(c) for _, b, c, _, _ in timed)

    # ---- e2e: the public calls with HOST (pinned) inputs and a HOST (pinned) result, copies inside the timed region:
    # lift + DBGNN as ONE number (24 B per event + the first-order features in, the activations out); for a lift-only
    # workload the finished max-order layer (edge index + weights) is what comes back.
    host_graph = pp.TemporalGraph.from_tensors(ei_pin, t_pin, cfg["n"])   # the caller's graph: pinned HOST tensors
    out_pin = torch.empty((cfg["n"], cfg["classes"]), dtype=torch.float32).pin_memory() if with_dbgnn else None
    stage_pin = torch.empty(64 << 20, dtype=torch.uint8).pin_memory() if not with_dbgnn else None
    d2h_bytes = [0]

    def copy_back(tensor):
        """device -> pinned host through a fixed staging buffer (the consumer reads it there)."""
        flat = tensor.contiguous().view(torch.uint8).reshape(-1)
        for a in range(0, flat.numel(), stage_pin.numel()):
            b = min(flat.numel(), a + stage_pin.numel())
            stage_pin[:b - a].copy_(flat[a:b], non_blocking=True)
        d2h_bytes[0] += flat.numel()

    def e2e_step():
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d2h_bytes[0] = 0
        a.record(stream)
        model_ = pp.MultiOrderModel.from_temporal_graph(host_graph, delta=cfg["delta"], max_order=K, device=dev)
        if with_dbgnn:
            x_dev = x_pin.to(dev, non_blocking=True)
            out_pin.copy_(dbgnn_step(model_, x_dev), non_blocking=True)
            d2h_bytes[0] = out_pin.numel() * 4
        else:
            top = model_.layers[K].data
            copy_back(top.edge_index.as_tensor())
            copy_back(top.edge_weight)
        b.record(stream)
        return a, b

    e2e_steps = max(3, min(steps, 20))
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize(dev)
    rec = [e2e_step() for _ in range(e2e_steps)]
    torch.cuda.synchronize(dev)
    e2e_ms = sum(a.elapsed_time(b) for a, b in rec) / e2e_steps
    h2d = ei_pin.numel() * 8 + t_pin.numel() * 8 + (x_pin.numel() * 4 if with_dbgnn else 0)

    # ---- dominant kernel: every radix digit pass launched inside three more real steps, timed by the library's hook
    launches = []
    for _ in range(3):
        launches += profiled_kernels(lib, lambda: one_step(None))
    lift_ms, dbgnn_ms, e2e_ms = max_over_ranks([lift_ms, dbgnn_ms, e2e_ms], dev, world)
    if rank != 0:
        return None
    peaks = load_peaks()
    lift_bytes = alg_bytes_lift(cfg["n"], cfg["m"], line_sizes, layers)
    line = {
        "metric": "k-order lift edges/s",
        "value": world * lifted * steps / (lift_ms / 1e3),
        "unit": "lifted edges/s",
        "n_gpus": world,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": (lift_ms + dbgnn_ms) / steps,
        "ms_per_step_lift": lift_ms / steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64 indices (int32/u64 keys inside) + fp32 weights/activations",
        "data": "synthetic",
        "config": static_config(name, world),
        "workload_stats": {"lifted_edges_per_step": lifted, "line_graph_columns": {str(k): v for k, v in line_sizes.items()},
                           "layers_nodes_edges": {str(k): v for k, v in layers.items()},
                           "l2": "flushed between steps (512 MB write)" if flush is not None else "inputs and intermediates larger than L2"},
        "lift_stage_roofline": stage_roofline(lift_bytes, lift_ms / steps, peaks),
        "roofline": kernel_roofline(launches, peaks, 3, step_ms=(lift_ms + dbgnn_ms) / steps,
                                    gcn_bytes={n_: 8 * e_ + 8 * n_ + 8 * H * n_ for n_, e_ in layers.values()} if with_dbgnn else None),
        "e2e": {"value": world * lifted / (e2e_ms / 1e3), "unit": "lifted edges/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h_bytes[0], "ms_per_step": e2e_ms,
                "what": ("from_temporal_graph(host graph, device=...) + DBGNN forward as ONE timed region: pinned host edge index, "
                         "time stamps and first-order features in, pinned host activations out; higher-order features are "
                         "device-resident (their row count is an output of the lift)") if with_dbgnn else
                        "from_temporal_graph(host graph, device=...): pinned host events in, the max-order layer (edge index + weights) "
                        "back to pinned host memory"},
        "gpu_launches": timed[-1][3] + timed[-1][4],
        "gpu_launches_lift": timed[-1][3],
        "wall_s_timed_region": wall,
        "clocks": clocks.summary(),
    }
    if with_dbgnn:
        db_bytes = alg_bytes_dbgnn(cfg["n"], layers[1][1], layers[K][0], layers[K][1], H, cfg["classes"])
        db_min = fused_min_bytes_dbgnn(cfg["n"], layers[1][1], layers[K][0], layers[K][1], H, cfg["classes"])
        line["ms_per_step_dbgnn"] = dbgnn_ms / steps
        line["gpu_launches_dbgnn"] = timed[-1][4]
        line["dbgnn"] = {"metric": "DBGNN forward nodes/s", "value": world * nodes * steps / (dbgnn_ms / 1e3), "unit": "nodes/s",
                         "nodes_per_step": nodes, "ms_per_step": dbgnn_ms / steps,
                         "roofline": stage_roofline(db_bytes, dbgnn_ms / steps, peaks, bytes_model="SURVEY 8d, unfused (every edge gather at full price)"),
                         "roofline_fused_min": stage_roofline(db_min, dbgnn_ms / steps, peaks,
                                                              bytes_model="compulsory traffic of the fused layers (each feature row once)")}
    return line


# ------------------------------------------------------------------------------------------------ cfg5
def single_gpu_reference_build(pp, parallel, ei, t, cfg, dev, timed_steps):
    """The whole stream on ONE device through MultiOrderModel.from_temporal_graph: digests of every layer, lifted
    edges, layer sizes and (optionally) the device time of `timed_steps` further builds."""
    K = cfg["order"]
    tg = pp.TemporalGraph.from_tensors(ei.to(dev), t.to(dev), cfg["n"])
    model = pp.MultiOrderModel.from_temporal_graph(tg, delta=cfg["delta"], max_order=K)
    digests = torch.stack([parallel.layer_digest(model.layers[k].data.edge_index, model.layers[k].data.edge_weight,
                                                 model.layers[k].data.node_sequence) for k in range(1, K + 1)])
    layers = {k: (g.n, g.m) for k, g in model.layers.items()}
    line_sizes = {k: int(round(float(model.layers[k].data.edge_weight.double().sum()))) for k in range(2, K + 1)}
    del model
    ms = []
    for _ in range(timed_steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model = pp.MultiOrderModel.from_temporal_graph(tg, delta=cfg["delta"], max_order=K)
        b.record()
        torch.cuda.synchronize(dev)
        ms.append(a.elapsed_time(b))
        del model
    del tg
    torch.cuda.empty_cache()
    return digests, layers, line_sizes, ms


def run_dist(args, name, rank, world, local_rank, as_extra=False):
    import pathpyg_b200 as pp
    from pathpyg_b200 import _lib, parallel

    cfg = WORKLOADS[name]
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    K = cfg["order"]
    dist = torch.distributed
    own_group = False
    if not dist.is_initialized():   # one GPU: the same code path on a one-rank group (records are routed to the rank itself)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        own_group = True
    steps = args.steps if not as_extra else max(3, min(args.steps, 5))
    warmup = args.warmup if not as_extra else 3
    ei_h, t_h = make_stream(cfg, seed=0)                      # every rank draws the SAME stream and keeps its time range
    lo, hi = parallel.partition_stream(cfg["m"], rank, world, last_share=1.0 + ghost_overhead(cfg, world))
    ei_pin, t_pin = ei_h[:, lo:hi].contiguous().pin_memory(), t_h[lo:hi].contiguous().pin_memory()
    ei_l, t_l = ei_pin.to(dev), t_pin.to(dev)
    stream = torch.cuda.current_stream(dev)

    # ---- the same stream on ONE GPU (rank 0): digests for the parity check, and its time for the strong-scaling ratio
    single = None
    if rank == 0:
        single = single_gpu_reference_build(pp, parallel, ei_h, t_h, cfg, dev, timed_steps=3)
    del ei_h, t_h
    want = single[0] if rank == 0 else torch.zeros((K, 2), dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(want, src=0)

    def step():
        return parallel.distributed_temporal_layers(ei_l, t_l, cfg["n"], cfg["delta"], K)

    parity_ok = True
    for i in range(warmup):
        layers = step()
        if i == 0:   # bit-exactness of the distributed build, checked inside the measured run
            got = torch.stack([layers[k].digest() for k in range(1, K + 1)])
            if world > 1:
                dist.all_reduce(got)
            parity_ok = bool(torch.equal(got, want))
        del layers
    torch.cuda.synchronize(dev)

    timed = []
    barrier(world)
    torch.cuda.synchronize(dev)
    launches = 0
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0 = lib.ppg_launch_count()
            a.record(stream)
            layers = step()
            b.record(stream)
            launches = lib.ppg_launch_count() - c0
            timed.append((a, b))
            del layers
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - wall0
    barrier(world)
    step_ms = sum(a.elapsed_time(b) for a, b in timed)

    # ---- e2e: pinned HOST slice in, the max-order layer (edge index + weights of the owned rows) back to pinned host memory
    stage_pin = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    d2h = [0]

    def copy_back(tensor):
        flat = tensor.contiguous().view(torch.uint8).reshape(-1)
        for a in range(0, flat.numel(), stage_pin.numel()):
            b = min(flat.numel(), a + stage_pin.numel())
            stage_pin[:b - a].copy_(flat[a:b], non_blocking=True)
        d2h[0] += flat.numel()

    def e2e_step():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d2h[0] = 0
        a.record(stream)
        ei_d, t_d = ei_pin.to(dev, non_blocking=True), t_pin.to(dev, non_blocking=True)
        out = parallel.distributed_temporal_layers(ei_d, t_d, cfg["n"], cfg["delta"], K)
        copy_back(out[K].edge_index)
        copy_back(out[K].edge_weight)
        b.record(stream)
        return a, b

    e2e_step()
    torch.cuda.synchronize(dev)
    barrier(world)
    rec = [e2e_step() for _ in range(3)]
    torch.cuda.synchronize(dev)
    e2e_ms = sum(a.elapsed_time(b) for a, b in rec) / len(rec)
    hot = profiled_kernels(lib, step)
    step_ms, e2e_ms = max_over_ranks([step_ms, e2e_ms], dev, world)
    h2d_all = torch.tensor([ei_pin.numel() * 8 + t_pin.numel() * 8, d2h[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_all)
    if own_group:
        dist.destroy_process_group()
    if rank != 0:
        return None
    peaks = load_peaks()
    _, layers_single, line_sizes, single_ms = single
    lifted = sum(line_sizes.values())
    one_gpu_ms = min(single_ms)
    ms = step_ms / steps
    exchange_ms = ms
    if world == 1:
        # one GPU: what a user runs there is MultiOrderModel.from_temporal_graph (no exchange); the exchange path on a
        # one-rank group (every record routed to the rank itself) is reported beside it
        ms, step_ms = one_gpu_ms, one_gpu_ms * steps
    lift_bytes = alg_bytes_lift(cfg["n"], cfg["m"], line_sizes, layers_single)
    return {
        "metric": "k-order lift edges/s",
        "value": lifted * steps / (step_ms / 1e3),
        "unit": "lifted edges/s",
        "n_gpus": world,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "int64 indices (u32 ids and record fields inside) + fp32 weights",
        "data": "synthetic",
        "config": static_config(name, world),
        "parity_ok": parity_ok,
        "parity_how": "64-bit order-sensitive digests of edge index, weights and node sequences of every layer, summed over the ranks "
                      "(all-reduce), equal the digests of MultiOrderModel.from_temporal_graph on the whole stream on one GPU (same run)",
        "strong_scaling": {"one_gpu_ms_same_stream": one_gpu_ms, "one_gpu_path": "MultiOrderModel.from_temporal_graph (no exchange), rank 0, same run",
                           "n_gpu_ms": exchange_ms, "n_gpu_path": "parallel.distributed_temporal_layers (ghost zone + per-order exchange)",
                           "speedup": one_gpu_ms / exchange_ms, "efficiency": one_gpu_ms / exchange_ms / world},
        "workload_stats": {"lifted_edges_per_step": lifted, "line_graph_columns": {str(k): v for k, v in line_sizes.items()},
                           "layers_nodes_edges": {str(k): v for k, v in layers_single.items()}, "l2": "inputs and intermediates larger than L2"},
        "collectives_per_step": {"records into the owners' symmetric buffers (NVLink peer stores; all_to_all_v without peer access)": K,
                                 "all_to_all_v (merged-edge indices back)": K, "all_gather (counts)": 2 * K,
                                 "ghost zone all_to_all_v + 2 all_gather": 1},
        "lift_stage_roofline": stage_roofline(lift_bytes, ms, peaks, note="whole-job algorithmic bytes of the single-device formulation / step time; "
                                              "aggregate peak = n_gpus x per-GPU peak", aggregate_frac=lift_bytes / (ms / 1e3) / 1e9 / (peaks["hbm_gbs"] * world)),
        "roofline": kernel_roofline(hot, peaks, 1, step_ms=exchange_ms),
        "e2e": {"value": lifted / (e2e_ms / 1e3), "unit": "lifted edges/s", "h2d_bytes_per_step": int(h2d_all[0]), "d2h_bytes_per_step": int(h2d_all[1]),
                "ms_per_step": e2e_ms, "what": "pinned host slices of the stream in (24 B per event), distributed_temporal_layers, the max-order "
                                               "layer's owned edges + weights back to pinned host memory (bytes summed over the ranks)"},
        "gpu_launches": launches,
        "wall_s_timed_region": wall,
        "clocks": clocks.summary(),
    }


# ------------------------------------------------------------------------------------------------ cfg4
def run_train(args, name, rank, world, local_rank, as_extra=False):
    import pathpyg_b200 as pp
    from pathpyg_b200 import _lib, parallel

    cfg = WORKLOADS[name]
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    H, n = cfg["hidden"], cfg["n"]
    steps = args.steps if not as_extra else max(3, min(args.steps, 5))
    warmup = args.warmup if not as_extra else 3
    flat, lengths = make_walks(cfg, seed=0)
    lo, hi = parallel.shard_walks(lengths, rank, world)
    starts = torch.cumsum(lengths, 0) - lengths
    a, b = int(starts[lo]), int(starts[hi - 1] + lengths[hi - 1])
    # every shard must contain every first-order node (lift_order.py:133-143): one covering walk per shard
    my_flat = torch.cat([flat[a:b], torch.arange(n)])
    my_len = torch.cat([lengths[lo:hi], torch.tensor([n])])
    del flat, lengths
    flat_pin, len_pin = my_flat.pin_memory(), my_len.pin_memory()
    flat_d, len_d = flat_pin.to(dev), len_pin.to(dev)
    w_d = torch.ones(len_d.numel(), device=dev)
    x = torch.randn(n, H, generator=torch.Generator().manual_seed(1)).to(dev)
    y = torch.randint(0, cfg["classes"], (n,), generator=torch.Generator().manual_seed(2)).to(dev)
    torch.manual_seed(0)
    net = pp.nn.DBGNN(num_classes=cfg["classes"], num_features=(H, H), hidden_dims=[H, H, H]).to(dev)
    parallel.broadcast_parameters(net)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    stream = torch.cuda.current_stream(dev)
    state = {}

    def step(flat_in, len_in, timed=None):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        paths = pp.PathData(device=dev)
        paths.append_index_walks(flat_in, len_in, w_d)
        model = pp.MultiOrderModel.from_path_data(paths, max_order=2)
        e1.record(stream)
        model.layers[1].data.x = x
        n2 = model.layers[2].n
        if state.get("x_h") is None or state["x_h"].size(0) != n2:
            state["x_h"] = torch.randn(n2, H, generator=torch.Generator(device=dev).manual_seed(7 + rank), device=dev)
        data = model.to_dbgnn_data(max_order=2, x_h=state["x_h"])
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(net(data), y)
        loss.backward()
        parallel.allreduce_gradients(net)
        opt.step()
        e2.record(stream)
        if timed is not None:
            timed.append((e0, e1, e2))
        return model, loss

    for _ in range(warmup):
        model, loss = step(flat_d, len_d)
    torch.cuda.synchronize(dev)
    layers = {k: (g.n, g.m) for k, g in model.layers.items()}
    lifted_local = int(round(float(model.layers[2].data.edge_weight.double().sum())))     # second-order walk edges of the shard
    nodes_local = n + layers[2][0]
    first_loss = float(loss.detach())
    del model
    timed = []
    barrier(world)
    torch.cuda.synchronize(dev)
    with ClockSampler(local_rank) as clocks:
        c0 = lib.ppg_launch_count()
        for _ in range(steps):
            _, loss = step(flat_d, len_d, timed)
        torch.cuda.synchronize(dev)
        launches = (lib.ppg_launch_count() - c0) // steps
    barrier(world)
    lift_ms = sum(a_.elapsed_time(b_) for a_, b_, _ in timed)
    train_ms = sum(b_.elapsed_time(c_) for _, b_, c_ in timed)
    # replicas must still agree after the timed steps (same gradients everywhere)
    flat_w = torch.cat([q.detach().reshape(-1) for q in net.parameters()])
    ref_w = flat_w.clone()
    if world > 1:
        torch.distributed.broadcast(ref_w, src=0)
    in_sync = torch.tensor([int(torch.equal(ref_w, flat_w))], device=dev)
    if world > 1:
        torch.distributed.all_reduce(in_sync, op=torch.distributed.ReduceOp.MIN)

    loss_pin = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        _, l = step(flat_pin.to(dev, non_blocking=True), len_pin.to(dev, non_blocking=True))
        loss_pin.copy_(l.detach(), non_blocking=True)
        b_.record(stream)
        return a_, b_

    e2e_step()
    torch.cuda.synchronize(dev)
    rec = [e2e_step() for _ in range(3)]
    torch.cuda.synchronize(dev)
    e2e_ms = sum(a_.elapsed_time(b_) for a_, b_ in rec) / len(rec)
    lift_ms, train_ms, e2e_ms = max_over_ranks([lift_ms, train_ms, e2e_ms], dev, world)
    totals = torch.tensor([lifted_local, nodes_local, flat_pin.numel() * 8 + len_pin.numel() * 8], dtype=torch.int64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(totals)
    if rank != 0:
        return None
    lifted, nodes, h2d = (int(v) for v in totals.tolist())
    total_ms = lift_ms + train_ms
    return {
        "metric": "k-order lift edges/s",
        "value": lifted * steps / (total_ms / 1e3),
        "unit": "lifted edges/s",
        "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "ms_per_step_lift": lift_ms / steps, "ms_per_step_train": train_ms / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int64 indices + fp32 weights/activations/gradients", "data": "synthetic",
        "config": static_config(name, world),
        "what": "one training step = from_path_data of the rank's walk shard (order 2) + DBGNN forward + cross-entropy + backward + "
                "flat NCCL all-reduce of the gradients + Adam; value = second-order walk edges of all shards per second of the WHOLE step",
        "dbgnn": {"metric": "DBGNN training nodes/s", "value": nodes * steps / (train_ms / 1e3), "unit": "nodes/s", "ms_per_step": train_ms / steps},
        "workload_stats": {"rank0_layers_nodes_edges": {str(k): v for k, v in layers.items()}, "lifted_edges_per_step": lifted,
                           "loss_first_last": [first_loss, float(loss.detach())], "replicas_in_sync": bool(int(in_sync))},
        "e2e": {"value": lifted / (e2e_ms / 1e3), "unit": "lifted edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": e2e_ms, "what": "pinned host walks in (8 B per position), the loss back"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }


RUNNERS = {"lift": run_lift, "dist": run_dist, "train": run_train}


# ------------------------------------------------------------------------------------------------ CPU arm
_stream_cache = {}


def cpu_lift_sample(cfg, budget_s, threads, stamps=None):
    """The reference's CPU implementation of the path (oracle port, reference operation order) on a bounded sample:
    ``oracle.lift.lift_order_temporal`` over the first S distinct source time stamps against the FULL stream (the
    per-time-stamp work is exactly the full run's), then ``oracle.lift.aggregate_edge_index`` of the order-2 layer for
    the event pairs produced.  Streams beyond 2M events aggregate over the events the pairs touch only (the reference's
    ``torch.unique(dim=0)`` over all [m, 2] rows alone takes minutes there: 0.24 M rows/s)."""
    from oracle import lift

    torch.set_num_threads(threads)
    key = (cfg["n"], cfg["m"], cfg["T"])
    if key not in _stream_cache:
        _stream_cache.clear()
        _stream_cache[key] = make_stream(cfg, seed=0)
    ei, t = _stream_cache[key]
    stats = {}
    t0 = time.perf_counter()
    ho = lift.lift_order_temporal(ei, t, cfg["delta"], max_source_stamps=stamps, budget_s=budget_s, stats=stats)
    t_lift = time.perf_counter() - t0
    t1 = time.perf_counter()
    if cfg["m"] > 2_000_000:
        used, compact = torch.unique(ho, return_inverse=True)
        lift.aggregate_edge_index(compact, ei[:, used].t().contiguous(), None)
    else:
        lift.aggregate_edge_index(ho, ei.t().contiguous(), None)
    t_agg = time.perf_counter() - t1
    return {"edges": int(ho.size(1)), "seconds": t_lift + t_agg, "stamps": stats["stamps"], "of_stamps": stats["of_stamps"],
            "t_lift": t_lift, "t_aggregate": t_agg}


def cpu_train_sample(cfg, threads, walks=20_000):
    """cfg4 on the CPU: the oracle's from_path_data (order 2) on the first `walks` walks plus one walk that covers the
    node set (as every GPU shard carries)."""
    from oracle import mom

    torch.set_num_threads(threads)
    if "walks" not in _stream_cache:
        _stream_cache.clear()
        _stream_cache["walks"] = make_walks(cfg, seed=0)
    flat, lengths = _stream_cache["walks"]
    lengths = torch.cat([lengths[:walks], torch.tensor([cfg["n"]])])
    flat = torch.cat([flat[:int(lengths[:-1].sum())], torch.arange(cfg["n"])])
    t0 = time.perf_counter()
    pos = torch.arange(int(lengths.sum()))
    chain = torch.stack([pos[:-1], pos[1:]])
    keep = torch.ones(chain.size(1), dtype=torch.bool)
    keep[torch.cumsum(lengths, 0)[:-1] - 1] = False                  # path_data.py:144-151
    store = mom.Walks(chain[:, keep], flat.unsqueeze(1), torch.ones(lengths.numel()), lengths - 1, lengths)
    layers = mom.from_path_data(store, max_order=2)
    dt = time.perf_counter() - t0
    return {"edges": int(round(float(layers[2].edge_weight.sum()))), "seconds": dt, "walks": walks}


def cpu_baseline(name, budget_s):
    cfg = WORKLOADS[name]
    threads = os.cpu_count() or 1
    s = cpu_lift_sample(cfg, budget_s, threads)
    return {"value": s["edges"] / s["seconds"], "unit": "lifted edges/s", "cores": threads, "kind": "port",
            "extrapolated": s["stamps"] < s["of_stamps"],
            "sample": f"oracle.lift.lift_order_temporal over the first {s['stamps']} of {s['of_stamps']} source time stamps against the "
                      f"full {cfg['m']}-edge stream ({s['t_lift']:.1f} s) + oracle.lift.aggregate_edge_index of the order-2 layer on the "
                      f"{s['edges']} event edges produced ({s['t_aggregate']:.1f} s); the rate is per produced edge, the full run is "
                      f"{s['of_stamps']}/{s['stamps']} times longer"}


def cpu_dbgnn_forward(cfg, threads):
    """Oracle DBGNN forward (plain torch CPU restatement of PyG GCNConv + the reference's bipartite operator) on the
    workload's full layers; the event graph comes from the oracle's closed-form lift, because the reference-order loop
    does not finish at this size (that loop is what the lift line samples)."""
    from oracle import dbgnn as odb
    from oracle import lift

    torch.set_num_threads(threads)
    ei, t = make_stream(cfg, seed=0)
    H = cfg["hidden"]
    eg = torch.from_numpy(lift.lift_order_temporal_closed_form(ei.numpy(), t.numpy(), cfg["delta"]))
    l1 = lift.aggregate_edge_index(ei, torch.arange(cfg["n"]).unsqueeze(1), None)
    l2 = lift.aggregate_edge_index(eg, ei.t().contiguous(), None)
    g = torch.Generator().manual_seed(3)
    data = {"x": torch.randn(cfg["n"], H, generator=g), "x_h": torch.randn(l2.num_nodes, H, generator=g), "num_nodes": cfg["n"],
            "edge_index": l1.edge_index, "edge_weights": l1.edge_weight.float(), "edge_index_higher_order": l2.edge_index,
            "edge_weights_higher_order": l2.edge_weight.float(),
            "bipartite_edge_index": torch.stack([torch.arange(l2.num_nodes), l2.node_sequence[:, 1]])}
    params = odb.init_params(cfg["classes"], (H, H), [H, H, H], seed=1)
    odb.dbgnn_forward(params, data)
    t0 = time.perf_counter()
    odb.dbgnn_forward(params, data)
    dt = time.perf_counter() - t0
    nodes = cfg["n"] + l2.num_nodes
    return {"metric": "DBGNN forward nodes/s", "value": nodes / dt, "unit": "nodes/s", "seconds": dt, "nodes": nodes, "cores": threads,
            "kind": "port", "how": "oracle.dbgnn.dbgnn_forward (torch CPU) on the full cfg2 layers (built with the oracle's closed-form lift)"}


def run_reference(args, name, rank, world):
    if rank != 0:
        return
    cfg = WORKLOADS[name]
    threads = os.cpu_count() or 1
    if cfg["kind"] == "train":
        for _ in range(args.warmup):
            cpu_train_sample(cfg, threads, walks=2_000)
        edges, seconds, last = 0, 0.0, None
        for _ in range(args.steps):
            last = cpu_train_sample(cfg, threads)
            edges += last["edges"]
            seconds += last["seconds"]
        sample = f"per step: oracle.mom.from_path_data (order 2, torch CPU, {threads} threads) on the first {last['walks']} of {cfg['walks']} walks"
        extrapolated, dbgnn = True, None
    else:
        full = args.steps <= 1   # one honest full run when asked for a single step
        per_step_budget = None if full else max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
        for _ in range(0 if full else args.warmup):
            cpu_lift_sample(cfg, per_step_budget, threads)
        edges, seconds, last = 0, 0.0, None
        for _ in range(args.steps):
            last = cpu_lift_sample(cfg, per_step_budget, threads)
            edges += last["edges"]
            seconds += last["seconds"]
        extrapolated = last["stamps"] < last["of_stamps"]
        sample = (f"per step: oracle.lift.lift_order_temporal (reference operation order, torch CPU, {threads} threads) over the first "
                  f"{last['stamps']} of {last['of_stamps']} source time stamps against the full stream + oracle.lift.aggregate_edge_index of the order-2 layer")
        dbgnn = cpu_dbgnn_forward(cfg, threads) if cfg.get("dbgnn") else None
    value = edges / seconds
    line = {
        "impl": "reference", "metric": "k-order lift edges/s", "value": value, "unit": "lifted edges/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if cfg["kind"] != "lift" else "weak", "vs_baseline": None, "dtype": "int64 + fp32 (torch CPU)", "data": "synthetic",
        "config": static_config(name, world),
        "extrapolated": extrapolated,
        "cpu_baseline": {"value": value, "unit": "lifted edges/s", "cores": threads, "kind": "port", "sample": sample, "extrapolated": extrapolated},
        "e2e": {"value": value, "unit": "lifted edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if dbgnn is not None:
        line["dbgnn"] = dbgnn
    print(json.dumps(line))


def default_workload(world: int) -> str:
    return "cfg2" if world == 1 else "cfg5"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads reported under `extra`")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps on the GPU arm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload or default_workload(world)
    if args.impl == "reference":
        run_reference(args, name, rank, world)
        return
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        line = RUNNERS[WORKLOADS[name]["kind"]](args, name, rank, world, local_rank)
        extras = {}
        if not args.no_extras and args.workload is None:
            # the other workloads that make sense at this world size, shortened
            for other in (["cfg3", "cfg5", "cfg4"] if world == 1 else ["cfg4"]):
                try:
                    res = RUNNERS[WORKLOADS[other]["kind"]](args, other, rank, world, local_rank, as_extra=True)
                except Exception as exc:  # an extra must never take the headline line down
                    res = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                    torch.cuda.empty_cache()
                if rank == 0:
                    extras[other] = res
        if rank == 0:
            if extras:
                line["extra"] = extras
            if world == 1 and WORKLOADS[name]["kind"] == "lift":
                line["cpu_baseline"] = cpu_baseline(name, budget_s=args.cpu_budget)
            print(json.dumps(line))
    finally:
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


def _only_json_on_stdout(fn):
    """The contract is ONE JSON line on stdout: libraries that print there (NCCL announces its version on the first
    communicator) are sent to stderr for the duration of the run; `print` inside this module goes to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    try:
        fn()
    finally:
        sys.stdout.flush()


if __name__ == "__main__":
    _only_json_on_stdout(main)
