"""Summarise `ncu --set full` reports into markdown + the traffic JSON bench.py reads (development aid).

    python scripts/ncu_summary.py <tag>      # reads gpurun_out/<tag>_{sort1p8m,sort64m,gcn}.ncu-rep
"""
import csv
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB read"),
    ("dram__bytes_write.sum", "MB written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"name": r[hdr.index("Kernel Name")]}
        for m, _ in METRICS:
            if m in hdr:
                v, u = float(r[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
                if m.startswith("dram__bytes"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
                if m == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                d[m] = v
        out.append(d)
    return out


def table(f, title, rows):
    f.write(f"\n## {title}\n\n| kernel | " + " | ".join(lbl for _, lbl in METRICS) + " |\n|---|" + "---:|" * len(METRICS) + "\n")
    for d in rows:
        name = d["name"].replace("void ", "").replace("ppg::", "")[:60]
        f.write(f"| `{name}` | " + " | ".join(f"{d.get(m, float('nan')):.1f}" for m, _ in METRICS) + " |\n")


def main(tag):
    traffic = {}
    with open(f"profiles/{tag}_ncu_full_summary.md", "w") as f:
        f.write(f"# ncu --set full summaries ({tag})\n\nCaptured with `--clock-control none --import-source on` under gpurun on one B200; "
                "per-launch values (cold cache, serialised by ncu). DRAM % is of ncu's own peak; the roofline fractions in the bench line use "
                "the measured copy bandwidth of MEASURED_PEAKS.json instead.\n")
        for key, name, title in (("pairs_1.8M", "sort1p8m", "onesweep digit pass, 1.8M (u64 key, u32 payload) pairs = the order-2 coalesce of cfg2"),
                                 ("pairs_64M", "sort64m", "onesweep digit pass, 64M pairs"),
                                 (None, "gcn", "DBGNN layers at cfg2 (first two launches: first-order graph n=100k; next two: order-2 graph n=1M; last: bipartite)")):
            try:
                rows = raw(f"gpurun_out/{tag}_{name}.ncu-rep")
            except Exception as e:  # noqa: BLE001
                f.write(f"\n({name}: {e})\n")
                continue
            table(f, title, rows)
            if key:
                per = [d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in rows]
                traffic[key] = {"dram_bytes_per_launch": int(1e6 * sum(per) / len(per)), "launches": len(per),
                                "source": f"profiles/{tag}_ncu_full_summary.md"}
    with open("profiles/r01_sort_traffic.json", "w") as f:
        json.dump(traffic, f, indent=1)
    print(open(f"profiles/{tag}_ncu_full_summary.md").read())


if __name__ == "__main__":
    main(sys.argv[1])
