"""Summarise `ncu --set full` captures of the chain kernels into markdown + profiles/kernel_traffic.json (development aid).

    python scripts/ncu_chain_summary.py <tag> <out.md> rep1.ncu-rep [rep2.ncu-rep ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import METRICS, raw  # noqa: E402

EXTRA = [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
         ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-sb"),
         ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
         ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short-sb")]


def main(tag, out, reps):
    import csv
    import subprocess
    rows = []
    for rep in reps:
        base = raw(rep)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        table = list(csv.reader(txt.splitlines()))
        hdr = table[0]
        for d, r in zip(base, table[2:]):
            for m, _ in EXTRA:
                if m in hdr:
                    d[m] = float(r[hdr.index(m)].replace(",", ""))
            d["rep"] = os.path.basename(rep)
            rows.append(d)
    cols = METRICS + EXTRA
    with open(out, "w") as f:
        f.write(f"# ncu --set full: chain kernels ({tag})\n\nCaptured with `--clock-control none --import-source on` under gpurun on one B200; per-launch "
                "values (cold cache, serialised by ncu).  DRAM % is of ncu's own peak.\n\n| capture | kernel | " +
                " | ".join(lbl for _, lbl in cols) + " |\n|---|---|" + "---:|" * len(cols) + "\n")
        for d in rows:
            name = d["name"].replace("void ", "").replace("ppg::", "")[:48]
            f.write(f"| {d['rep']} | `{name}` | " + " | ".join(f"{d.get(m, float('nan')):.1f}" for m, _ in cols) + " |\n")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_traffic.json")
    traffic = json.load(open(path)) if os.path.isfile(path) else {}
    for d in rows:
        if "chain_tile_kernel" in d["name"] and "launch__grid_size" in d:
            # items are not in the report; key by grid size, resolved by the caller (see the markdown)
            traffic.setdefault("_chain_tile_by_grid", {})[str(int(d["launch__grid_size"]))] = {
                "dram_bytes_per_launch": int((d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) * 1e6), "source": f"{tag}: {d['rep']}"}
    json.dump(traffic, open(path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
