"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel summary (markdown)."""
import collections
import csv
import re
import sys


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)[:100]
        v = float(row["Metric Value"])
        unit = row["Metric Unit"]
        us = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += us
        n += 1
    total = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary: `{path}`\n\n{n} launches, {total:.0f} us total device time "
                "(cold-cache, serialised by ncu: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / total:.1f}% |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
