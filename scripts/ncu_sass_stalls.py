"""Stall reasons and hottest SASS instructions of the first kernel in an ncu report (development aid).

    python scripts/ncu_sass_stalls.py report.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[i_s]) for r in data)
    tot_i = sum(int(r[i_e]) for r in data)
    print(rows[0][1][:100])
    print("samples", tot, "warp instructions", tot_i, "SASS lines", len(data))
    agg = {s: sum(int(r[hdr.index(s)]) for r in data) for s in stalls}
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"  {k:28s} {100 * v / tot:5.1f}%")
    hot = sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:top]
    for i in sorted(hot):
        r = data[i]
        st = sorted(((s, int(r[hdr.index(s)])) for s in stalls), key=lambda x: -x[1])[:2]
        print(f"{i:5d} {100 * int(r[i_s]) / tot:5.1f}% inst {100 * int(r[i_e]) / tot_i:4.1f}%  {r[1].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
