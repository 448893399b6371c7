"""Top source lines by warp-stall samples from an ncu report (development aid).

    python scripts/ncu_hot_lines.py report.ncu-rep <kernel-id> [top]
"""
import csv
import subprocess
import sys


def main(rep, kid, top=30):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, out, name = None, None, [], None
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            name = r[1][:90]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            try:
                s = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            if s > 0:
                g = lambda k: r[hdr.index(k)]
                out.append((s, cur, r[0], r[1].strip()[:100], "short", g("stall_short_sb"), "long", g("stall_long_sb"),
                            "bar", g("stall_barrier"), "wait", g("stall_wait"), "mio", g("stall_mio"), "math", g("stall_math")))
    tot = sum(o[0] for o in out)
    print(name, "total samples", tot)
    for o in sorted(out, reverse=True)[:top]:
        print(f"{100 * o[0] / tot:5.1f}%", *o[1:])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
