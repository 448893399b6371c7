"""Source lines of a kernel ranked by executed warp instructions, from an ncu report (development aid).

    python scripts/ncu_inst_lines.py report.ncu-rep <kernel-id> [top]
"""
import csv
import subprocess
import sys


def main(rep, kid, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                         capture_output=True, text=True).stdout
    cur, hdr, out = None, None, []
    for r in csv.reader(txt.splitlines()):
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] != "":
            try:
                out.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), cur, r[0], r[1].strip()[:110]))
            except ValueError:
                pass
    tot, stot = sum(o[0] for o in out), sum(o[1] for o in out)
    print("warp instructions executed:", tot, " stall samples:", stot)
    for o in sorted(out, reverse=True)[:top]:
        print(f"{100 * o[0] / tot:5.1f}% inst {100 * o[1] / max(stot, 1):5.1f}% samples  {o[2]}:{o[3]}  {o[4]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
