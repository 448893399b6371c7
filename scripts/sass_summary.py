"""SASS mnemonic counts per kernel of the shipped library -> profiles/<tag>_sass_summary.md (development aid).

    python scripts/sass_summary.py r02
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pathpyg_b200", "_C", "libpathpyg_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "LDGDEPBAR", "SYNCS", "NANOSLEEP", "ATOMS", "REDUX", "MATCH"]


def main(tag):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    rows, name, counts = [], None, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, counts))
            name, counts = m.group(1), dict.fromkeys(COLS, 0)
            continue
        if name:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1).split(".")[0]
                if op in counts:
                    counts[op] += 1
    if name:
        rows.append((name, counts))
    demangled = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    keep = [(d, c) for d, (_, c) in zip(demangled, rows) if any(c[k] for k in ("UTCHMMA", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "MATCH"))]
    out = [f"# SASS mnemonic counts of the shipped library ({tag})", "",
           "`cuobjdump -sass pathpyg_b200/_C/libpathpyg_b200.so`, counted per kernel (sm_100a; kernels without any of the listed",
           "instructions are left out).  `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, `LDGSTS` = cp.async,",
           "`LDGDEPBAR` = cp.async.commit_group, `SYNCS` = mbarrier operations, `NANOSLEEP` = parked mbarrier waits, `MATCH` = match.any.",
           "The staged GCN layer moves its rows with `LDGSTS.E.BYPASS.128` (16 bytes per lane, one row per 16 lanes): the rows are gathered",
           "by index, which a bulk / tensor copy cannot do, and the two places where a 1-D bulk copy was tried (the index tile of the ring",
           "variant, the nodes' own rows as a chunk of the stages) were measured and dropped (DESIGN.md section 4.2) -- there is no",
           "`UBLKCP` / `UTMALDG` in the library.", "",
           "| kernel | " + " | ".join(COLS) + " |", "|---|" + "---:|" * len(COLS)]
    for d, c in keep:
        out.append(f"| `{d[:110]}` | " + " | ".join(str(c[k]) for k in COLS) + " |")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, len(keep), "kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
