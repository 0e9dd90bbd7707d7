"""Join an `ncu --page source --csv` SASS listing with `nvdisasm -gi` line info and aggregate executed
instructions / stall samples by source line (developer tool; needs the cubin the profile was taken with).

usage: python tools/ncu_by_line.py <src.csv> <cubin> <mangled kernel name> [inner|outer]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def line_map(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(out) if l.startswith(".text." + kernel + ":"))
    m, cur_inner, cur_outer, pending = {}, None, None, []
    for l in out[start + 1:]:
        if l.startswith("//---") and ".text." in l:
            break
        g = re.match(r'\s*//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', l)
        if g:
            pending.append((g.group(1).split("/")[-1], int(g.group(2))))
            continue
        a = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", l)
        if a:
            if pending:
                cur_inner, cur_outer = pending[0], pending[-1]
                pending = []
            m[int(a.group(1), 16)] = (cur_inner, cur_outer, a.group(2).strip())
    return m


def main():
    src, cubin, kernel = sys.argv[1:4]
    mode = sys.argv[4] if len(sys.argv) > 4 else "outer"
    m = line_map(cubin, kernel)
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: j for j, h in enumerate(hdr)}
    data = []
    for r in rows[hi + 1:]:
        if r and r[0] in ("Kernel Name", "Address"):
            break                      # only the first kernel of the listing
        if len(r) == len(hdr):
            data.append(r)
    base = int(data[0][0], 16)
    agg = defaultdict(lambda: defaultdict(float))
    tot = defaultdict(float)
    cols = ["Instructions Executed", "# Samples", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_selected", "stall_not_selected",
            "stall_barrier", "stall_branch_resolving", "stall_math", "stall_mio", "stall_lg", "L1 Wavefronts Shared"]
    for r in data:
        off = int(r[0], 16) - base
        inner, outer, _ = m.get(off, (None, None, ""))
        key = inner if mode == "inner" else outer
        for c in cols:
            v = float(r[ci[c]] or 0)
            agg[key][c] += v
            tot[c] += v
    print("%-28s" % "line" + "".join("%12s" % c.replace("stall_", "")[:11] for c in cols))
    for key in sorted(agg, key=lambda k: (k is None, k)):
        a = agg[key]
        if a["Instructions Executed"] < 0.002 * tot["Instructions Executed"] and a["# Samples"] < 0.002 * tot["# Samples"]:
            continue
        print("%-28s" % ("%s:%d" % key if key else "?") + "".join("%12.0f" % a[c] for c in cols))
    print("%-28s" % "TOTAL" + "".join("%12.0f" % tot[c] for c in cols))


if __name__ == "__main__":
    main()
