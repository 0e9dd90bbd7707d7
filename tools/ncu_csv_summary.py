"""Summarise the CSV pages of an ncu --set full capture (made on the GPU box: `ncu -i rep --page raw --csv`, `--page source --csv`)
into the text committed under profiles/: per-launch table, pipe utilisation, stall reasons, DRAM / L2 traffic, and for the
longest launch the opcode mix and the per-source-line table (SASS joined with `nvdisasm -gi` line info of the shipped library).

usage: python tools/ncu_csv_summary.py <raw.csv> [<source.csv.gz> <libtmvb.so>] > profiles/<name>.txt
"""
import collections
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: j for j, h in enumerate(hdr)}

    def g(r, n):
        return r[ci[n]] if n in ci else ""

    cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"), ("launch__registers_per_thread", "regs"),
            ("launch__shared_mem_per_block_dynamic", "dsmem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("smsp__inst_executed.sum", "inst"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
            ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaH%"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smemWF%"),
            ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"), ("lts__t_bytes.sum", "l2B"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sectors_op_red.sum", "l2red")]
    cols = [(n, s) for n, s in cols if n in ci]
    print("launches: %d   (units: %s)" % (len(data), ", ".join("%s=%s" % (s, units[ci[n]]) for n, s in cols if units[ci[n]])))
    print("%-4s %-62s" % ("#", "kernel") + "".join("%11s" % s for _, s in cols))
    for k, r in enumerate(data):
        name = re.sub(r"\(tmvb::|\(int\)|\(bool\)|tmvb::|void ", "", g(r, "Kernel Name"))[:62]
        print("%-4d %-62s" % (k, name) + "".join("%11s" % g(r, n)[:10] for n, _ in cols))
    st = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    if st:
        print("\nwarp stall cycles per issued instruction (top reasons per launch)")
        for k, r in enumerate(data):
            vals = sorted(((fnum(r[ci[h]]), h.split("stalled_")[1].split("_per_issue")[0]) for h in st), reverse=True)[:7]
            print("%-4d " % k + "  ".join("%s=%.2f" % (n, v) for v, n in vals))
    tot_t = sum(fnum(g(r, "gpu__time_duration.sum")) for r in data)
    tot_r = sum(fnum(g(r, "dram__bytes_read.sum")) for r in data)
    tot_w = sum(fnum(g(r, "dram__bytes_write.sum")) for r in data)
    print("\nsum over launches: time %.1f %s, dram read %.3f + write %.3f %s" % (tot_t, units[ci["gpu__time_duration.sum"]], tot_r, tot_w,
                                                                          units[ci["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in ci else ""))
    longest = max(range(len(data)), key=lambda k: fnum(g(data[k], "gpu__time_duration.sum"))) if data else 0
    return longest, [g(r, "Kernel Name") for r in data]


def line_map(so, mangled_hint):
    """address -> (file:line of the outermost inlining frame, innermost frame) for the kernel whose demangled name contains the hint"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
        heads = [i for i, l in enumerate(out) if l.startswith(".text.")]
        for i in heads:
            sym = out[i][6:].rstrip(":")
            dem = subprocess.run(["cu++filt", sym], capture_output=True, text=True).stdout.strip()
            if re.sub(r"\s", "", mangled_hint) == re.sub(r"\s", "", dem):
                m, pending, cur = {}, [], (None, None)
                for l in out[i + 1:]:
                    if l.startswith("//---") and ".text." in l:
                        break
                    gq = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
                    if gq:
                        pending.append((gq.group(1).split("/")[-1], int(gq.group(2))))
                        continue
                    a = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", l)
                    if a:
                        if pending:
                            cur = (pending[-1], pending[0])
                            pending = []
                        m[int(a.group(1), 16)] = cur
                return m
    return {}


def source_summary(path, kernel_index, names, so):
    rows = list(csv.reader(gzip.open(path, "rt")))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    if kernel_index >= len(idx) - 1:
        kernel_index = 0
    a, b = idx[kernel_index], idx[kernel_index + 1]
    name = rows[a][1]
    hdr = rows[a + 1]
    ci = {h: j for j, h in enumerate(hdr)}
    data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    print("\n==== launch %d: %s" % (kernel_index, name))
    ops, samp, tot, stot = collections.Counter(), collections.Counter(), 0.0, 0.0
    for r in data:
        s = re.sub(r"^@!?U?P\d+\s+", "", r[ci["Source"]].strip())
        op = s.split()[0].rstrip(";") if s else "?"
        n, sm = fnum(r[ci["Instructions Executed"]]), fnum(r[ci["# Samples"]])
        ops[op] += n
        samp[op] += sm
        tot += n
        stot += sm
    print("warp instructions executed %.0f, stall samples %.0f; opcode mix (>= 0.4%%):" % (tot, stot))
    for op, n in ops.most_common(60):
        if n >= 0.004 * tot:
            print("  %-44s %6.2f%% of instructions  %6.2f%% of samples" % (op, 100 * n / max(tot, 1), 100 * samp[op] / max(stot, 1)))
    if not so:
        return
    m = line_map(so, name)
    if not m:
        print("(no line info for this kernel in %s)" % so)
        return
    base = int(data[0][0], 16)
    scols = [c for c in ("Instructions Executed", "# Samples", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_math",
                         "stall_branch_resolving", "stall_mio", "stall_lg", "stall_not_selected", "L1 Wavefronts Shared") if c in ci]
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in data:
        key = m.get(int(r[0], 16) - base, (None, None))[0]
        for c in scols:
            agg[key][c] += fnum(r[ci[c]])
    print("\nper source line (outermost frame; lines with >= 0.5%% of instructions or samples):")
    print("%-30s" % "line" + "".join("%12s" % c.replace("stall_", "").replace("Instructions Executed", "inst").replace("L1 Wavefronts Shared", "smem_wf")[:11]
                                     for c in scols))
    for key in sorted(agg, key=lambda k: (k is None, k)):
        a_ = agg[key]
        if a_["Instructions Executed"] < 0.005 * tot and a_["# Samples"] < 0.005 * stot:
            continue
        print("%-30s" % ("%s:%d" % key if key else "?") + "".join("%12.0f" % a_[c] for c in scols))


if __name__ == "__main__":
    longest, names = raw_summary(sys.argv[1])
    if len(sys.argv) > 2:
        source_summary(sys.argv[2], longest, names, sys.argv[3] if len(sys.argv) > 3 else None)
