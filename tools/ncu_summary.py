"""Summarise an .ncu-rep (read here, no GPU): per-launch table + stall breakdown + opcode mix of one kernel.
usage: python tools/ncu_summary.py rep.ncu-rep [kernel-id-for-source]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(n): return hdr.index(n)
cols = [("gpu__time_duration.sum", "ms"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smemKB"),
        ("launch__occupancy_limit_shared_mem", "occ_sm"), ("launch__occupancy_limit_registers", "occ_rg"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "inst"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("lts__t_sectors_op_red.sum", "l2red"), ("lts__t_bytes.sum", "l2B")]
print(" ".join("%10s" % c[1] for c in cols))
for r in data:
    out = []
    for name, _ in cols:
        out.append(r[col(name)][:10] if name in hdr else "-")
    print(" ".join("%10s" % o for o in out))
# stall reasons (pct of warp-active) for each launch
st = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
if st:
    print("\nstall cycles per issued instruction (top reasons)")
    for k, r in enumerate(data):
        vals = sorted(((float(r[col(h)] or 0), h.split("stalled_")[1].split("_per_issue")[0]) for h in st), reverse=True)[:6]
        print(k, " ".join("%s=%.2f" % (n, v) for v, n in vals))
if len(sys.argv) > 2:
    kid = sys.argv[2]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h2 = rows[1]
    ia, ie, isamp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
    seen, ops, tot, samples = set(), collections.Counter(), 0, collections.Counter()
    stall_cols = [i for i, n in enumerate(h2) if n.startswith("stall_") and "Not Issued" not in n]
    stall_tot = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ie or r[0] in seen: continue
        seen.add(r[0])
        try: n = int(r[ie])
        except ValueError: continue
        toks = r[ia].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += n; tot += n
        for i in stall_cols:
            try: stall_tot[h2[i]] += int(r[i])
            except ValueError: pass
    print("\nkernel", kid, "warp instructions", tot)
    print(" ".join("%s=%.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(16)))
    ts = sum(stall_tot.values())
    print("stall samples:", " ".join("%s=%.1f%%" % (k[6:], 100 * v / ts) for k, v in stall_tot.most_common(8)))
