#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics per captured launch + stall / opcode breakdown.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
want = [w for w in want if w in hdr]
lines = [want, [units[hdr.index(c)] for c in want]] + [[r[hdr.index(c)] for c in want] for r in rows[2:]]
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        csv.writer(f).writerows(lines)
for i, c in enumerate(want):
    print(c.ljust(70), [l[i] for l in lines[2:]][:3], lines[1][i])

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) < 10:
        break
    data.append(r)
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
agg = {h: sum(int(r[hdr.index(h)] or 0) for r in data) for h in stalls}
print("samples", tot, "warp-inst", sum(int(r[ie] or 0) for r in data))
print("stalls", [(k, round(100.0 * v / max(tot, 1), 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
byop, byinst = collections.Counter(), collections.Counter()
for r in data:
    t = r[1].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += int(r[si] or 0)
    byinst[op] += int(r[ie] or 0)
print("inst by op", byinst.most_common(16))
print("samples by op", byop.most_common(12))
for i, r in sorted(enumerate(data), key=lambda x: -int(x[1][si] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    st = sorted(((h, int(r[hdr.index(h)] or 0)) for h in stalls), key=lambda x: -x[1])[:2]
    print(str(i).rjust(5), r[si].rjust(6), r[ie].rjust(9), r[1].strip()[:64].ljust(64), st)
