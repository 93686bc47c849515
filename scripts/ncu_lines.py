#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample totals from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; lines = []; fn = 0
for r in rows:
    if r and r[0] == "Function Name":
        fn += 1
        if fn > 1: break
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0] != "":
        lines.append(r)
def _i(v):
    try:
        return int(v or 0)
    except ValueError:
        return 0
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed")
tot_i = sum(_i(r[ie]) for r in lines); tot_s = sum(_i(r[si]) for r in lines)
print("total warp-inst", tot_i, "samples", tot_s)
for r in sorted(lines, key=lambda r: -_i(r[ie]))[:top]:
    i = _i(r[ie]); t = _i(r[te])
    print(r[0].rjust(4), ("%5.1f%%" % (100.0 * i / tot_i)), ("%5.1f%%s" % (100.0 * _i(r[si]) / max(tot_s, 1))), ("act %4.1f" % (t / max(i, 1))), r[1].strip()[:110])
