#!/usr/bin/env python
"""Full SASS listing of the first kernel in an .ncu-rep with per-instruction executed counts, active lanes and stall
samples, annotated with the CUDA source line taken from `nvdisasm -g` of the object file the kernel was built from
(must be the same build).
usage: python scripts/ncu_sass.py rep obj mangled_substring [out.txt]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, obj, fn = sys.argv[1], sys.argv[2], sys.argv[3]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) < 10:
        break
    data.append(r)
si, ie, te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][0], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, None, False
for l in dis:
    if l.startswith(".text."):
        inside = fn in l and "$" not in l.split(fn)[-1][:40].replace("E:", "")
        inside = l.strip().rstrip(":").endswith(fn) or (fn in l and l.strip().endswith("E:"))
        continue
    if not inside:
        continue
    m = re.search(r'line (\d+)', l)
    if "//## File" in l and m:
        cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = open(sys.argv[4], "w") if len(sys.argv) > 4 else sys.stdout
tot = sum(int(r[ie] or 0) for r in data)
for k, r in enumerate(data):
    i = int(r[ie] or 0); t = int(r[te] or 0)
    off = int(r[0], 16) - base
    st = sorted(((h[6:], int(r[hdr.index(h)] or 0)) for h in stalls), key=lambda x: -x[1])[:2]
    st = " ".join("%s:%d" % s for s in st if s[1] > 0)
    out.write("%5d %04x %9d %5.1f %4d L%-5s %-60s %s\n" % (k, off, i, t / max(i, 1), int(r[si] or 0), line_of.get(off, "?"), r[1].strip()[:60], st))
out.write("total %d\n" % tot)
