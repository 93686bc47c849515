#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from a `make` log with -Xptxas -v.  usage: ptxas_info.py make.log [filter]"""
import re, subprocess, sys
lines = open(sys.argv[1]).read().splitlines()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
name = None
for i, l in enumerate(lines):
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:70]
        frame = None
    elif name and "bytes stack frame" in l and frame is None:
        frame = l.strip()
    elif name and "Used" in l:
        if flt in name:
            print(name.ljust(72), l.split(":")[1].strip(), "|", frame)
        name = None
