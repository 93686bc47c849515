#!/usr/bin/env python
"""fp32 throughput mode vs fp64 parity mode on the same world (GPU): how often do discrete outputs agree, how far
do positions drift.  usage: python scripts/f32_vs_f64.py [workload] [steps]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_b200 import native, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3_sf_50k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w0, _ = synth.make_city(**synth.CITY_CONFIGS[name])
out = {"workload": name, "cars": w0.n_cars, "checkpoints": []}
a, b = w0.copy(), w0.copy()
s64 = native.DeviceSim(a, precision=native.F64)
s32 = native.DeviceSim(b, precision=native.F32)
done = 0
for chunk in (1, 9, 90, 400, 500, 1000, 3000):
    if done >= steps:
        break
    n = min(chunk, steps - done)
    s64.step(1e-3, n, 1); s32.step(1e-3, n, 1); done += n
    s64.download_world(a); s32.download_world(b)
    dpos = np.abs(a.pos - b.pos)
    rel = (dpos / np.maximum(np.abs(a.pos), 1.0)).max(1)
    rec = {"steps": done,
           "cursor_equal": float((a.cursor == b.cursor).mean()), "cell_equal": float((a.cell == b.cell).mean()),
           "leader_equal": float((a.leader == b.leader).mean()), "face_go_equal": float((a.face_go == b.face_go).mean()),
           "d_light_code_equal": float(((a.d_light != 0) == (b.d_light != 0)).mean()),
           "pos_abs_err_m": {"median": float(np.median(dpos)), "p99": float(np.percentile(dpos, 99)), "max": float(dpos.max())},
           "pos_rel_err": {"p99": float(np.percentile(rel, 99)), "p999": float(np.percentile(rel, 99.9)), "max": float(rel.max())},
           "frac_within_1e-5_rel": float((rel <= 1e-5).mean())}
    out["checkpoints"].append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/f32_vs_f64_%s.json" % name, "w"), indent=1)
