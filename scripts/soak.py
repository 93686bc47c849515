#!/usr/bin/env python
"""Long run of the 1M-car city: step cost and state statistics as the simulation ages (cars queue, then park)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from traffic_b200 import native, synth

w, _ = synth.make_city(**synth.CITY_CONFIGS["c4_city_1m"])
sim = native.DeviceSim(w)
out, done = [], 0
for chunk in (1000, 1000, 1000, 2000, 5000, 10000):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(1e-3, chunk, 0); e1.record(); torch.cuda.synchronize()
    done += chunk
    sim.download_world(w)
    rec = {"steps": done, "us_per_step": e0.elapsed_time(e1) / chunk * 1e3, "parked": float((w.cursor >= w.path_end).mean()),
           "leaders": float((w.leader >= 0).mean()), "red": float((w.d_light != 0).mean()), "stopped": float((np.abs(w.vel).sum(1) < 1e-9).mean()),
           "nan": int(np.isnan(w.pos).any(1).sum())}
    out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/soak_c4.json", "w"), indent=1)
