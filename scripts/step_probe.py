#!/usr/bin/env python
"""Where the driver-run number comes from: per-batch step time of the 1M-car city as the run ages (batches of 20, the
driver's --steps), and the cost of a launch that also writes the observation-only columns (the last one of a batch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from traffic_b200 import native, synth

w, _ = synth.make_city(**synth.CITY_CONFIGS["c4_city_1m"])
sim = native.DeviceSim(w, write_aux=True)
def timed(n):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(1e-3, n, 0); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
sim.step(1e-3, 5, 0)
out = {"batches_of_20_us_per_step": [round(timed(20), 2) for _ in range(12)],
       "single_step_batches_us": [round(timed(1), 2) for _ in range(12)],
       "batches_of_200_us_per_step": [round(timed(200), 2) for _ in range(4)]}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/step_probe.json", "w"))
