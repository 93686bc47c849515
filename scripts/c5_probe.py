"""probe: c5 rollout behaviour - how many of R replicas see the agent arrive within max_steps, and at what cost"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from traffic_b200 import synth, ensemble, native
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
agent = int(sys.argv[3]) if len(sys.argv) > 3 else 17
cfg = dict(synth.CITY_CONFIGS["c2_manhattan_2k"]); cfg["seed"] = 0
w, _ = synth.make_city(**cfg)
print("agent route points", w.path_end[agent] - w.path_start[agent])
ens = ensemble.Ensemble(w, np.arange(R), agent_car=agent, seed=0)
torch.cuda.synchronize(); t0 = time.time()
res = ens.rollout(1e-3, max_steps)
torch.cuda.synchronize(); el = time.time() - t0
print("R", R, "arrived", int(res[:, 2].sum()), "steps min/med/max", res[:, 1].min(), np.median(res[:, 1]), res[:, 1].max(), "time %.3f s" % el,
      "car-steps/s %.3g" % (w.n_cars * res[:, 1].sum() / el), "launches", ens.sim.launch_count)
