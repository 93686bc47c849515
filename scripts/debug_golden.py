"""debug: one golden fixture through DeviceSim, print first mismatching cars"""
import sys, numpy as np
sys.path.insert(0, '.')
from tests import util
from traffic_b200 import native
name = sys.argv[1] if len(sys.argv) > 1 else "c1_dt0.001_p2_s1"
g = util.load_golden(name); w = util.world_from_golden(g); start = w.path_start.copy()
sim = native.DeviceSim(w); dt, order = float(g["dt"]), util.order_code(g)
for t in range(3):
    sim.cars_update(dt) if order == 0 else sim.lights_update(dt)
    sim.lights_update(dt) if order == 0 else sim.cars_update(dt)
    sim.download_world(w)
    snap = util.snapshot(w, start)
    bad = np.flatnonzero(np.abs(snap["vx"] - g["vx"][t]) > 1e-6 * np.maximum(1, np.abs(g["vx"][t])))
    print("step", t, "bad cars", bad[:10])
    for i in bad[:5]:
        print(i, "vx", snap["vx"][i], g["vx"][t][i], "d_node", snap["d_node"][i], g["d_node"][t][i], "d_car", snap["d_car"][i], g["d_car"][t][i], "d_light", snap["d_light"][i], g["d_light"][t][i], "f", np.hypot(snap["vx"][i], snap["vy"][i]) / 1500, np.hypot(g["vx"][t][i], g["vy"][t][i]) / 1500)
