#!/usr/bin/env python
"""Runs the workload BASELINE.json's north_star names (the update loop the reference profiles: cars first, then lights,
dt = 0.1, about a hundred cars on a Piedmont-sized map) on the drop-in classes, with a synthetic graph instead of an
OSMnx download and the GPU-batched initialisers instead of two networkx Dijkstras per car.

    python examples/run_profile_workload.py --steps 1000 [--cars 100] [--seed 0]

The reference's own harness runs unchanged too when traffic_b200/dropin precedes the reference checkout on PYTHONPATH
(INTEGRATION.md, section 1).
"""
import argparse
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "traffic_b200", "dropin"), os.path.join(HERE, "..")]

import cars as dropin_cars                      # noqa: E402  -> traffic_b200/dropin/cars.py
from traffic_b200 import init, synth            # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--cars", type=int, default=100, help="n of init_random_node_start_location (yields n-1 cars)")
    ap.add_argument("--seed", type=int, default=0)
    opt = ap.parse_args()

    world = synth.SynthGraph(19, 19, seed=0)    # 361 nodes
    random.seed(opt.seed)
    t_init = time.perf_counter()
    fleet = dropin_cars.Cars(init.init_random_node_start_location(opt.cars, world), world)
    signals = dropin_cars.TrafficLights(init.init_traffic_lights(world, prescale=40), world)
    t_loop = time.perf_counter()
    state = None
    for _ in range(opt.steps):
        state = fleet.update(0.1, signals.state)
        signals.update(0.1)
    xs = state["x"].to_numpy()                  # first touch of the state: one device -> host refresh
    t_end = time.perf_counter()
    n = len(xs)
    print("init %.3f s; %d cars x %d steps in %.3f s = %.0f vehicle-updates/s (%.1f us per step); parked: %d"
          % (t_loop - t_init, n, opt.steps, t_end - t_loop, n * opt.steps / (t_end - t_loop),
             (t_end - t_loop) / opt.steps * 1e6, int((state["vx"].to_numpy() == 0).sum())))


if __name__ == "__main__":
    main()
