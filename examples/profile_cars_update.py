#!/usr/bin/env python
"""The workload BASELINE.json's north_star names - the loop of the reference's test/profile_cars_update.py:16-36 - on the
drop-in classes: same calls, same order (cars then lights, dt = 0.1), a synthetic Piedmont-sized graph instead of an
OSMnx download, the GPU-batched initialisers instead of two networkx Dijkstras per car.

    python examples/profile_cars_update.py -t 1000 [-n 100] [--profile]

With the reference checkout on PYTHONPATH *after* traffic_b200/dropin, the reference's own script runs unchanged as well
(INTEGRATION.md, section 1).
"""
import argparse
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "traffic_b200", "dropin"))
sys.path.insert(1, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from cars import Cars, TrafficLights          # noqa: E402  (resolves to traffic_b200/dropin/cars.py)
from traffic_b200 import init as sim          # noqa: E402  (drop-in for the reference's `simulation` initialisers)
from traffic_b200.synth import SynthGraph     # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("-t", "--timesteps", type=int, default=100)
parser.add_argument("-n", "--cars", type=int, default=100, help="n for init_random_node_start_location (n-1 cars)")
parser.add_argument("--profile", action="store_true")


def test_cars_update(timesteps, n, graph):
    init_cars = sim.init_random_node_start_location(n, graph)
    cars_object = Cars(init_state=init_cars, graph=graph)
    init_lights = sim.init_traffic_lights(graph, prescale=40)
    lights_object = TrafficLights(light_state=init_lights, graph=graph)
    cars_data, lights_data = [], []
    for i in range(timesteps):
        cars_data.append(cars_object.update(dt=0.1, lights=lights_object.state))
        lights_data.append(lights_object.update(dt=0.1))
    return {"cars": cars_data, "lights": lights_data}, cars_object


if __name__ == "__main__":
    args = parser.parse_args()
    graph = SynthGraph(19, 19, seed=0)        # "Piedmont-sized": 361 nodes
    random.seed(0)
    if args.profile:
        profiler = cProfile.Profile()
        profiler.enable()
    t0 = time.perf_counter()
    data, cars_object = test_cars_update(args.timesteps, args.cars, graph)
    x = data["cars"][-1]["x"].to_numpy()      # touching the state materialises it from the device
    el = time.perf_counter() - t0
    if args.profile:
        profiler.disable()
        pstats.Stats(profiler).sort_stats("cumtime").print_stats(25)
    print("%d cars x %d steps in %.3f s (init included) -> %.0f vehicle-updates/s; mean x = %.3f"
          % (len(x), args.timesteps, el, len(x) * args.timesteps / el, x.mean()))
