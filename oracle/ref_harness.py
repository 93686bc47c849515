"""TEST INFRASTRUCTURE - runs the UNMODIFIED reference (donjpierce/traffic, mounted at /root/reference) as the
ground truth and records golden vectors.  Only usable in the build container (the reference does not travel to
the GPU box); the vectors it writes are committed under tests/golden/ together with this script.

What it does:
* registers a stub module named ``osmnx`` (``navigation.py:10`` imports it at module load; its only use,
  ``ox.stats.count_streets_per_node`` at ``navigation.py:531``, is off the hot path);
* puts /root/reference on ``sys.path`` and imports the reference's own ``cars``, ``simulation`` modules;
* builds the synthetic duck-typed OGraph (``traffic_b200.synth.SynthGraph``), seeds ``random`` and calls the
  reference initialisers ``sim.init_random_node_start_location`` / ``sim.init_traffic_lights``
  (``simulation.py:189-256, 329-381``);
* runs the loop of ``test/profile_cars_update.py:28-34`` (cars then lights) or of ``animate.py:63-64`` /
  ``environment.py:167-168`` (lights then cars) and dumps, per step, every per-car output column.

Usage:  python oracle/ref_harness.py --all            # regenerate every fixture in tests/golden/
        python oracle/ref_harness.py --name c1_dt0.1_p40_s0
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import time
import types
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("TRAFFIC_REFERENCE", "/root/reference")
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def install_reference():
    """Make ``import cars, simulation, navigation, models`` resolve to the reference's own files."""
    if not os.path.isdir(REFERENCE):
        raise RuntimeError("reference tree not found at %s (golden vectors can only be made in the build container)"
                           % REFERENCE)
    if "osmnx" not in sys.modules:
        ox = types.ModuleType("osmnx")
        ox.stats = types.SimpleNamespace(count_streets_per_node=lambda G: {n: G.degree(n) // 2 for n in G.nodes})
        sys.modules["osmnx"] = ox
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    if REPO not in sys.path:
        sys.path.insert(1, REPO)
    import cars as ref_cars  # noqa
    import simulation as ref_sim  # noqa
    assert ref_cars.__file__.startswith(REFERENCE), ref_cars.__file__
    return ref_cars, ref_sim


# name -> (graph kwargs, n_arg for init_random_node_start_location, prescale, dt, steps, order, seed)
CONFIGS = {
    # c1: "Piedmont-sized", 33 cars, the profile script's dt=0.1 / prescale=40 / cars->lights
    "c1_dt0.1_p40_s0":   dict(g=dict(nx_=19, ny_=19, x0=566000.0, y0=4186000.0, seed=0), n=34, prescale=40, dt=0.1, steps=1000, order="cars_lights", seed=0),
    "c1_dt0.001_p40_s0": dict(g=dict(nx_=19, ny_=19, x0=566000.0, y0=4186000.0, seed=0), n=34, prescale=40, dt=0.001, steps=1000, order="cars_lights", seed=0),
    "c1_dt0.001_p2_s1":  dict(g=dict(nx_=19, ny_=19, x0=566000.0, y0=4186000.0, seed=0), n=34, prescale=2, dt=0.001, steps=1000, order="lights_cars", seed=1),
    "c1_dt0.1_p2_s2":    dict(g=dict(nx_=19, ny_=19, x0=566000.0, y0=4186000.0, seed=0), n=34, prescale=2, dt=0.1, steps=400, order="lights_cars", seed=2),
    # dense variant: 120 cars on a 12x12 grid -> many same-cell candidates and red lights
    "dense_dt0.001_p2_s3": dict(g=dict(nx_=12, ny_=12, x0=566000.0, y0=4186000.0, seed=3), n=121, prescale=2, dt=0.001, steps=600, order="lights_cars", seed=3),
    # c2: lower-Manhattan-sized 64x64 grid, 2,000 cars, prescale 10, first 12 steps (the oracle needs ~7 s/step)
    "c2_dt0.001_p10_s0": dict(g=dict(nx_=64, ny_=64, x0=583000.0, y0=4507000.0, seed=0), n=2001, prescale=10, dt=0.001, steps=12, order="lights_cars", seed=0),
    # hand-built: >= 3 cars in a bin with the two lowest indices off the linspace, two lights in one bin (first all green)
    "adversarial_dt0.001":  dict(g=dict(nx_=9, ny_=9, x0=566000.0, y0=4186000.0, seed=5), builder="adversarial", dt=0.001, steps=40, order="lights_cars", seed=0),
}


def adversarial_frames(g):
    """Hand-built frames that pin the first-candidate-only rules of navigation.car_obstacles / light_obstacles
    (navigation.py:442-452, 474-494) independently of any restatement:
      * bin A holds cars 0..4: car 2 has car 3 straight ahead on its linspace, but the first other car of the bin in
        DataFrame order is car 0 (and the next one car 1), both off the linspace -> the reference answers False;
        car 0 (the bin minimum) looks at car 1, which IS ahead of it -> a distance;
      * bin A holds two lights: light 0 lies on car 2's linspace with every face green, light 1 behind it has a red
        anti-parallel face -> the faces of light 0 are exhausted and the distance to light 1 is returned; car 3 has
        light 0 behind it (off its linspace) -> False at once although light 1 qualifies;
      * bin B holds car 5 and one all-green light on its linspace -> None."""
    import pandas as pd
    ax = g.axis
    X, Y = ax[0] + 230.0, ax[2] + 230.0          # inside bin (xbin 2, ybin 2)
    nodes = list(g.G.nodes())
    def car(x, y, path):
        return {"object": "car", "x": x, "y": y, "vx": 0, "vy": 0, "route-time": 0, "origin": nodes[0], "destination": nodes[-1],
                "route": [nodes[0], nodes[-1]], "xpath": [p[0] for p in path], "ypath": [p[1] for p in path],
                "distance-to-car": 0, "distance-to-node": 0, "distance-to-red-light": 0}
    cars = [
        car(X + 10.0, Y + 150.0, [(X + 190.0, Y + 153.0), (X + 420.0, Y + 160.0)]),     # 0: eastbound, top of the bin
        car(X + 100.0, Y + 151.5, [(X + 190.0, Y + 153.0), (X + 420.0, Y + 160.0)]),    # 1: ahead of car 0
        car(X + 20.0, Y + 20.0, [(X + 170.0, Y + 100.0), (X + 400.0, Y + 230.0)]),      # 2: diagonal, cars 0/1 off its line
        car(X + 80.0, Y + 52.0, [(X + 170.0, Y + 100.0), (X + 400.0, Y + 230.0)]),      # 3: straight ahead of car 2
        car(X + 50.0, Y + 36.0, [(X + 170.0, Y + 100.0), (X + 400.0, Y + 230.0)]),      # 4: between 2 and 3
        car(X + 240.0, Y + 30.0, [(X + 380.0, Y + 90.0), (X + 600.0, Y + 200.0)]),      # 5: next bin east, all-green light ahead
    ]
    def light(x, y, vecs, go, sw):
        return {"object": "light", "node": nodes[1], "degree": len(vecs), "x": x, "y": y, "switch-counter": 0, "switch-time": sw,
                "out-xpositions": [x + 0.3 * v[0] for v in vecs], "out-ypositions": [y + 0.3 * v[1] for v in vecs],
                "out-xvectors": [v[0] for v in vecs], "out-yvectors": [v[1] for v in vecs], "go-values": np.array(go)}
    d = np.array([150.0, 80.0]) / np.hypot(150.0, 80.0)
    lights = [
        light(X + 60.0, Y + 41.0, [(60 * d[0], 60 * d[1]), (-60 * d[0], -60 * d[1]), (-40.0, 45.0)], [True, True, True], 0.0),
        light(X + 120.0, Y + 73.0, [(50.0, -30.0), (-55 * d[0] + 1.0, -55 * d[1] - 1.5), (55 * d[0], 55 * d[1])], [True, False, True], 0.013),
        light(X + 310.0, Y + 60.0, [(-70.0, -30.0), (70.0, 30.0)], [True, True], 0.0),
    ]
    cars, lights = pd.DataFrame(cars), pd.DataFrame(lights)
    xb, yb = np.arange(ax[0], ax[1], 200), np.arange(ax[2], ax[3], 200)
    cars["xbin"], cars["ybin"] = np.digitize(cars["x"], xb), np.digitize(cars["y"], yb)
    lights["xbin"], lights["ybin"] = np.digitize(lights["x"], xb), np.digitize(lights["y"], yb)
    return cars, lights


def _falsy_encode(values, none_as_negzero=False):
    out = np.zeros(len(values), np.float64)
    for i, v in enumerate(values):
        if v is None:
            out[i] = -0.0 if none_as_negzero else 0.0
        elif v is False or (isinstance(v, (bool, np.bool_)) and not v):
            out[i] = 0.0
        else:
            out[i] = float(v)
    return out


def run_reference(cfg, verbose=True):
    from traffic_b200.synth import SynthGraph

    ref_cars, ref_sim = install_reference()
    warnings.simplefilter("ignore")
    g = SynthGraph(**cfg["g"])
    random.seed(cfg["seed"])
    t0 = time.time()
    if cfg.get("builder") == "adversarial":
        init_cars, init_lights = adversarial_frames(g)
    else:
        init_cars = ref_sim.init_random_node_start_location(cfg["n"], g)
        init_lights = ref_sim.init_traffic_lights(g, prescale=cfg["prescale"])
    t_init = time.time() - t0
    cars = ref_cars.Cars(init_state=init_cars, graph=g)
    lights = ref_cars.TrafficLights(light_state=init_lights, graph=g)
    N, L = len(init_cars), len(init_lights)
    T, dt = cfg["steps"], cfg["dt"]
    init_len = np.asarray([len(p) for p in init_cars["xpath"]], np.int64)

    rec = {k: np.zeros((T, N), np.float64) for k in ("x", "y", "vx", "vy", "route_time", "d_node", "d_car", "d_light")}
    rec.update({k: np.zeros((T, N), np.int32) for k in ("cursor", "xbin", "ybin")})
    F = int(sum(init_lights["degree"]))
    face_go = np.zeros((T, F), np.uint8)
    t0 = time.time()
    for t in range(T):
        if cfg["order"] == "lights_cars":
            lights.update(dt)
            st = cars.update(dt, lights.state)
        else:
            st = cars.update(dt, lights.state)
            lights.update(dt)
        rec["x"][t] = st["x"].to_numpy(np.float64)
        rec["y"][t] = st["y"].to_numpy(np.float64)
        rec["vx"][t] = st["vx"].to_numpy(np.float64)
        rec["vy"][t] = st["vy"].to_numpy(np.float64)
        rec["route_time"][t] = st["route-time"].to_numpy(np.float64)
        rec["d_node"][t] = st["distance-to-node"].to_numpy(np.float64)
        rec["d_car"][t] = _falsy_encode(list(st["distance-to-car"]))
        rec["d_light"][t] = _falsy_encode(list(st["distance-to-red-light"]), none_as_negzero=True)
        rec["cursor"][t] = init_len - np.asarray([len(p) for p in st["xpath"]], np.int64)
        rec["xbin"][t] = st["xbin"].to_numpy(np.int64)
        rec["ybin"][t] = st["ybin"].to_numpy(np.int64)
        face_go[t] = np.concatenate([np.asarray(v).astype(np.uint8) for v in lights.state["go-values"]])
        if verbose and (t + 1) % 100 == 0:
            print("  step %d/%d  %.1fs" % (t + 1, T, time.time() - t0), flush=True)
    t_loop = time.time() - t0

    # inputs in flat form (so the GPU-side tests need neither pandas DataFrames of lists nor networkx)
    from traffic_b200.state import world_from_dataframes

    w0 = world_from_dataframes(init_cars, init_lights, g)
    out = dict(
        axis=np.asarray(g.axis, np.float64), dt=np.float64(dt), steps=np.int64(T),
        order=np.bytes_(cfg["order"]), seed=np.int64(cfg["seed"]),
        pos0=w0.pos, path_ptr=np.concatenate([w0.path_start.astype(np.int64), [int(w0.path_end[-1])]]),
        path_xy=w0.path_xy, dest_xy=w0.dest_xy,
        light_xy=w0.light_xy, switch_time=w0.switch_time, face_ptr=w0.face_ptr, face_vec=w0.face_vec,
        face_go0=w0.face_go,
        light_xbin=init_lights["xbin"].to_numpy(np.int64), light_ybin=init_lights["ybin"].to_numpy(np.int64),
        face_go=face_go, ref_loop_seconds=np.float64(t_loop), ref_init_seconds=np.float64(t_init),
        ref_updates_per_s=np.float64(N * T / t_loop),
    )
    out.update(rec)
    if verbose:
        print("  reference: %d cars x %d steps in %.1fs -> %.0f car-updates/s (init %.1fs)"
              % (N, T, t_loop, N * T / t_loop, t_init), flush=True)
    return out, init_cars, init_lights, g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--name", action="append", default=[])
    args = ap.parse_args()
    names = list(CONFIGS) if args.all else args.name
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in names:
        print("[golden] %s" % name, flush=True)
        out, _, _, _ = run_reference(CONFIGS[name])
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print("  wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024), flush=True)


if __name__ == "__main__":
    main()
