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
}


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
