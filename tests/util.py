"""Shared helpers for the parity tests: golden fixtures -> World, and step-by-step comparison."""
from __future__ import annotations

import glob
import os

import numpy as np

from traffic_b200.state import Grid, make_world

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def world_from_golden(g):
    grid = Grid.from_axis(tuple(g["axis"]))
    lights = dict(xy=g["light_xy"], switch_time=g["switch_time"], face_ptr=g["face_ptr"], face_vec=g["face_vec"],
                  face_go=g["face_go0"])
    light_cell = g["light_ybin"] * grid.ncx + g["light_xbin"]
    return make_world(tuple(g["axis"]), g["pos0"], g["path_ptr"], g["path_xy"], g["dest_xy"], lights, light_cell)


def order_code(g):
    return 1 if bytes(g["order"]).decode() == "lights_cars" else 0


def expected_leader(g, t):
    """Leader index implied by the golden record: the first other car in my bin (navigation.py:438-442) when
    distance-to-car is truthy, else -1."""
    xb, yb = g["xbin"][t].astype(np.int64), g["ybin"][t].astype(np.int64)
    key = yb * 100000 + xb
    n = len(key)
    lead = np.full(n, -1, np.int32)
    first, second = {}, {}
    for i, k in enumerate(key):
        if k not in first:
            first[k] = i
        elif k not in second:
            second[k] = i
    for i, k in enumerate(key):
        if g["d_car"][t][i] != 0.0:
            lead[i] = first[k] if first[k] != i else second[k]
    return lead


DISCRETE = ("cursor", "cell", "leader", "face_go", "d_car_truthy", "d_light_code")
FLOATS = ("x", "y", "vx", "vy", "route_time", "d_node", "d_car", "d_light")


def snapshot(w, path_start):
    """Per-step outputs of a World in the golden record's vocabulary."""
    return dict(x=w.pos[:, 0].copy(), y=w.pos[:, 1].copy(), vx=w.vel[:, 0].copy(), vy=w.vel[:, 1].copy(),
                route_time=w.route_time.copy(), d_node=w.d_node.copy(), d_car=w.d_car.copy(), d_light=w.d_light.copy(),
                cursor=(w.cursor - path_start).astype(np.int32), cell=w.cell.copy(), leader=w.leader.copy(),
                face_go=w.face_go.copy())


def light_code(d):
    """0 = False, 1 = None, 2 = a distance."""
    return np.where(d != 0.0, 2, np.where(np.signbit(d), 1, 0)).astype(np.int8)


def compare_step(snap, g, t, ncx, rtol, report):
    """Compare one step against the golden record. Discrete fields must be identical; floats within rtol
    (relative to max(|ref|, 1)).  Appends human-readable mismatches to `report`."""
    ok = True
    ref_cell = (g["ybin"][t].astype(np.int64) * ncx + g["xbin"][t]).astype(np.int32)
    checks = [("cursor", snap["cursor"], g["cursor"][t]), ("cell", snap["cell"], ref_cell),
              ("face_go", snap["face_go"], g["face_go"][t]),
              ("d_car_truthy", snap["d_car"] != 0, g["d_car"][t] != 0),
              ("d_light_code", light_code(snap["d_light"]), light_code(g["d_light"][t])),
              ("leader", snap["leader"], expected_leader(g, t))]
    for name, a, b in checks:
        if not np.array_equal(a, b):
            bad = np.flatnonzero(np.asarray(a) != np.asarray(b))
            report.append("step %d: %s differs at %s (got %s want %s)" % (t, name, bad[:5], np.asarray(a)[bad[:5]], np.asarray(b)[bad[:5]]))
            ok = False
    worst = 0.0
    for name in FLOATS:
        a, b = snap[name], g[name][t]
        both_nan = np.isnan(a) & np.isnan(b)
        err = np.abs(a - b) / np.maximum(np.abs(b), 1.0)
        err[both_nan] = 0.0
        err[np.isnan(err)] = np.inf
        m = float(err.max()) if len(err) else 0.0
        worst = max(worst, m)
        if m > rtol:
            i = int(err.argmax())
            report.append("step %d: %s rel err %.3e at car %d (got %r want %r)" % (t, name, m, i, a[i], b[i]))
            ok = False
    return ok, worst


class FakeNodes(dict):
    pass


class FakeGraph:
    """Duck-typed OGraph for the facade tests: .axis and .G.nodes[node]['x'|'y'] (navigation.py:566-578)."""

    def __init__(self, axis, dest_xy):
        from types import SimpleNamespace
        self.axis = tuple(float(a) for a in axis)
        nodes = FakeNodes({10_000 + i: {"x": float(p[0]), "y": float(p[1])} for i, p in enumerate(dest_xy)})
        self.G = SimpleNamespace(nodes=nodes)
        self.fig = None
        self.ax = SimpleNamespace(axis=lambda: self.axis)


def frames_from_golden(g):
    """Rebuild the reference-schema DataFrames (simulation.py:226-252, 358-378) from a golden fixture."""
    import pandas as pd

    n = len(g["pos0"])
    ptr = g["path_ptr"]
    cars = pd.DataFrame({
        "object": ["car"] * n, "x": g["pos0"][:, 0], "y": g["pos0"][:, 1], "vx": 0, "vy": 0, "route-time": 0,
        "origin": np.arange(n), "destination": 10_000 + np.arange(n),
        "route": [[0, 1]] * n,
        "xpath": [g["path_xy"][ptr[i]:ptr[i + 1], 0].tolist() for i in range(n)],
        "ypath": [g["path_xy"][ptr[i]:ptr[i + 1], 1].tolist() for i in range(n)],
        "distance-to-car": 0, "distance-to-node": 0, "distance-to-red-light": 0})
    grid = Grid.from_axis(tuple(g["axis"]))
    ex, ey = grid.edges()
    cars["xbin"], cars["ybin"] = np.digitize(cars["x"], ex), np.digitize(cars["y"], ey)
    fp = g["face_ptr"]
    L = len(g["switch_time"])
    lights = pd.DataFrame({
        "object": ["light"] * L, "node": np.arange(L), "degree": np.diff(fp), "x": g["light_xy"][:, 0],
        "y": g["light_xy"][:, 1], "switch-counter": 0, "switch-time": g["switch_time"],
        "out-xpositions": [(g["light_xy"][i, 0] + 0.3 * g["face_vec"][fp[i]:fp[i + 1], 0]).tolist() for i in range(L)],
        "out-ypositions": [(g["light_xy"][i, 1] + 0.3 * g["face_vec"][fp[i]:fp[i + 1], 1]).tolist() for i in range(L)],
        "out-xvectors": [g["face_vec"][fp[i]:fp[i + 1], 0].tolist() for i in range(L)],
        "out-yvectors": [g["face_vec"][fp[i]:fp[i + 1], 1].tolist() for i in range(L)],
        "go-values": [g["face_go0"][fp[i]:fp[i + 1]].astype(bool) for i in range(L)],
        "xbin": g["light_xbin"], "ybin": g["light_ybin"]})
    return cars, lights, FakeGraph(g["axis"], g["dest_xy"])
