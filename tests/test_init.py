"""The drop-in initialisers (traffic_b200/init.py: GPU-batched routing behind the reference's init API) must rebuild,
from the same `random` seed, exactly the inputs the reference's own initialisers produced for the golden fixtures."""
import random

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1_dt0.1_p40_s0", "c1_dt0.001_p2_s1", "dense_dt0.001_p2_s3", "c2_dt0.001_p10_s0"])
def test_init_reproduces_reference_inputs(name):
    from oracle.ref_harness import CONFIGS
    from traffic_b200 import init, synth
    from traffic_b200.state import world_from_dataframes

    cfg = CONFIGS[name]
    graph = synth.SynthGraph(**cfg["g"])
    random.seed(cfg["seed"])
    cars_df = init.init_random_node_start_location(cfg["n"], graph)
    lights_df = init.init_traffic_lights(graph, prescale=cfg["prescale"])
    w = world_from_dataframes(cars_df, lights_df, graph)
    g = util.load_golden(name)
    w2 = util.world_from_golden(g)
    for f in ("pos", "path_xy", "dest_xy", "path_start", "path_end", "light_xy", "switch_time", "face_ptr", "face_vec",
              "face_go", "cell_light_ptr", "cell_light_idx", "cell"):
        assert np.array_equal(getattr(w, f), getattr(w2, f)), "%s: field %s differs from the reference's init" % (name, f)
    # schema
    assert list(cars_df.columns) == ["object", "x", "y", "vx", "vy", "route-time", "origin", "destination", "route", "xpath",
                                     "ypath", "distance-to-car", "distance-to-node", "distance-to-red-light", "xbin", "ybin"]
    assert list(lights_df.columns) == ["object", "node", "degree", "x", "y", "switch-counter", "switch-time", "out-xpositions",
                                       "out-ypositions", "out-xvectors", "out-yvectors", "go-values", "xbin", "ybin"]


def test_init_then_facade_loop_matches_golden_end_to_end():
    """random.seed -> our init -> our Cars/TrafficLights -> 200 steps == the reference's recorded run: the whole
    reference-facing surface (init + update) with no reference code in the loop."""
    from oracle.ref_harness import CONFIGS
    from traffic_b200 import init, synth
    from traffic_b200.facade import Cars, TrafficLights

    name = "c1_dt0.001_p2_s1"
    cfg = CONFIGS[name]
    g = util.load_golden(name)
    graph = synth.SynthGraph(**cfg["g"])
    random.seed(cfg["seed"])
    cars = Cars(init.init_random_node_start_location(cfg["n"], graph), graph)
    lights = TrafficLights(init.init_traffic_lights(graph, prescale=cfg["prescale"]), graph)
    for t in range(200):
        lights.update(cfg["dt"])
        st = cars.update(cfg["dt"], lights.state)
    assert np.allclose(st["x"].to_numpy(float), g["x"][199], rtol=1e-12, atol=0)
    assert np.allclose(st["y"].to_numpy(float), g["y"][199], rtol=1e-12, atol=0)
    go = np.concatenate([np.asarray(v).astype(np.uint8) for v in lights.state["go-values"]])
    assert np.array_equal(go, g["face_go"][199])


def test_init_on_a_plain_networkx_graph_equals_array_path():
    """An OGraph-like object that only has `.G` (networkx MultiDiGraph) and `.axis` - what the reference's callers pass -
    must give the same frames as the SynthGraph fast path."""
    from types import SimpleNamespace
    from traffic_b200 import init, synth

    graph = synth.SynthGraph(12, 12, seed=3)
    plain = SimpleNamespace(G=graph.G, axis=graph.axis)
    random.seed(7)
    c1 = init.init_random_node_start_location(60, graph)
    l1 = init.init_traffic_lights(graph, prescale=3)
    random.seed(7)
    c2 = init.init_random_node_start_location(60, plain)
    l2 = init.init_traffic_lights(plain, prescale=3)
    for col in ("x", "y", "origin", "destination", "xbin", "ybin"):
        assert np.array_equal(c1[col].to_numpy(), c2[col].to_numpy()), col
    assert all(a == b for a, b in zip(c1["xpath"], c2["xpath"])) and all(a == b for a, b in zip(c1["route"], c2["route"]))
    for col in ("node", "degree", "x", "y", "switch-time", "xbin", "ybin"):
        assert np.array_equal(l1[col].to_numpy(), l2[col].to_numpy()), col
    assert all(a == b for a, b in zip(l1["out-xvectors"], l2["out-xvectors"]))


def test_init_culdesac_reproduces_the_reference_fixture():
    """init_culdesac_start_location (simulation.py:259-326) against the frame the unmodified reference built on the same
    graph (oracle/ref_culdesac.py -> tests/golden/culdesac_init.json)."""
    import json
    import os
    from traffic_b200 import init, synth

    fx = json.load(open(os.path.join(util.GOLDEN_DIR, "culdesac_init.json")))
    graph = synth.SynthGraph(**fx["graph"])
    cars = init.init_culdesac_start_location(fx["n"], graph)
    assert len(cars) == len(fx["cars"]) == fx["n"]
    for (_, row), want in zip(cars.iterrows(), fx["cars"]):
        assert row["origin"] == want["origin"] and row["destination"] == want["destination"]
        assert list(row["route"]) == want["route"]
        assert row["xpath"] == want["xpath"] and row["ypath"] == want["ypath"]
        assert row["x"] == want["x"] and row["y"] == want["y"] and row["xbin"] == want["xbin"] and row["ybin"] == want["ybin"]
    with pytest.raises(ValueError):
        init.init_culdesac_start_location(fx["graph"]["culdesacs"] + 1, graph)
