"""The reference's own, UNMODIFIED callers on top of the drop-in ``cars`` module ("artist.py,
animate.py and learn.py run unchanged on top of it").

``traffic_b200/dropin`` is put ahead of /root/reference on ``sys.path``, so the reference's ``from cars import Cars,
TrafficLights`` (environment.py:2, test/profile_cars_update.py:5) binds the drop-in classes; then
  * the loop of ``test/profile_cars_update.py:16-36`` (``test_cars_update``),
  * ``environment.Env`` construction + ``Env.simulation_step`` (environment.py:11-41, 154-173),
  * ``animate.Animator.animate`` (animate.py:56-98) with stub matplotlib artists
run exactly as the reference ships them, and what they produce is compared with the same callers driven by the reference's
own ``cars.py`` (loaded under another module name).

These tests need the reference tree, which exists only in the build container - where there is no GPU; so the device
behind the drop-in classes is tests/oracle_device.py (the CPU oracle behind the DeviceSim interface; the oracle is pinned
bit-exact to the reference and the CUDA path is pinned to the oracle by the -m gpu tests).  What is proven here is the
boundary: constructors, attributes, DataFrame behaviour (``state.loc[agent]``, ``iterrows``, ``state['degree']``, ``len``),
call order and return conventions under the reference's real drivers.  tests/test_facade.py drives the same call patterns
through the CUDA library on the GPU box."""
import importlib.util
import os
import random
import sys
import types

import numpy as np
import pytest

REFERENCE = os.environ.get("TRAFFIC_REFERENCE", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")


@pytest.fixture()
def ref(monkeypatch):
    """Reference modules importable, `cars` = the drop-in, DeviceSim = the oracle-backed look-alike."""
    from tests.oracle_device import OracleDeviceSim
    from traffic_b200 import native, synth
    for m in ("cars", "simulation", "navigation", "models", "environment", "animate", "osm_request", "profile_cars_update"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    ox = types.ModuleType("osmnx")
    ox.stats = types.SimpleNamespace(count_streets_per_node=lambda G: {n: G.degree(n) // 2 for n in G.nodes})
    monkeypatch.setitem(sys.modules, "osmnx", ox)
    osm = types.ModuleType("osm_request")           # osm_request.py needs osmnx + matplotlib + a network
    osm.OGraph = lambda *a, **k: synth.SynthGraph(12, 12, seed=3)
    monkeypatch.setitem(sys.modules, "osm_request", osm)
    monkeypatch.syspath_prepend(REFERENCE)
    monkeypatch.syspath_prepend(os.path.join(REFERENCE, "test"))
    monkeypatch.syspath_prepend(os.path.join(REPO, "traffic_b200", "dropin"))   # first: shadows the reference's cars.py
    monkeypatch.setattr(native, "DeviceSim", OracleDeviceSim)
    monkeypatch.setattr(native, "lib", lambda: None)
    import cars as dropin_cars
    assert "traffic_b200" in dropin_cars.Cars.__module__
    spec = importlib.util.spec_from_file_location("reference_cars", os.path.join(REFERENCE, "cars.py"))
    ref_cars = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_cars)                # the reference's own classes, as ground truth
    return types.SimpleNamespace(cars=dropin_cars, ref_cars=ref_cars, graph=synth.SynthGraph(12, 12, seed=3))


def _same_cars(a, b, what):
    for col in ("x", "y", "vx", "vy", "route-time", "distance-to-node"):
        assert np.allclose(np.asarray(a[col], float), np.asarray(b[col], float), rtol=1e-12, atol=0, equal_nan=True), (what, col)
    for col in ("xbin", "ybin"):
        assert np.array_equal(np.asarray(a[col]), np.asarray(b[col])), (what, col)
    for col in ("distance-to-car", "distance-to-red-light"):
        for u, v in zip(a[col], b[col]):
            assert (u is None) == (v is None) and bool(u) == bool(v), (what, col, u, v)
            if u:
                assert abs(u - v) <= 1e-9 * abs(v)
    assert [len(p) for p in a["xpath"]] == [len(p) for p in b["xpath"]], what


def test_profile_cars_update_loop_runs_unchanged(ref):
    import profile_cars_update as prof            # the reference's test/profile_cars_update.py
    import simulation as sim
    assert prof.Cars is ref.cars.Cars and prof.Cars is not ref.ref_cars.Cars
    prof.graph = ref.graph
    random.seed(4)
    data = prof.test_cars_update(timesteps=60)    # unmodified: init, 60 x (cars.update, lights.update)
    assert len(data["cars"]) == 60 and data["cars"][0] is data["cars"][-1]    # update returns self.state (cars.py:95)
    # ground truth: the same function body on the reference's own classes
    random.seed(4)
    cars_object = ref.ref_cars.Cars(init_state=sim.init_random_node_start_location(100, ref.graph), graph=ref.graph)
    lights_object = ref.ref_cars.TrafficLights(light_state=sim.init_traffic_lights(ref.graph, prescale=40), graph=ref.graph)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(60):
            cars_object.update(dt=0.1, lights=lights_object.state)
            lights_object.update(dt=0.1)
    _same_cars(data["cars"][-1], cars_object.state, "profile loop")
    for g_ours, g_ref in zip(data["lights"][-1]["go-values"], lights_object.state["go-values"]):
        assert np.array_equal(np.asarray(g_ours, bool), np.asarray(g_ref, bool))


def test_env_simulation_step_runs_unchanged_and_reads_one_row(ref):
    import environment
    from traffic_b200 import native
    assert environment.Cars is ref.cars.Cars
    random.seed(9)
    env = environment.Env(n=40, graph=ref.graph, agent=7, dt=1e-3)      # Env.__init__ unmodified (prescale=40 lights)
    sim_obj = None
    for i in range(250):
        arrived = env.simulation_step(i)                                 # FrontView(state.loc[agent]) + updates
        assert arrived in (True, False)
        sim_obj = env.cars_object._sim or sim_obj
    # state.loc[agent] must cost a few elements per field, never a full-state download
    n = env.cars_object._n_rows
    assert sim_obj is not None and sim_obj.downloads
    assert max(cnt for f, cnt in sim_obj.downloads if f != native.FACE_GO) <= 2, "Env.simulation_step pulled whole columns"
    # ground truth: the same Env driven by the reference's own classes
    random.seed(9)
    environment.Cars, environment.TrafficLights = ref.ref_cars.Cars, ref.ref_cars.TrafficLights
    try:
        env_ref = environment.Env(n=40, graph=ref.graph, agent=7, dt=1e-3)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(250):
                env_ref.simulation_step(i)
    finally:
        environment.Cars, environment.TrafficLights = ref.cars.Cars, ref.cars.TrafficLights
    _same_cars(env.cars_object.state, env_ref.cars_object.state, "Env.simulation_step")
    assert abs(env.cars_object.state.loc[7]["route-time"] - env_ref.cars_object.state.loc[7]["route-time"]) < 1e-12
    assert len(sim_obj.downloads) > 0 and n == 39


class _Artist:
    def __init__(self):
        self.data, self.color = None, None

    def set_data(self, x, y):
        self.data = (x, y)

    def set_color(self, c):
        self.color = c


def _stub_axes():
    ax = types.SimpleNamespace(plot=lambda *a, **k: [_Artist()], axis=lambda: (0, 1, 0, 1), annotate=lambda *a, **k: None,
                               set_xlim=lambda *a: None, set_ylim=lambda *a: None)
    fig = types.SimpleNamespace(canvas=types.SimpleNamespace(draw=lambda: None))
    return fig, ax


def test_animator_animate_runs_unchanged(ref):
    import animate
    import simulation as sim
    outs = []
    for mod in (ref.cars, ref.ref_cars):
        random.seed(2)
        cars_object = mod.Cars(init_state=sim.init_random_node_start_location(30, ref.graph), graph=ref.graph)
        lights_object = mod.TrafficLights(light_state=sim.init_traffic_lights(ref.graph, prescale=10), graph=ref.graph)
        fig, ax = _stub_axes()
        anim = animate.Animator(fig, ax, cars_object, lights_object, num=(1, 1), n=len(cars_object.state))   # animate.py:5-21
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(40):
                artists = anim.animate(i)                                  # lights.update, cars.update, iterrows (56-98)
        outs.append([(a.data, a.color) for a in artists])
    ours, theirs = outs
    assert len(ours) == len(theirs) > 30
    for (d1, c1), (d2, c2) in zip(ours, theirs):
        assert c1 == c2
        assert np.allclose(np.asarray(d1, float), np.asarray(d2, float), rtol=1e-12, atol=0)
