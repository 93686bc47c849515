"""The drop-in Cars / TrafficLights classes (traffic_b200/facade.py, importable as module `cars` through
traffic_b200/dropin) driven exactly like the reference's callers drive cars.py, checked against the golden record."""
import importlib
import os
import sys

import numpy as np
import pytest

from tests import util

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_module_is_named_cars_and_has_reference_api():
    sys.path.insert(0, os.path.join(REPO, "traffic_b200", "dropin"))
    try:
        sys.modules.pop("cars", None)
        cars = importlib.import_module("cars")
        assert cars.__file__.startswith(os.path.join(REPO, "traffic_b200", "dropin"))
        for cls, methods in ((cars.Cars, ("update", "find_obstacles", "write_state")), (cars.TrafficLights, ("update",))):
            for m in methods:
                assert callable(getattr(cls, m))
        import inspect
        assert list(inspect.signature(cars.Cars.__init__).parameters) == ["self", "init_state", "graph", "serialize"]
        assert list(inspect.signature(cars.Cars.update).parameters) == ["self", "dt", "lights"]
        assert list(inspect.signature(cars.TrafficLights.__init__).parameters) == ["self", "light_state", "graph"]
        assert list(inspect.signature(cars.TrafficLights.update).parameters) == ["self", "dt"]
    finally:
        sys.path.pop(0)
        sys.modules.pop("cars", None)


def test_facade_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from traffic_b200 import facade, native
    g = util.load_golden(util.golden_names()[0])
    cars_df, lights_df, graph = util.frames_from_golden(g)
    with pytest.raises(native.NativeError):
        facade.TrafficLights(lights_df, graph)


def test_frames_roundtrip_through_packer():
    """DataFrame -> World packing reproduces the flat arrays stored in the fixture."""
    from traffic_b200.state import world_from_dataframes
    for name in util.golden_names()[:2]:
        g = util.load_golden(name)
        cars_df, lights_df, graph = util.frames_from_golden(g)
        w = world_from_dataframes(cars_df, lights_df, graph)
        w2 = util.world_from_golden(g)
        for f in ("pos", "path_xy", "dest_xy", "path_start", "path_end", "path_eff_end", "light_xy", "switch_time",
                  "face_ptr", "face_vec", "face_go", "cell_light_ptr", "cell_light_idx", "cell"):
            assert np.array_equal(getattr(w, f), getattr(w2, f)), f


def _state_snapshot(cars_obj, lights_obj, g, start_len):
    st = cars_obj.state
    dc = np.array([0.0 if (v is False or v is None) else float(v) for v in st["distance-to-car"]])
    dl = np.array([(-0.0 if v is None else 0.0) if (v is False or v is None) else float(v) for v in st["distance-to-red-light"]])
    ncx = len(np.arange(g["axis"][0], g["axis"][1], 200)) + 1
    lead = np.full(len(dc), -1, np.int32)
    return dict(x=st["x"].to_numpy(float), y=st["y"].to_numpy(float), vx=st["vx"].to_numpy(float),
                vy=st["vy"].to_numpy(float), route_time=st["route-time"].to_numpy(float),
                d_node=st["distance-to-node"].to_numpy(float), d_car=dc, d_light=dl,
                cursor=(start_len - np.array([len(p) for p in st["xpath"]])).astype(np.int32),
                cell=(st["ybin"].to_numpy(np.int64) * ncx + st["xbin"].to_numpy(np.int64)).astype(np.int32),
                leader=lead,
                face_go=np.concatenate([np.asarray(v).astype(np.uint8) for v in lights_obj.state["go-values"]]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_dt0.001_p2_s1", "c1_dt0.1_p40_s0", "dense_dt0.001_p2_s3"])
def test_facade_loop_matches_reference_golden(name):
    from traffic_b200.facade import Cars, TrafficLights
    g = util.load_golden(name)
    cars_df, lights_df, graph = util.frames_from_golden(g)
    cars_obj = Cars(init_state=cars_df, graph=graph)
    lights_obj = TrafficLights(light_state=lights_df, graph=graph)
    assert len(cars_obj.state) == len(cars_df) and cars_obj.state["x"][0] == cars_df["x"][0]
    start_len = np.array([len(p) for p in cars_df["xpath"]])
    dt, T, order = float(g["dt"]), min(int(g["steps"]), 300), util.order_code(g)
    report = []
    for t in range(T):
        if order == 1:      # animate.py:63-64 / environment.py:167-168
            lights_obj.update(dt)
            ret = cars_obj.update(dt, lights_obj.state)
        else:               # test/profile_cars_update.py:29-34
            ret = cars_obj.update(dt=dt, lights=lights_obj.state)
            lights_obj.update(dt=dt)
        assert ret is cars_obj.state
        if t < 10 or t % 25 == 0 or t == T - 1:
            snap = _state_snapshot(cars_obj, lights_obj, g, start_len)
            snap["leader"] = util.expected_leader(g, t)  # the DataFrame has no leader column
            ok, _ = util.compare_step(snap, g, t, 0, 1e-9, report) if False else util.compare_step(
                snap, g, t, len(np.arange(g["axis"][0], g["axis"][1], 200)) + 1, 1e-9, report)
            assert ok, "\n".join(report[:6])
    assert abs(cars_obj.time_elapsed - sum([dt] * T)) < 1e-9
    assert np.isnan(cars_obj.state["route"].astype(float)).all()   # Appendix B-7


@pytest.mark.gpu
def test_facade_accepts_foreign_lights_dataframe():
    """Cars.update(dt, lights) with a plain DataFrame (e.g. the reference's own TrafficLights.state): the go-values
    are taken from the DataFrame every step."""
    from traffic_b200.facade import Cars
    name = "c1_dt0.001_p2_s1"
    g = util.load_golden(name)
    cars_df, lights_df, graph = util.frames_from_golden(g)
    cars_obj = Cars(cars_df, graph)
    fp = g["face_ptr"]
    start_len = np.array([len(p) for p in cars_df["xpath"]])
    report = []
    for t in range(60):
        go = g["face_go"][t]  # order lights->cars: the state the reference's cars saw at step t
        lights_df["go-values"] = [go[fp[i]:fp[i + 1]].astype(bool) for i in range(len(lights_df))]
        cars_obj.update(float(g["dt"]), lights_df)
    st = cars_obj.state
    assert np.allclose(st["x"].to_numpy(float), g["x"][59], rtol=1e-12, atol=0)
    assert np.array_equal(start_len - np.array([len(p) for p in st["xpath"]]), g["cursor"][59])


@pytest.mark.gpu
def test_write_state_parquet_roundtrip(tmp_path):
    """Cars(serialize=True) dumps the state every update like cars.py:43-58 (pyarrow instead of the absent fastparquet)."""
    import pandas as pd
    from traffic_b200.facade import Cars, TrafficLights
    g = util.load_golden("c1_dt0.001_p2_s1")
    cars_df, lights_df, graph = util.frames_from_golden(g)
    cars_obj = Cars(cars_df, graph, serialize=True)
    cars_obj.serialize_path = str(tmp_path / "data_store")
    lights_obj = TrafficLights(lights_df, graph)
    for _ in range(3):
        lights_obj.update(1e-3)
        cars_obj.update(1e-3, lights_obj.state)
    files = sorted((tmp_path / "data_store").rglob("*.parquet.gz"))
    assert len(files) == 3
    back = pd.read_parquet(files[-1])
    assert np.allclose(back["x"].to_numpy(float), g["x"][2], rtol=1e-12, atol=0) and len(back) == len(cars_df)
