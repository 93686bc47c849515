"""animate.py-side frame consumer: FrameStream / FastAnimator against what the reference's
Animator.animate pattern (animate.py:56-98: lights.update, cars.update, iterrows over cars, per-light lists) extracts from
the drop-in classes, and against the golden record of the reference itself."""
import types

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

NAME = "c1_dt0.001_p2_s1"      # recorded in the Animator's order: lights then cars, dt = 1e-3


def _pair(g):
    from traffic_b200.facade import Cars, TrafficLights
    cars_df, lights_df, graph = util.frames_from_golden(g)
    return Cars(init_state=cars_df, graph=graph), TrafficLights(light_state=lights_df, graph=graph)


def _animator_pattern(cars_object, lights_object, dt):
    """What animate.py:63-86 reads after one frame (one step)."""
    lights_object.update(dt)
    cars_object.update(dt, lights_object.state)
    xy = np.array([[row["x"], row["y"]] for _, row in cars_object.state.iterrows()])
    colors, face_xy = [], []
    for _, light in lights_object.state.iterrows():
        face_xy += list(zip(light["out-xpositions"], light["out-ypositions"]))
        colors += [bool(c) for c in light["go-values"]]
    return xy, np.array(colors), np.array(face_xy)


@pytest.mark.parametrize("k", [1, 7])
def test_frame_stream_equals_animator_pattern_and_golden(k):
    from traffic_b200.frames import FrameStream
    g = util.load_golden(NAME)
    dt = float(g["dt"])
    a_cars, a_lights = _pair(g)
    b_cars, b_lights = _pair(g)
    stream = FrameStream(b_cars, b_lights, dt, k=k)
    for frame in range(12):
        for _ in range(k):
            xy_ref, colors_ref, face_xy_ref = _animator_pattern(a_cars, a_lights, dt)
        xy, go = stream.next()
        if frame < 11:
            stream.prefetch()                  # the next frame is computed while this one is looked at
        assert np.array_equal(xy, xy_ref), "frame %d" % frame
        assert np.array_equal(go.astype(bool), colors_ref)
        assert np.allclose(stream.face_xy, face_xy_ref, rtol=0, atol=1e-9)
        t = (frame + 1) * k - 1
        assert np.allclose(xy[:, 0], g["x"][t], rtol=1e-9, atol=0) and np.allclose(xy[:, 1], g["y"][t], rtol=1e-9, atol=0)
        assert np.array_equal(go, g["face_go"][t])
    # the facade objects stay usable and consistent after streaming
    assert abs(b_cars.time_elapsed - a_cars.time_elapsed) < 1e-12
    assert np.array_equal(b_cars.state["x"].to_numpy(), a_cars.state["x"].to_numpy())


class _Scatter:
    def __init__(self):
        self.offsets, self.colors = None, None

    def set_offsets(self, xy):
        self.offsets = np.array(xy, float)

    def set_color(self, c):
        self.colors = np.array(c)


def test_fast_animator_draws_what_the_reference_animator_draws():
    from traffic_b200.frames import FastAnimator
    g = util.load_golden(NAME)
    dt = float(g["dt"])
    a_cars, a_lights = _pair(g)
    b_cars, b_lights = _pair(g)
    ax = types.SimpleNamespace(scatter=lambda *a, **k: _Scatter(), axis=lambda: (0, 1, 0, 1), annotate=lambda *a, **k: None)
    fig = types.SimpleNamespace(canvas=types.SimpleNamespace(draw=lambda: None))
    n = len(g["pos0"])
    anim = FastAnimator(fig, ax, b_cars, b_lights, num=(1, 1), dt=dt, n=n)
    assert anim.number_of_lights == len(g["switch_time"]) and anim.number_of_faces == len(g["face_go0"])
    anim.reset((1, 3))
    for i in range(25):
        xy_ref, colors_ref, face_xy_ref = _animator_pattern(a_cars, a_lights, dt)
        cars_art, lights_art, faces_art = anim.animate(i)
        assert np.array_equal(cars_art.offsets, xy_ref)
        assert np.array_equal(faces_art.colors == "green", colors_ref)
        assert np.allclose(faces_art.offsets, face_xy_ref, rtol=0, atol=1e-9)
        assert np.allclose(lights_art.offsets, g["light_xy"], rtol=0, atol=0)
