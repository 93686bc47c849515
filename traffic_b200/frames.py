"""Frame extraction for ``animate.py``-style consumers.

The reference's ``Animator.animate`` (``animate.py:56-98``) steps the simulation once per frame and then walks the cars
DataFrame row by row (``iterrows``) and every light's Python lists to move one matplotlib artist per car, per light and
per light face.  Here

* ``FrameStream`` advances the device simulation ``k`` steps per frame through the pipelined host-buffer entry point
  (``tb2_step_host_async``: the frame is snapshotted on the device and travels to pinned host memory on a private copy
  stream while the next frame is already being computed) and hands out, per frame, the car positions ``[N, 2]`` and the
  per-face go flags ``[F]`` as flat arrays - what a renderer needs, nothing else crosses PCIe;
* ``FastAnimator`` keeps the reference's ``Animator`` interface (constructor arguments, ``reset``, ``animate(i)``) but
  draws with THREE collections (cars, lights, faces: ``ax.scatter`` + ``set_offsets`` / ``set_color``) instead of
  N + L + F line artists.
"""
from __future__ import annotations

import numpy as np

from . import native
from .facade import Cars, TrafficLights


class FrameStream:
    """Iterator over frames of a device-resident simulation.  ``next()`` returns ``(xy, go)``: views of pinned host
    buffers holding the positions and face states after the next ``k`` steps (valid until the call after next)."""

    def __init__(self, cars: Cars, lights: TrafficLights, dt: float, k: int = 1, order: int = native.ORDER_LIGHTS_CARS):
        import torch

        self.cars, self.lights, self.dt, self.k, self.order = cars, lights, float(dt), int(k), order
        cars.lights = lights.state
        cars._bind(lights.state)                      # one combined device simulation (as the first update() would)
        self.sim = cars._sim
        n, f = self.sim.n_cars, self.sim.n_faces
        real = torch.float64 if self.sim.precision == native.F64 else torch.float32
        self._xy = [torch.empty((n, 2), dtype=real).pin_memory() for _ in range(2)]
        self._go = [torch.empty((f,), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._slot = 0
        self._pending = False
        self.face_xy = self._face_positions()

    def _face_positions(self):
        """Static: where the reference draws every light face ('out-xpositions' / 'out-ypositions', animate.py:74-81)."""
        st = self.lights.init_state
        xs = [x for row in st["out-xpositions"] for x in row]
        ys = [y for row in st["out-ypositions"] for y in row]
        return np.stack([np.asarray(xs, np.float64), np.asarray(ys, np.float64)], 1) if xs else np.zeros((0, 2))

    def _enqueue(self):
        s = self._slot
        self.sim.step_host_async(self.dt, self.k, self.order, None, self._xy[s], self._go[s])
        for _ in range(self.k):                        # the facade objects' clocks (Python float running sums)
            self.cars.time_elapsed += self.dt
            self.lights.time_elapsed += self.dt
        self.cars._updates += self.k
        self.cars._stale = set(_ALL_GROUPS)
        self.lights._dirty = True
        self._pending = True

    def next(self):
        if not self._pending:
            self._enqueue()
        self.sim.frame_wait()                          # this frame has landed ...
        s = self._slot
        self._slot ^= 1
        self._pending = False
        xy = self._xy[s].numpy()
        if self.sim.precision != native.F64:
            xy = xy.astype(np.float64) + self.sim.origin
        return xy, self._go[s].numpy()

    def prefetch(self):
        """Start computing the next frame now (overlaps with whatever the caller does with the current one)."""
        if not self._pending:
            self._enqueue()

    def __iter__(self):
        return self

    def __next__(self):
        return self.next()


def _all_groups():
    from .facade import _GROUPS
    return _GROUPS


_ALL_GROUPS = _all_groups()


class FastAnimator:
    """``animate.Animator`` with the same constructor and ``reset`` / ``animate(i)`` protocol (``animate.py:4-98``) on
    top of ``FrameStream``: one ``set_offsets`` for all cars, one ``set_color`` for all faces per frame."""

    def __init__(self, fig, ax, cars_object, lights_object, num, frame_rate=1000, dt=1 / 1000, n=1, focus=None,
                 write_frames=False, steps_per_frame=1):
        self.fig, self.ax, self.num = fig, ax, num
        self.frame_rate, self.dt, self.N, self.focus, self.write_frames = frame_rate, dt, n, focus, write_frames
        self.cars_object, self.lights_object = cars_object, lights_object
        self.number_of_lights = len(lights_object.state)
        self.number_of_faces = sum(lights_object.state["degree"])
        self.stream = FrameStream(cars_object, lights_object, dt, steps_per_frame)
        self.cars = ax.scatter([], [], color="blue", marker="o", s=1)
        self.lights = ax.scatter([], [], color="red", marker="+", s=4)
        self.faces = ax.scatter([], [], color="red", marker="^", s=4)
        st = lights_object.init_state
        self._light_xy = np.stack([st["x"].to_numpy(np.float64), st["y"].to_numpy(np.float64)], 1) if len(st) else np.zeros((0, 2))

    def reset(self, num=None):
        empty = np.zeros((0, 2))
        for artist in (self.cars, self.lights, self.faces):
            artist.set_offsets(empty)
        self.num = num if num else self.num
        axis = self.ax.axis()
        self.ax.annotate("Episode {} of {}".format(self.num[0], self.num[1]), xy=(axis[0] + 10, axis[2] + 10))
        self.fig.canvas.draw()
        return [self.cars, self.lights, self.faces]

    def animate(self, i):
        xy, go = self.stream.next()
        self.stream.prefetch()                         # frame i+1 is computed while frame i is drawn
        self.cars.set_offsets(xy[: self.N])            # the reference draws the first n cars (animate.py:19, 66)
        self.lights.set_offsets(self._light_xy)
        self.faces.set_offsets(self.stream.face_xy)
        self.faces.set_color(np.where(go.astype(bool), "green", "red"))
        self.fig.canvas.draw()
        return [self.cars, self.lights, self.faces]
