"""TEST INFRASTRUCTURE - a ``native.DeviceSim`` look-alike backed by the CPU oracle, so that the drop-in ``Cars`` /
``TrafficLights`` classes can be driven by the reference's own, unmodified callers in a container that has the
reference tree but no GPU (tests/test_reference_callers.py).  Never imported by the product."""
from __future__ import annotations

import numpy as np

from oracle import oracle
from traffic_b200 import native


class OracleDeviceSim:
    precision = native.F64
    n_replicas = 1

    def __init__(self, world, precision=native.F64, write_aux=True, **_):
        self.w = world.copy()
        self.origin = np.array([world.grid.axis[0], world.grid.axis[2]], np.float64)
        self.n_cars, self.n_lights, self.n_faces = world.n_cars, world.n_lights, world.n_faces
        self.launch_count = 0
        self.closed = False
        self._refs = 0
        self.downloads = []          # (field, n_elements) of every device->host copy, for the tests

    def retain(self):
        self._refs += 1
        return self

    def release(self):
        self._refs -= 1
        if self._refs <= 0:
            self.close()

    def close(self):
        self.closed = True

    def _field(self, field):
        w = self.w
        return {native.POS: w.pos, native.VEL: w.vel, native.ROUTE_TIME: w.route_time, native.CURSOR: w.cursor,
                native.D_NODE: w.d_node, native.D_CAR: w.d_car, native.D_LIGHT: w.d_light, native.CELL: w.cell,
                native.LEADER: w.leader, native.FACE_GO: w.face_go,
                native.LIGHTS_TIME: np.array([w.lights_time])}[field]

    def upload(self, field, arr):
        assert not self.closed
        if field == native.FACE_GO:
            self.w.face_go[:] = np.asarray(arr, np.uint8)
        else:
            raise NotImplementedError(field)

    def download(self, field):
        assert not self.closed, "simulation used after close()"
        a = np.array(self._field(field)).reshape(-1)
        self.downloads.append((field, a.size))
        return a.copy()

    def download_range(self, field, first, count):
        assert not self.closed
        self.downloads.append((field, count))
        return np.array(self._field(field)).reshape(-1)[first:first + count].copy()

    def cars_update(self, dt):
        assert not self.closed
        oracle.cars_update(self.w, dt)
        self.launch_count += 1

    def lights_update(self, dt):
        assert not self.closed
        oracle.lights_update(self.w, dt)
        self.launch_count += 1

    def step(self, dt, n_steps=1, order=0):
        oracle.run(self.w, dt, n_steps, order)

    def synchronize(self):
        pass
