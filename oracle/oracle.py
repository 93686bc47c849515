"""TEST INFRASTRUCTURE - ctypes front end of oracle/traffic_oracle.c (the fp64 CPU restatement of the reference's
``Cars.update`` + ``TrafficLights.update``, ``cars.py:61-133``).  Operates in place on a ``traffic_b200.state.World``.
Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORDER_CARS_LIGHTS = 0  # test/profile_cars_update.py:28-34
ORDER_LIGHTS_CARS = 1  # animate.py:63-64, environment.py:167-168


class ToParams(C.Structure):
    _fields_ = [("n_cars", C.c_int32), ("n_lights", C.c_int32), ("n_faces", C.c_int32),
                ("nbx", C.c_int32), ("nby", C.c_int32),
                ("ex0", C.c_double), ("ex1", C.c_double), ("exd", C.c_double),
                ("ey0", C.c_double), ("ey1", C.c_double), ("eyd", C.c_double),
                ("speed_limit", C.c_double), ("stop_distance", C.c_double), ("free_distance", C.c_double),
                ("default_acceleration", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtraffic_oracle.so")
    src = os.path.join(_HERE, "traffic_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libtraffic_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.to_abi_version.restype = C.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
    return a.ctypes.data_as(C.c_void_p)


def params_of(w) -> ToParams:
    g = w.grid
    p = w.params
    return ToParams(w.n_cars, w.n_lights, w.n_faces, g.nbx, g.nby, g.ex[0], g.ex[1], g.ex[2], g.ey[0], g.ey[1], g.ey[2],
                    p["speed_limit"], p["stop_distance"], p["free_distance"], p["default_acceleration"])


def run(w, dt: float, n_steps: int, order: int = ORDER_CARS_LIGHTS, threads: int | None = None) -> None:
    """Advance ``w`` in place by ``n_steps`` steps."""
    L = lib()
    L.to_set_threads(C.c_int(1 if threads is None else int(threads)))
    P = params_of(w)
    t = C.c_double(w.lights_time)
    L.to_run(C.byref(P), C.c_double(dt), C.c_int(n_steps), C.c_int(order), C.byref(t),
             _p(w.pos), _p(w.vel), _p(w.route_time), _p(w.cursor), _p(w.path_end), _p(w.path_eff_end),
             _p(w.path_xy), _p(w.dest_xy), _p(w.d_node), _p(w.d_car), _p(w.d_light), _p(w.cell), _p(w.leader),
             _p(w.light_xy), _p(w.switch_time), _p(w.face_ptr), _p(w.face_vec), _p(w.face_go),
             _p(w.cell_light_ptr), _p(w.cell_light_idx))
    w.lights_time = t.value
    # Cars.time_elapsed is a Python float running sum (cars.py:75)
    for _ in range(n_steps):
        w.cars_time += dt


def cars_update(w, dt: float, threads: int | None = None) -> None:
    """``Cars.update(dt, lights)`` alone (cars.py:61-95), lights untouched."""
    L = lib()
    L.to_set_threads(C.c_int(1 if threads is None else int(threads)))
    P = params_of(w)
    pos_new = np.empty((max(w.n_cars, 1), 2), np.float64)
    min12 = np.empty(2 * w.grid.n_cells, np.int32)
    L.to_cars_update(C.byref(P), C.c_double(dt), _p(w.pos), _p(w.vel), _p(w.route_time), _p(w.cursor), _p(w.path_end),
                     _p(w.path_eff_end), _p(w.path_xy), _p(w.dest_xy), _p(w.d_node), _p(w.d_car), _p(w.d_light), _p(w.cell),
                     _p(w.leader), _p(w.light_xy), _p(w.face_ptr), _p(w.face_vec), _p(w.face_go), _p(w.cell_light_ptr),
                     _p(w.cell_light_idx), _p(pos_new), _p(min12))
    w.cars_time += dt


def lights_update(w, dt: float) -> None:
    """``TrafficLights.update(dt)`` alone (cars.py:123-133)."""
    L = lib()
    t = C.c_double(w.lights_time)
    L.to_lights_update(C.c_int(w.n_lights), C.byref(t), C.c_double(dt), _p(w.switch_time), _p(w.face_ptr), _p(w.face_go))
    w.lights_time = t.value
