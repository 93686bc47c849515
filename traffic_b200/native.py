"""ctypes binding of ``libtraffic_b200.so`` (C ABI in ``include/traffic_b200.h``) + ``DeviceSim``, the device-resident
simulation object the ``Cars`` / ``TrafficLights`` facade drives.

PyTorch is plumbing here: it owns the device arena (one ``torch.uint8`` tensor from the caching allocator, so
re-creating a simulation per episode - ``environment.py:50-51,85`` - costs no cudaMalloc), provides the stream and
zero-copy tensor views of the fields.  All compute is in the hand-written kernels behind the C ABI.  There is no CPU
fallback: if the library or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .state import World

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.environ.get("TB2_LIB", os.path.join(_CSRC, "libtraffic_b200.so"))  # TB2_LIB: kernel-variant experiments

ABI_VERSION = 2
ORDER_CARS_LIGHTS = 0  # test/profile_cars_update.py:29-34
ORDER_LIGHTS_CARS = 1  # animate.py:63-64, environment.py:167-168
F64, F32 = 0, 1
NEIGHBOR_ATOMIC, NEIGHBOR_SORT = 0, 1

# tb2_field
(POS, VEL, ROUTE_TIME, CURSOR, PATH_END, PATH_EFF_END, PATH_XY, DEST_XY, D_NODE, D_CAR, D_LIGHT, CELL, LEADER,
 LIGHT_XY, SWITCH_TIME, FACE_PTR, FACE_VEC, FACE_GO, CELL_LIGHT_PTR, CELL_LIGHT_IDX, LIGHTS_TIME,
 DONE, TRIP_TIME, TRIP_STEP) = range(24)
FIELD_COUNT = 24

EXPORTS = ["tb2_abi_version", "tb2_last_error", "tb2_arena_bytes", "tb2_create", "tb2_destroy", "tb2_field_info",
           "tb2_field_ptr", "tb2_upload", "tb2_download", "tb2_prepare", "tb2_step", "tb2_step_host", "tb2_cars_update", "tb2_lights_update",
           "tb2_launch_count", "tb2_step_count", "tb2_route_paths", "tb2_route_last_error", "tb2_step_host_async",
           "tb2_frame_wait", "tb2_rollout", "tb2_download_range", "tb2_debug_bounds_enabled", "tb2_debug_bounds_violations", "tb2_route_release",
           "tb2_route_polylines"]


class Tb2Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("precision", C.c_int32),
                ("n_cars", C.c_int32), ("n_lights", C.c_int32), ("n_faces", C.c_int32), ("n_cell_lights", C.c_int32),
                ("n_path_points", C.c_int64),
                ("nbx", C.c_int32), ("nby", C.c_int32),
                ("ex0", C.c_double), ("ex1", C.c_double), ("exd", C.c_double),
                ("ey0", C.c_double), ("ey1", C.c_double), ("eyd", C.c_double),
                ("speed_limit", C.c_double), ("stop_distance", C.c_double), ("free_distance", C.c_double),
                ("default_acceleration", C.c_double),
                ("origin_x", C.c_double), ("origin_y", C.c_double),
                ("write_aux", C.c_int32), ("neighbor_mode", C.c_int32),
                ("n_replicas", C.c_int32), ("agent_car", C.c_int32)]


class NativeError(RuntimeError):
    pass


_LIB = None


def build(force: bool = False) -> str:
    """Compile the C-ABI library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _CSRC, "-s"] + (["-B"] if force else []), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise NativeError("building libtraffic_b200.so failed:\n" + r.stdout[-4000:])
    return LIB_PATH


def lib():
    """Load the library; fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError("%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.tb2_abi_version.restype = C.c_int
        L.tb2_last_error.restype = C.c_char_p
        L.tb2_arena_bytes.restype = C.c_size_t
        L.tb2_arena_bytes.argtypes = [C.POINTER(Tb2Config)]
        L.tb2_create.argtypes = [C.POINTER(Tb2Config), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.tb2_destroy.argtypes = [C.c_void_p]
        L.tb2_field_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.tb2_field_ptr.restype = C.c_void_p
        L.tb2_field_ptr.argtypes = [C.c_void_p, C.c_int]
        L.tb2_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.tb2_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.tb2_download_range.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.tb2_debug_bounds_violations.restype = C.c_uint
        L.tb2_prepare.argtypes = [C.c_void_p, C.c_void_p]
        L.tb2_step.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p]
        L.tb2_rollout.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p]
        L.tb2_step_host.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
        L.tb2_step_host_async.argtypes = L.tb2_step_host.argtypes
        L.tb2_frame_wait.argtypes = [C.c_void_p, C.c_int32]
        L.tb2_cars_update.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.tb2_lights_update.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.tb2_route_paths.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tb2_route_polylines.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.POINTER(C.c_double)), C.c_void_p]
        L.tb2_route_last_error.restype = C.c_char_p
        L.tb2_launch_count.restype = C.c_int64
        L.tb2_launch_count.argtypes = [C.c_void_p]
        L.tb2_step_count.restype = C.c_int64
        L.tb2_step_count.argtypes = [C.c_void_p]
        if L.tb2_abi_version() != ABI_VERSION:
            raise NativeError("libtraffic_b200.so ABI %d != binding ABI %d" % (L.tb2_abi_version(), ABI_VERSION))
        _LIB = L
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        raise NativeError("%s failed (status %d): %s" % (what, rc, lib().tb2_last_error().decode()))


def config_of(w: World, precision: int = F64, write_aux: bool = True, n_replicas: int = 1, agent_car: int = -1,
              neighbor_mode: int = 0) -> Tb2Config:
    """``w`` describes ONE replica; per-car arrays of an ensemble are tiled by the caller (``traffic_b200.ensemble``)."""
    g, p = w.grid, w.params
    return Tb2Config(ABI_VERSION, precision, w.n_cars, w.n_lights, w.n_faces, len(w.cell_light_idx), len(w.path_xy),
                     g.nbx, g.nby, g.ex[0], g.ex[1], g.ex[2], g.ey[0], g.ey[1], g.ey[2],
                     p["speed_limit"], p["stop_distance"], p["free_distance"], p["default_acceleration"],
                     g.axis[0], g.axis[2], 1 if write_aux else 0, neighbor_mode, n_replicas, agent_car)


_NP_OF_FIELD = {POS: np.float64, VEL: np.float64, ROUTE_TIME: np.float64, CURSOR: np.int32, PATH_END: np.int32,
                PATH_EFF_END: np.int32, PATH_XY: np.float64, DEST_XY: np.float64, D_NODE: np.float64,
                D_CAR: np.float64, D_LIGHT: np.float64, CELL: np.int32, LEADER: np.int32, LIGHT_XY: np.float64,
                SWITCH_TIME: np.float64, FACE_PTR: np.int32, FACE_VEC: np.float64, FACE_GO: np.uint8,
                CELL_LIGHT_PTR: np.int32, CELL_LIGHT_IDX: np.int32, LIGHTS_TIME: np.float64,
                DONE: np.int32, TRIP_TIME: np.float64, TRIP_STEP: np.int32}
_REAL_FIELDS = (POS, VEL, ROUTE_TIME, PATH_XY, DEST_XY, D_NODE, D_CAR, D_LIGHT, LIGHT_XY, FACE_VEC)


class DeviceSim:
    """One simulation resident on one GPU.  ``step`` is asynchronous on the current torch stream."""

    def __init__(self, world: World, precision: int = F64, write_aux: bool = True, device=None,
                 n_replicas: int = 1, agent_car: int = -1, switch_time=None, face_go=None,
                 neighbor_mode: int = NEIGHBOR_ATOMIC):
        """``world`` describes one replica.  With ``n_replicas`` > 1 the car state is tiled and ``switch_time``
        ``[R, L]`` / ``face_go`` ``[R, F]`` may give every replica its own light timings (default: the world's)."""
        import torch

        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: traffic_b200 has no CPU path")
        self._torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.L = lib()
        self.precision = precision
        self.cfg = config_of(world, precision, write_aux, n_replicas, agent_car, neighbor_mode)
        self.n_replicas = int(n_replicas)
        self._switch_time, self._face_go = switch_time, face_go
        self.n_cars, self.n_lights, self.n_faces = world.n_cars, world.n_lights, world.n_faces
        self.origin = np.array([world.grid.axis[0], world.grid.axis[2]], np.float64)
        nbytes = self.L.tb2_arena_bytes(C.byref(self.cfg))
        if nbytes == 0:
            raise NativeError("tb2_arena_bytes: %s" % self.L.tb2_last_error().decode())
        with torch.cuda.device(self.device):
            self.arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = self.arena.data_ptr()
            self._arena_base = (base + 255) // 256 * 256
            h = C.c_void_p()
            _check(self.L.tb2_create(C.byref(self.cfg), C.c_void_p(self._arena_base), C.c_size_t(nbytes), C.byref(h)),
                   "tb2_create")
        self.h = h
        self._refs = 0   # facade objects sharing this simulation (retain / release)
        self._keep = []  # host staging arrays that must outlive async copies
        self.upload_world(world)

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _real(self):
        return np.float64 if self.precision == F64 else np.float32

    def _dtype(self, field):
        return self._real() if field in _REAL_FIELDS else _NP_OF_FIELD[field]

    def retain(self):
        self._refs += 1
        return self

    def release(self):
        """Drop one holder; the device simulation is destroyed when the last holder lets go."""
        self._refs -= 1
        if self._refs <= 0:
            self.close()

    def upload(self, field: int, arr: np.ndarray):
        a = np.ascontiguousarray(arr, self._dtype(field))
        if len(self._keep) >= 32:   # a driver loop that never reads state back must not pile up staging arrays
            self.synchronize()
        self._keep.append(a)
        _check(self.L.tb2_upload(self.h, field, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), self._stream()),
               "tb2_upload(%d)" % field)

    def download(self, field: int) -> np.ndarray:
        n, eb = C.c_size_t(), C.c_size_t()
        _check(self.L.tb2_field_info(self.h, field, C.byref(n), C.byref(eb)), "tb2_field_info")
        out = np.empty(n.value, self._dtype(field))
        assert out.itemsize == eb.value
        _check(self.L.tb2_download(self.h, field, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes), self._stream()),
               "tb2_download(%d)" % field)
        self.synchronize()
        return out

    def download_range(self, field: int, first: int, count: int) -> np.ndarray:
        """Elements [first, first + count) of a field (one small device->host copy)."""
        out = np.empty(count, self._dtype(field))
        _check(self.L.tb2_download_range(self.h, field, C.c_size_t(first), C.c_size_t(count),
                                         out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes), self._stream()),
               "tb2_download_range(%d)" % field)
        self.synchronize()
        return out

    def synchronize(self):
        self._torch.cuda.current_stream(self.device).synchronize()
        self._keep.clear()

    def tensor(self, field: int):
        """Zero-copy torch view of a device field (valid until the next step for POS / FACE_GO / CELL)."""
        torch = self._torch
        n, eb = C.c_size_t(), C.c_size_t()
        _check(self.L.tb2_field_info(self.h, field, C.byref(n), C.byref(eb)), "tb2_field_info")
        ptr = self.L.tb2_field_ptr(self.h, field)
        off = ptr - self.arena.data_ptr()
        tdt = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8}[self._dtype(field)]
        return self.arena[off: off + n.value * eb.value].view(tdt)

    # -- state ------------------------------------------------------------------------------------------------
    def _rel(self, xy):
        return xy if self.precision == F64 else xy - self.origin

    def upload_world(self, w: World):
        R = self.n_replicas

        def tile(a):  # replicas back to back
            return a if R == 1 else np.tile(a, (R,) + (1,) * (a.ndim - 1))

        with self._torch.cuda.device(self.device):
            self.upload(POS, tile(self._rel(w.pos)))
            self.upload(VEL, tile(w.vel))
            self.upload(ROUTE_TIME, tile(w.route_time))
            self.upload(CURSOR, tile(w.cursor))
            self.upload(PATH_END, tile(w.path_end))
            self.upload(PATH_EFF_END, tile(w.path_eff_end))
            self.upload(PATH_XY, self._rel(w.path_xy))
            self.upload(DEST_XY, tile(self._rel(w.dest_xy)))
            self.upload(D_NODE, tile(w.d_node))
            self.upload(D_CAR, tile(w.d_car))
            self.upload(D_LIGHT, tile(w.d_light))
            self.upload(LEADER, tile(w.leader))
            self.upload(LIGHT_XY, self._rel(w.light_xy))
            self.upload(SWITCH_TIME, tile(w.switch_time) if self._switch_time is None else np.asarray(self._switch_time, np.float64).reshape(-1))
            self.upload(FACE_PTR, w.face_ptr)
            self.upload(FACE_VEC, w.face_vec)
            self.upload(FACE_GO, tile(w.face_go) if self._face_go is None else np.asarray(self._face_go, np.uint8).reshape(-1))
            self.upload(CELL_LIGHT_PTR, w.cell_light_ptr)
            self.upload(CELL_LIGHT_IDX, w.cell_light_idx)
            self.upload(LIGHTS_TIME, np.array([w.lights_time], np.float64))
            _check(self.L.tb2_prepare(self.h, self._stream()), "tb2_prepare")
            self.synchronize()

    def step(self, dt: float, n_steps: int = 1, order: int = ORDER_CARS_LIGHTS):
        with self._torch.cuda.device(self.device):
            _check(self.L.tb2_step(self.h, C.c_double(dt), C.c_int32(n_steps), C.c_int32(order), self._stream()), "tb2_step")

    def rollout(self, dt: float, max_steps: int, order: int = ORDER_LIGHTS_CARS):
        """``Env.step``'s rollout loop for every replica in one launch (``tb2_rollout``)."""
        with self._torch.cuda.device(self.device):
            _check(self.L.tb2_rollout(self.h, C.c_double(dt), C.c_int32(max_steps), C.c_int32(order), self._stream()),
                   "tb2_rollout")

    def cars_update(self, dt: float):
        with self._torch.cuda.device(self.device):
            _check(self.L.tb2_cars_update(self.h, C.c_double(dt), self._stream()), "tb2_cars_update")

    def lights_update(self, dt: float):
        with self._torch.cuda.device(self.device):
            _check(self.L.tb2_lights_update(self.h, C.c_double(dt), self._stream()), "tb2_lights_update")

    def step_host(self, dt, n_steps, order, face_go_in=None, frame_out=None, face_go_out=None):
        """End-to-end call with host buffers (pinned numpy/torch memory recommended)."""
        def p(a):
            return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())
        with self._torch.cuda.device(self.device):
            _check(self.L.tb2_step_host(self.h, C.c_double(dt), C.c_int32(n_steps), C.c_int32(order), p(face_go_in),
                                        p(frame_out), p(face_go_out), self._stream()), "tb2_step_host")

    def step_host_async(self, dt, n_steps, order, face_go_in=None, frame_out=None, face_go_out=None):
        """Pipelined ``step_host``: returns once enqueued; ``frame_wait()`` blocks until the last frame has landed."""
        def p(a):
            return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())
        with self._torch.cuda.device(self.device):
            _check(self.L.tb2_step_host_async(self.h, C.c_double(dt), C.c_int32(n_steps), C.c_int32(order), p(face_go_in),
                                              p(frame_out), p(face_go_out), self._stream()), "tb2_step_host_async")

    def frame_wait(self, faces_only: bool = False):
        _check(self.L.tb2_frame_wait(self.h, 1 if faces_only else 0), "tb2_frame_wait")

    def download_world(self, w: World, replica: int = 0) -> World:
        """Refresh the mutable fields of ``w`` from the device (of one replica when this is an ensemble)."""
        if self.n_replicas > 1:
            return self._download_replica(w, replica)
        pos = self.download(POS).reshape(-1, 2).astype(np.float64)
        w.pos = pos if self.precision == F64 else pos + self.origin
        w.vel = self.download(VEL).reshape(-1, 2).astype(np.float64)
        w.route_time = self.download(ROUTE_TIME).astype(np.float64)
        w.cursor = self.download(CURSOR)
        w.d_node = self.download(D_NODE).astype(np.float64)
        w.d_car = self.download(D_CAR).astype(np.float64)
        w.d_light = self.download(D_LIGHT).astype(np.float64)
        w.cell = self.download(CELL)
        w.leader = self.download(LEADER)
        w.face_go = self.download(FACE_GO)
        w.lights_time = float(self.download(LIGHTS_TIME)[0])
        return w

    def _download_replica(self, w: World, r: int) -> World:
        N, F = self.n_cars, self.n_faces
        cs, fs = slice(r * N, (r + 1) * N), slice(r * F, (r + 1) * F)
        pos = self.download(POS).reshape(-1, 2)[cs].astype(np.float64)
        w.pos = pos if self.precision == F64 else pos + self.origin
        w.vel = self.download(VEL).reshape(-1, 2)[cs].astype(np.float64)
        w.route_time = self.download(ROUTE_TIME)[cs].astype(np.float64)
        w.cursor = self.download(CURSOR)[cs].copy()
        w.d_node = self.download(D_NODE)[cs].astype(np.float64)
        w.d_car = self.download(D_CAR)[cs].astype(np.float64)
        w.d_light = self.download(D_LIGHT)[cs].astype(np.float64)
        w.cell = self.download(CELL)[cs].copy()
        w.leader = self.download(LEADER)[cs].copy()
        w.face_go = self.download(FACE_GO)[fs].copy()
        w.lights_time = float(self.download(LIGHTS_TIME)[0])
        return w

    @property
    def launch_count(self) -> int:
        return int(self.L.tb2_launch_count(self.h))

    @property
    def step_count(self) -> int:
        return int(self.L.tb2_step_count(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.tb2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
