"""Host-side structure-of-arrays image of the simulation state + the DataFrame <-> SoA packing boundary.

The reference keeps cars and lights as pandas DataFrames whose ``xpath``/``ypath``/``go-values`` cells hold Python
lists (schemas: ``simulation.py:226-252`` for cars, ``simulation.py:358-378`` for lights).
``World`` is the flat image of the same information that the device kernels (and the CPU oracle) consume:

* per-car routes become one CSR of polyline points (``path_xy[P,2]``, ``path_start/path_end[N]``) plus a
  ``cursor`` - popping the head of the Python list (``simulation.py:47-52``) is ``cursor += 1``;
* the destination-node fallback (``navigation.py:84,89,566-578``) becomes ``dest_xy[N,2]``;
* light faces become a CSR (``face_ptr[L+1]``, ``face_vec[F,2]``, ``face_go[F]``);
* the "same 200 m bin" masks (``navigation.py:438-441, 474-476``) become a static cell -> lights CSR in
  lights-DataFrame order and a per-step cell -> two-smallest-car-index table built on the device.

Numpy only; no torch here so the oracle-side tests can use the packer without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .synth import BIN_SIZE

# reference literals (SURVEY.md §5 "Config"); simulation.py:13-16, cars.py:41
SPEED_LIMIT = 1500.0
STOP_DISTANCE = 20.0
FREE_DISTANCE = 60.0
DEFAULT_ACCELERATION = 4.0


@dataclass
class Grid:
    """Bin edges of ``models.determine_bins`` (``models.py:7-19``) in a closed form the kernels can evaluate.

    ``np.arange(a0, a1, 200)`` fills ``e[0] = a0``, ``e[1] = a0 + 200`` and ``e[k] = a0 + k * (e[1] - e[0])`` for
    k >= 2 (numpy's ``DOUBLE_fill``), so ``delta`` may differ from 200 by an ulp of ``a0``; we carry
    ``e0, e1, delta`` per axis and verify the closed form against ``np.arange`` itself at pack time.
    """
    axis: tuple
    nbx: int
    nby: int
    ex: tuple  # (e0, e1, delta)
    ey: tuple

    @property
    def ncx(self) -> int:
        return self.nbx + 1

    @property
    def ncy(self) -> int:
        return self.nby + 1

    @property
    def n_cells(self) -> int:
        return self.ncx * self.ncy

    @staticmethod
    def from_axis(axis, bin_size: float = BIN_SIZE) -> "Grid":
        out = []
        for lo, hi in ((axis[0], axis[1]), (axis[2], axis[3])):
            e = np.arange(lo, hi, bin_size)
            nb = len(e)
            e0 = float(e[0]) if nb > 0 else float(lo)
            e1 = float(e[1]) if nb > 1 else float(np.float64(lo) + np.float64(bin_size))
            delta = float(np.float64(e1) - np.float64(e0))
            k = np.arange(nb, dtype=np.float64)
            closed = np.float64(e0) + k * np.float64(delta)
            if nb > 0:
                closed[0] = e0
            if nb > 1:
                closed[1] = e1
            if not np.array_equal(closed, e):
                raise ValueError("bin-edge closed form does not reproduce np.arange for axis %r" % ((lo, hi),))
            out.append((nb, (e0, e1, delta)))
        return Grid(tuple(float(a) for a in axis), out[0][0], out[1][0], out[0][1], out[1][1])

    def edges(self):
        return (np.arange(self.axis[0], self.axis[1], BIN_SIZE), np.arange(self.axis[2], self.axis[3], BIN_SIZE))


@dataclass
class World:
    grid: Grid
    # ---- cars (N) ----
    pos: np.ndarray            # (N,2) f64
    vel: np.ndarray            # (N,2) f64
    route_time: np.ndarray     # (N,)  f64
    cursor: np.ndarray         # (N,)  i32 absolute index into path_xy of the current path head
    path_start: np.ndarray     # (N,)  i32
    path_end: np.ndarray       # (N,)  i32 one past the last point of the car's polyline
    path_eff_end: np.ndarray   # (N,)  i32 has_path <=> cursor < path_eff_end  (navigation.py:37-39 `.any()` rule)
    path_xy: np.ndarray        # (P,2) f64
    dest_xy: np.ndarray        # (N,2) f64
    d_node: np.ndarray         # (N,)  f64
    d_car: np.ndarray          # (N,)  f64   0.0 encodes the reference's `False`
    d_light: np.ndarray        # (N,)  f64   0.0 encodes `False`, -0.0 encodes `None` (both falsy, Appendix B-4)
    cell: np.ndarray           # (N,)  i32   ybin * ncx + xbin of the start-of-step position
    leader: np.ndarray         # (N,)  i32   index of the car accepted as obstacle, -1 if none
    # ---- lights (L), faces (F) ----
    light_xy: np.ndarray       # (L,2) f64
    switch_time: np.ndarray    # (L,)  f64
    face_ptr: np.ndarray       # (L+1,) i32
    face_vec: np.ndarray       # (F,2) f64
    face_go: np.ndarray        # (F,)  u8
    cell_light_ptr: np.ndarray  # (C+1,) i32
    cell_light_idx: np.ndarray  # (.,)   i32 lights of a cell in lights-DataFrame order
    lights_time: float = 0.0   # TrafficLights.time_elapsed (cars.py:117,130)
    cars_time: float = 0.0     # Cars.time_elapsed (cars.py:36,75)
    params: dict = field(default_factory=lambda: dict(speed_limit=SPEED_LIMIT, stop_distance=STOP_DISTANCE,
                                                      free_distance=FREE_DISTANCE,
                                                      default_acceleration=DEFAULT_ACCELERATION))

    @property
    def n_cars(self) -> int:
        return len(self.route_time)

    @property
    def n_lights(self) -> int:
        return len(self.switch_time)

    @property
    def n_faces(self) -> int:
        return len(self.face_go)

    def copy(self) -> "World":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.copy() if isinstance(v, np.ndarray) else (dict(v) if isinstance(v, dict) else v)
        return World(**kw)


def digitize_cells(grid: Grid, pos: np.ndarray) -> np.ndarray:
    """cell id per position with numpy's own digitize (host-side helper, used for lights and for tests)."""
    ex, ey = grid.edges()
    xb = np.digitize(pos[:, 0], ex)
    yb = np.digitize(pos[:, 1], ey)
    return (yb * grid.ncx + xb).astype(np.int32)


def _effective_end(path_xy: np.ndarray, start: np.ndarray, end: np.ndarray) -> np.ndarray:
    """``has_path`` in the reference is ``np.array(xpath).any() and np.array(ypath).any()`` on the *remaining*
    list (``navigation.py:37-39``, ``simulation.py:38-39``): true while a non-zero x and a non-zero y are both still
    ahead of the cursor.  eff_end = 1 + min(last index with x != 0, last index with y != 0)."""
    n = len(start)
    eff = np.array(start, dtype=np.int32, copy=True)
    if len(path_xy) == 0:
        return eff
    nzx = ~(path_xy[:, 0] == 0)
    nzy = ~(path_xy[:, 1] == 0)
    if nzx.all() and nzy.all():
        return np.array(end, dtype=np.int32, copy=True)
    for i in range(n):  # rare slow path (synthetic tests only)
        s, e = int(start[i]), int(end[i])
        lx = np.flatnonzero(nzx[s:e])
        ly = np.flatnonzero(nzy[s:e])
        eff[i] = s + min(lx[-1], ly[-1]) + 1 if len(lx) and len(ly) else s
    return eff


def build_cell_lights(grid: Grid, light_cell: np.ndarray):
    """Static cell -> lights CSR, lights kept in DataFrame order inside a cell (``navigation.py:474-477``)."""
    C = grid.n_cells
    light_cell = np.asarray(light_cell, np.int64)
    ok = (light_cell >= 0) & (light_cell < C)
    idx = np.flatnonzero(ok)
    order = idx[np.argsort(light_cell[idx], kind="stable")]
    counts = np.bincount(light_cell[idx], minlength=C)
    ptr = np.zeros(C + 1, np.int32)
    np.cumsum(counts, out=ptr[1:])
    return ptr, order.astype(np.int32)


def make_world(axis, pos, path_ptr, path_xy, dest_xy, lights: dict | None = None, light_cell=None,
               vel=None, route_time=None, params=None) -> World:
    """Assemble a ``World`` from flat arrays (the large-config generators and the golden fixtures use this)."""
    grid = Grid.from_axis(axis)
    pos = np.ascontiguousarray(pos, np.float64)
    n = len(pos)
    path_ptr = np.asarray(path_ptr, np.int64)
    if path_ptr[-1] >= 2 ** 31:
        raise ValueError("more than 2^31 path points")
    start = path_ptr[:-1].astype(np.int32)
    end = path_ptr[1:].astype(np.int32)
    path_xy = np.ascontiguousarray(path_xy, np.float64).reshape(-1, 2)
    if lights is None:
        lights = dict(xy=np.zeros((0, 2)), switch_time=np.zeros(0), face_ptr=np.zeros(1, np.int32),
                      face_vec=np.zeros((0, 2)), face_go=np.zeros(0, np.uint8))
    light_xy = np.ascontiguousarray(lights["xy"], np.float64).reshape(-1, 2)
    if light_cell is None:
        light_cell = digitize_cells(grid, light_xy) if len(light_xy) else np.zeros(0, np.int32)
    clp, cli = build_cell_lights(grid, light_cell)
    w = World(
        grid=grid, pos=pos,
        vel=np.zeros((n, 2)) if vel is None else np.ascontiguousarray(vel, np.float64),
        route_time=np.zeros(n) if route_time is None else np.ascontiguousarray(route_time, np.float64),
        cursor=start.copy(), path_start=start, path_end=end, path_eff_end=_effective_end(path_xy, start, end),
        path_xy=path_xy, dest_xy=np.ascontiguousarray(dest_xy, np.float64).reshape(-1, 2),
        d_node=np.zeros(n), d_car=np.zeros(n), d_light=np.zeros(n),
        cell=digitize_cells(grid, pos) if n else np.zeros(0, np.int32), leader=np.full(n, -1, np.int32),
        light_xy=light_xy, switch_time=np.ascontiguousarray(lights["switch_time"], np.float64),
        face_ptr=np.ascontiguousarray(lights["face_ptr"], np.int32),
        face_vec=np.ascontiguousarray(lights["face_vec"], np.float64).reshape(-1, 2),
        face_go=np.ascontiguousarray(lights["face_go"]).astype(np.uint8),
        cell_light_ptr=clp, cell_light_idx=cli)
    if params:
        w.params.update(params)
    return w


# ----------------------------------------------------------------------------------------------------------------
# DataFrame boundary
# ----------------------------------------------------------------------------------------------------------------

def lights_from_dataframe(lights_df) -> tuple[dict, np.ndarray]:
    """Flatten the lights DataFrame (``simulation.py:358-378``).  Returns (lights dict, (xbin, ybin) columns)."""
    L = len(lights_df)
    if L == 0:
        return None, (np.zeros(0, np.int64), np.zeros(0, np.int64))
    xy = np.stack([lights_df["x"].to_numpy(np.float64), lights_df["y"].to_numpy(np.float64)], 1)
    degree = np.asarray([int(d) for d in lights_df["degree"]], np.int32)
    face_ptr = np.zeros(L + 1, np.int32)
    np.cumsum(degree, out=face_ptr[1:])
    fx = [np.asarray(v, np.float64)[:d] for v, d in zip(lights_df["out-xvectors"], degree)]
    fy = [np.asarray(v, np.float64)[:d] for v, d in zip(lights_df["out-yvectors"], degree)]
    go = [np.asarray(v).astype(bool)[:d] for v, d in zip(lights_df["go-values"], degree)]
    face_vec = np.stack([np.concatenate(fx), np.concatenate(fy)], 1) if face_ptr[-1] else np.zeros((0, 2))
    face_go = np.concatenate(go).astype(np.uint8) if face_ptr[-1] else np.zeros(0, np.uint8)
    lights = dict(xy=xy, switch_time=lights_df["switch-time"].to_numpy(np.float64), face_ptr=face_ptr,
                  face_vec=face_vec, face_go=face_go, degree=degree)
    return lights, (lights_df["xbin"].to_numpy(np.int64), lights_df["ybin"].to_numpy(np.int64))


def world_from_dataframes(cars_df, lights_df, graph, params=None) -> World:
    """``Cars(init_state, graph)`` + ``TrafficLights(light_state, graph)`` -> ``World`` (cars.py:23-41, 110-121)."""
    grid = Grid.from_axis(graph.axis)
    n = len(cars_df)
    pos = np.stack([cars_df["x"].to_numpy(np.float64), cars_df["y"].to_numpy(np.float64)], 1) if n else np.zeros((0, 2))
    xs = [np.asarray(p, np.float64).ravel() for p in cars_df["xpath"]] if n else []
    ys = [np.asarray(p, np.float64).ravel() for p in cars_df["ypath"]] if n else []
    lens = np.asarray([len(p) for p in xs], np.int64)
    if any(len(a) != len(b) for a, b in zip(xs, ys)):
        raise ValueError("xpath / ypath lengths differ")
    path_ptr = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=path_ptr[1:])
    path_xy = np.stack([np.concatenate(xs), np.concatenate(ys)], 1) if path_ptr[-1] else np.zeros((0, 2))
    nodes = graph.G.nodes
    dest_xy = np.asarray([[nodes[d]["x"], nodes[d]["y"]] for d in cars_df["destination"]], np.float64).reshape(-1, 2)
    lights, (lxb, lyb) = lights_from_dataframe(lights_df) if lights_df is not None else (None, (None, None))
    light_cell = None
    if lights is not None:
        inr = (lxb >= 0) & (lxb < grid.ncx) & (lyb >= 0) & (lyb < grid.ncy)
        light_cell = np.where(inr, lyb * grid.ncx + lxb, -1)
    w = make_world(graph.axis, pos, path_ptr, path_xy, dest_xy, lights, light_cell,
                   vel=np.stack([cars_df["vx"].to_numpy(np.float64), cars_df["vy"].to_numpy(np.float64)], 1) if n else None,
                   route_time=cars_df["route-time"].to_numpy(np.float64) if n else None, params=params)
    return w
