"""Drop-in ``Cars`` / ``TrafficLights`` (the reference's ``cars.py`` API) backed by the device-resident simulation.

Same constructors, attributes, ``update`` signatures and return conventions as ``cars.py:22-133`` so that
``animate.py``, ``environment.py``, ``learn.py``, ``artist.py`` and ``test/profile_cars_update.py`` run unchanged when
``traffic_b200/dropin`` is placed before the reference on ``sys.path`` (it provides a module named ``cars``).

Differences a caller can observe, all deliberate:
* ``.state`` / the value returned by ``update`` is a ``LazyFrame``: it behaves like the pandas DataFrame for reading
  (``iterrows``, ``.loc``, ``[col]``, ``len``) but is only materialised from device memory when it is touched, so a
  driver loop that does not look at the state every step (``environment.py:120-124``) pays no device->host copy.
* ``TrafficLights`` keeps its arrays on the GPU from construction on; when a ``Cars`` object first sees its state in
  ``update(dt, lights)`` both are bound into one device simulation.  A foreign lights DataFrame (e.g. the reference's
  own ``TrafficLights``) is accepted too; its ``go-values`` are then uploaded every step.
* no CPU fallback: constructing either class without the CUDA library / a GPU raises.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import native
from .state import World, make_world, world_from_dataframes, lights_from_dataframe, Grid


def _reference_params():
    """Honour monkey-patched module constants of the reference's ``simulation`` module (simulation.py:13-16)."""
    sim = sys.modules.get("simulation")
    if sim is None:
        return None
    try:
        return dict(speed_limit=float(sim.speed_limit), stop_distance=float(sim.stop_distance),
                    free_distance=float(sim.free_distance), default_acceleration=float(sim.default_acceleration))
    except AttributeError:
        return None


class _LazyLoc:
    """``state.loc``: a single integer row label is served from a few small device->host copies (what
    ``Env.simulation_step`` / ``Env.step`` do every step, ``environment.py:126,161``); everything else falls through to
    the fully materialised DataFrame."""

    __slots__ = ("_owner",)

    def __init__(self, owner):
        self._owner = owner

    def __getitem__(self, key):
        row = getattr(self._owner, "_row", None)
        if row is not None and isinstance(key, (int, np.integer)) and not isinstance(key, bool):
            return row(int(key))
        return self._owner._materialize().loc[key]


class LazyFrame:
    """Read-only stand-in for the reference's ``state`` DataFrame.  Device state is copied back only when it is looked
    at, and only as much as is looked at: ``state[col]`` refreshes that column (group), ``state.loc[i]`` that row,
    anything else (``iterrows``, arbitrary attribute access) the whole frame."""

    __slots__ = ("_owner",)

    def __init__(self, owner):
        object.__setattr__(self, "_owner", owner)

    def _df(self):
        return self._owner._materialize()

    @property
    def loc(self):
        return _LazyLoc(self._owner)

    def __getattr__(self, name):
        return getattr(self._df(), name)

    def __getitem__(self, key):
        cols = getattr(self._owner, "_refresh_columns", None)
        if cols is not None and (isinstance(key, str) or (isinstance(key, list) and all(isinstance(k, str) for k in key))):
            return cols([key] if isinstance(key, str) else key)[key]
        return self._df()[key]

    def __setitem__(self, key, value):
        raise TypeError("traffic_b200 state is device-resident; build a new Cars/TrafficLights object instead of "
                        "assigning into .state")

    def __len__(self):
        return self._owner._n_rows

    def __iter__(self):
        return iter(self._df())

    def __repr__(self):
        return repr(self._df())


class TrafficLights:
    """cars.TrafficLights (cars.py:109-133)."""

    def __init__(self, light_state, graph):
        self.init_state = light_state
        self.time_elapsed = 0
        self.xbins = np.arange(graph.axis[0], graph.axis[1], 200)
        self.ybins = np.arange(graph.axis[2], graph.axis[3], 200)
        self._graph = graph
        self._frame = light_state.copy()
        self._n_rows = len(light_state)
        self._dirty = False
        self._state = LazyFrame(self)
        lights, (xb, yb) = lights_from_dataframe(light_state)
        self._lights = lights
        self._face_ptr = lights["face_ptr"] if lights is not None else np.zeros(1, np.int32)
        grid = Grid.from_axis(graph.axis)
        cell = None
        if lights is not None:
            inr = (xb >= 0) & (xb < grid.ncx) & (yb >= 0) & (yb < grid.ncy)
            cell = np.where(inr, yb * grid.ncx + xb, -1)
        self._light_cell = cell
        # lights-only device simulation until a Cars object adopts these lights
        w = make_world(graph.axis, np.zeros((0, 2)), np.zeros(1, np.int64), np.zeros((0, 2)), np.zeros((0, 2)),
                       lights, cell)
        self._sim = native.DeviceSim(w).retain()

    @property
    def state(self):
        return self._state

    def update(self, dt):
        self.time_elapsed += dt
        self._sim.lights_update(dt)
        self._dirty = True
        return self._state

    # -- internals -------------------------------------------------------------------------------------------
    def _current_go(self):
        return self._sim.download(native.FACE_GO)

    def _materialize(self):
        if self._dirty:
            go = self._current_go().astype(bool)
            fp = self._face_ptr
            self._frame["go-values"] = [go[fp[i]:fp[i + 1]] for i in range(self._n_rows)]
            self._dirty = False
        return self._frame

    def _adopt(self, sim):
        """A Cars object built a combined simulation holding these lights: move onto it."""
        if self._sim is not sim:
            self._sim.release()
            self._sim = sim.retain()


# column groups of the cars DataFrame: one group = one (set of) device field(s)
_GROUP_OF = {"x": "pos", "y": "pos", "vx": "vel", "vy": "vel", "route-time": "route-time",
             "distance-to-node": "dist", "distance-to-car": "dist", "distance-to-red-light": "dist",
             "xbin": "bins", "ybin": "bins", "xpath": "paths", "ypath": "paths", "route": "route"}
_GROUPS = tuple(sorted(set(_GROUP_OF.values())))


def _falsy(d, none):
    """Mixed-type distance columns of the reference (Appendix B-4): a distance, False, or (lights only) None."""
    out = d.astype(object)
    zero = d == 0.0
    if none:
        out[zero & ~np.signbit(d)] = False
        out[zero & np.signbit(d)] = None
    else:
        out[zero] = False
    return out


class Cars:
    """cars.Cars (cars.py:22-106)."""

    def __init__(self, init_state, graph, serialize=False):
        native.lib()  # fail loudly now if the CUDA library is missing
        self.init_state = init_state
        self.time_elapsed = 0
        self.lights = 0
        self.graph = graph
        self.serialize = serialize
        self.serialize_path = "data_store"
        self.axis = self.graph.axis
        self.stop_distance = 5
        self._frame = init_state.copy()
        self._n_rows = len(init_state)
        self._stale = set()       # column groups of self._frame that are older than the device state
        self._state = LazyFrame(self)
        self._sim = None
        self._world = None
        self._lights_owner = None
        self._foreign = None      # the caller's own lights DataFrame when it is not one of our TrafficLights
        self._updates = 0

    @property
    def state(self):
        return self._state

    # ------------------------------------------------------------------------------------------------------
    def _bind(self, lights):
        owner = lights._owner if isinstance(lights, LazyFrame) and isinstance(lights._owner, TrafficLights) else None
        if self._sim is not None and owner is self._lights_owner and (
                (owner is not None and owner._sim is self._sim) or (owner is None and lights is self._foreign)):
            return owner   # (owner._sim differs when another Cars object has adopted these lights since: rebind)
        lights_df = owner._materialize() if owner is not None else lights
        # current car state (if we are re-binding mid-run, continue from the device state)
        cars_df = self._materialize()
        w = world_from_dataframes(cars_df, lights_df, self.graph, params=_reference_params())
        if owner is not None:
            w.lights_time = float(owner.time_elapsed)
        # (on a re-bind mid-run the frame's xpath/ypath already hold only the remaining route, so cursors carry over)
        if self._sim is not None:
            self._sim.release()    # a TrafficLights object that adopted it keeps it alive
        self._sim = native.DeviceSim(w).retain()
        self._world = w
        self._path_start0 = w.path_start.copy()
        self._lights_owner = owner
        self._foreign = None if owner is not None else lights
        if owner is not None:
            owner._adopt(self._sim)
        return owner

    def update(self, dt, lights):
        self.lights = lights
        self.time_elapsed += dt
        owner = self._bind(lights)
        if owner is None:
            # foreign lights DataFrame: its go-values are this step's input (H2D every step)
            go = np.concatenate([np.asarray(v).astype(np.uint8) for v in lights["go-values"]]) if len(lights) else np.zeros(0, np.uint8)
            self._sim.upload(native.FACE_GO, go)
        self._sim.cars_update(dt)
        self._stale = set(_GROUPS)
        self._updates += 1
        if self.serialize:
            self.write_state()
        return self._state

    def find_obstacles(self):
        """cars.py:98-106 - the three distance lists of the last update."""
        st = self._materialize()
        return list(st["distance-to-node"]), list(st["distance-to-car"]), list(st["distance-to-red-light"])

    def write_state(self):
        """cars.py:43-58 (pyarrow instead of the absent fastparquet; mixed-type distance columns become floats)."""
        st = self._materialize().copy()
        for c in ("distance-to-car", "distance-to-red-light"):
            st[c] = [float(v) if v not in (False, None) else 0.0 for v in st[c]]
        st["route"] = st["route"].astype(str)
        t = str(round(self.time_elapsed, 4)).replace(".", "").zfill(4)
        d = os.path.join(self.serialize_path, "dt=%s" % t)
        os.makedirs(d, exist_ok=True)
        st.to_parquet(os.path.join(d, "cars_state_at_%s.parquet.gz" % t), engine="pyarrow", compression="gzip")

    # ------------------------------------------------------------------------------------------------------
    def _materialize(self):
        """The whole DataFrame, current (every stale column group is refreshed)."""
        if self._stale:
            self._refresh_columns(list(self._frame.columns))
        return self._frame

    def _refresh_columns(self, cols):
        """Bring the column groups that hold ``cols`` up to date with the device; returns the frame."""
        need = {_GROUP_OF[c] for c in cols if c in _GROUP_OF} & self._stale
        if not need:
            return self._frame
        sim, f, w, n = self._sim, self._frame, self._world, self._n_rows
        if "pos" in need:
            pos = sim.download(native.POS).reshape(-1, 2).astype(np.float64)
            if sim.precision != native.F64:
                pos = pos + sim.origin
            w.pos = pos
            f["x"], f["y"] = pos[:, 0], pos[:, 1]
        if "vel" in need:
            w.vel = sim.download(native.VEL).reshape(-1, 2).astype(np.float64)
            f["vx"], f["vy"] = w.vel[:, 0], w.vel[:, 1]
        if "route-time" in need:
            w.route_time = sim.download(native.ROUTE_TIME).astype(np.float64)
            f["route-time"] = w.route_time
        if "dist" in need:
            w.d_node = sim.download(native.D_NODE).astype(np.float64)
            w.d_car = sim.download(native.D_CAR).astype(np.float64)
            w.d_light = sim.download(native.D_LIGHT).astype(np.float64)
            f["distance-to-node"] = w.d_node
            f["distance-to-car"] = _falsy(w.d_car, none=False)
            f["distance-to-red-light"] = _falsy(w.d_light, none=True)
        if "bins" in need:
            w.cell = sim.download(native.CELL)
            ncx = w.grid.ncx
            f["xbin"] = (w.cell % ncx).astype(np.int64)
            f["ybin"] = (w.cell // ncx).astype(np.int64)
        if "paths" in need:
            # remaining route of every car as Python lists (the reference pops the head of these lists, simulation.py:47-52)
            w.cursor = sim.download(native.CURSOR)
            px, py = w.path_xy[:, 0].tolist(), w.path_xy[:, 1].tolist()
            cur, end = w.cursor.tolist(), w.path_end.tolist()
            xs = np.empty(n, object)
            ys = np.empty(n, object)
            for i, (a, b) in enumerate(zip(cur, end)):   # element-wise: numpy would broadcast equal-length lists to 2-D
                xs[i] = px[a:b]
                ys[i] = py[a:b]
            f["xpath"], f["ypath"] = xs, ys
        if "route" in need and "route" in f.columns:
            f["route"] = np.nan  # the reference wipes this column on the first update (simulation.py:30,76; B-7)
        self._stale -= need
        return f

    def _row(self, i):
        """``state.loc[i]`` as a Series: the stale cells of row i come from a few bytes of device memory each."""
        f, sim, w = self._frame, self._sim, self._world
        if not self._stale:
            return f.loc[i]
        if i < 0 or i >= self._n_rows:
            raise KeyError(i)
        row = f.loc[i].copy()
        st = self._stale
        if "pos" in st:
            p = sim.download_range(native.POS, 2 * i, 2).astype(np.float64)
            if sim.precision != native.F64:
                p = p + sim.origin
            row["x"], row["y"] = p[0], p[1]
        if "vel" in st:
            v = sim.download_range(native.VEL, 2 * i, 2)
            row["vx"], row["vy"] = float(v[0]), float(v[1])
        if "route-time" in st:
            row["route-time"] = float(sim.download_range(native.ROUTE_TIME, i, 1)[0])
        if "dist" in st:
            row["distance-to-node"] = float(sim.download_range(native.D_NODE, i, 1)[0])
            row["distance-to-car"] = _falsy(sim.download_range(native.D_CAR, i, 1).astype(np.float64), none=False)[0]
            row["distance-to-red-light"] = _falsy(sim.download_range(native.D_LIGHT, i, 1).astype(np.float64), none=True)[0]
        if "bins" in st:
            c = int(sim.download_range(native.CELL, i, 1)[0])
            row["xbin"], row["ybin"] = c % w.grid.ncx, c // w.grid.ncx
        if "paths" in st:
            a, b = int(sim.download_range(native.CURSOR, i, 1)[0]), int(w.path_end[i])
            row["xpath"], row["ypath"] = w.path_xy[a:b, 0].tolist(), w.path_xy[a:b, 1].tolist()
        if "route" in st and "route" in row.index:
            row["route"] = np.nan
        return row
