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


class LazyFrame:
    """Read-only stand-in for the reference's ``state`` DataFrame, materialised on first touch after each update."""

    __slots__ = ("_owner",)

    def __init__(self, owner):
        object.__setattr__(self, "_owner", owner)

    def _df(self):
        return self._owner._materialize()

    def __getattr__(self, name):
        return getattr(self._df(), name)

    def __getitem__(self, key):
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
        self._sim = native.DeviceSim(w)
        self._shared = False

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
            if not self._shared:
                self._sim.close()
            self._sim = sim
            self._shared = True


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
        self._dirty = False
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
        if self._sim is not None and owner is self._lights_owner and (owner is not None or lights is self._foreign):
            return owner
        lights_df = owner._materialize() if owner is not None else lights
        # current car state (if we are re-binding mid-run, continue from the device state)
        cars_df = self._materialize()
        w = world_from_dataframes(cars_df, lights_df, self.graph, params=_reference_params())
        if owner is not None:
            w.lights_time = float(owner.time_elapsed)
        # (on a re-bind mid-run the frame's xpath/ypath already hold only the remaining route, so cursors carry over)
        if self._sim is not None:
            self._sim.close()
        self._sim = native.DeviceSim(w)
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
        self._dirty = True
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
        if not self._dirty:
            return self._frame
        w = self._sim.download_world(self._world)
        f = self._frame
        n = self._n_rows
        f["x"], f["y"] = w.pos[:, 0], w.pos[:, 1]
        f["vx"], f["vy"] = w.vel[:, 0], w.vel[:, 1]
        f["route-time"] = w.route_time
        f["distance-to-node"] = w.d_node
        # mixed-type columns of the reference (Appendix B-4): a distance, False, or (lights only) None
        dc = w.d_car.astype(object)
        dc[w.d_car == 0.0] = False
        dl = w.d_light.astype(object)
        zero = w.d_light == 0.0
        dl[zero & ~np.signbit(w.d_light)] = False
        dl[zero & np.signbit(w.d_light)] = None
        f["distance-to-car"] = dc
        f["distance-to-red-light"] = dl
        ncx = w.grid.ncx
        f["xbin"] = (w.cell % ncx).astype(np.int64)
        f["ybin"] = (w.cell // ncx).astype(np.int64)
        # remaining route of every car as Python lists (the reference pops the head of these lists, simulation.py:47-52)
        px, py = w.path_xy[:, 0].tolist(), w.path_xy[:, 1].tolist()
        cur, end = w.cursor.tolist(), w.path_end.tolist()
        xs = np.empty(n, object)
        ys = np.empty(n, object)
        for i, (a, b) in enumerate(zip(cur, end)):   # element-wise: numpy would broadcast equal-length lists to 2-D
            xs[i] = px[a:b]
            ys[i] = py[a:b]
        f["xpath"], f["ypath"] = xs, ys
        if self._updates > 0 and "route" in f.columns:
            f["route"] = np.nan  # the reference wipes this column on the first update (simulation.py:30,76; B-7)
        self._dirty = False
        return f
