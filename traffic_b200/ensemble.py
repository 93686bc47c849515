"""Replica ensembles: many independent rollouts of the same world (same graph, routes, light geometry) that differ in
their light timings - the parallel axis of ``learn.py`` / ``environment.Env`` (one RL "step" is a whole traffic rollout,
``environment.py:97-152``) and of randomised light-timing sweeps.

* ``Ensemble`` batches R replicas into ONE device simulation: the step kernel offsets the per-cell car table and the
  light-face states by the replica index, so a launch advances all replicas; the agent's arrival
  (``FrontView.end_of_route``, ``navigation.py:113-128``, tested before every step as in ``Env.simulation_step``,
  ``environment.py:154-173``) is recorded on the device (``DONE`` / ``TRIP_TIME`` / ``TRIP_STEP``), no per-step host sync.
* ``shard`` / ``gather_results`` split the replicas over ranks (one process per GPU) and collect the per-replica results
  with a single ``all_gather`` (NCCL on GPUs, gloo in the CPU tests); there is no data-path collective.
* ``reward`` restates the reward bookkeeping of ``Env.step`` (``environment.py:126-150``).
"""
from __future__ import annotations

import numpy as np

from . import native
from .state import World


def shard(n_replicas: int, world_size: int, rank: int):
    """Round-robin assignment replica -> rank (balances the rollout-length tail)."""
    return np.arange(rank, n_replicas, world_size, dtype=np.int64)


def random_switch_times(world: World, n_replicas: int, seed: int):
    """Per-replica light timings: ``round(U(0,1) * degree, 2)`` as ``models.determine_traffic_light_timer``
    (``models.py:31-38``); replica r uses generator stream ``seed + r`` so shards are reproducible on any rank."""
    degree = np.diff(world.face_ptr).astype(np.float64)
    out = np.empty((n_replicas, world.n_lights), np.float64)
    for r in range(n_replicas):
        out[r] = np.round(np.random.default_rng(seed + r).random(world.n_lights) * degree, 2)
    return out


def pick_agent(world: World, min_points: int = 8, max_points: int = 24) -> int:
    """A car whose rollout CAN terminate.  ``Env.simulation_step`` ends an episode when the agent is within 5 m of its
    destination on BOTH axes (``navigation.py:113-128``), but a car parks as soon as it "crosses" its last route point - a
    window of 1e-5 x |coordinate|, i.e. ~5.7 m in x but ~42 m in y at UTM scale - so only cars
    whose last edge runs mostly along x get close enough.  ``learn.py``'s own choice (car 17) never arrives on most
    worlds.  Returns the first car with a route of min_points..max_points points (long enough to meet lights, so that
    the replicas' light timings matter) and an x-dominant last edge; falls back to any x-dominant route; -1 if none."""
    lens = world.path_end - world.path_start
    prev = np.where((lens >= 2)[:, None], world.path_xy[np.maximum(world.path_end - 2, 0)], world.pos)
    last = world.path_xy[np.maximum(world.path_end - 1, 0)] - prev
    xdom = (lens >= 2) & (np.abs(last[:, 0]) > 2.0 * np.abs(last[:, 1]))
    idx = np.flatnonzero(xdom & (lens >= min_points) & (lens <= max_points))
    if not len(idx):
        idx = np.flatnonzero(xdom)
    return int(idx[0]) if len(idx) else -1


def reward(route_times, num, dt, high=10.0, thresh=5):
    """``Env.step`` reward for the episode whose trip time was just appended to ``route_times``
    (``environment.py:126-150``); ``num`` = (episode index, total)."""
    route_times = list(route_times)
    if len(route_times) < thresh:
        bonus, done = 0.0, False
    elif np.isclose(0, route_times[-1] - np.min(route_times), atol=5 * dt).all():
        bonus, done = high, True
    else:
        bonus, done = 0.0, False
    if num[0] < 1:
        r = 0.0
    else:
        delta = route_times[num[0] - 1] - route_times[num[0]] + bonus
        r = delta if delta > 0 else 0.0
    return r, done


def rewards_of(trip_times, dt, high=10.0, thresh=5):
    """Rewards ``Env.step`` would hand out if the given trip times were successive episodes (``environment.py:126-150``):
    episode k gets ``max(0, t[k-1] - t[k] + bonus)``, bonus = ``high`` when t[k] is within 5 dt of the best so far."""
    out, seen = [], []
    for k, t in enumerate(trip_times):
        seen.append(t)
        out.append(reward(seen, (k, len(trip_times)), dt, high, thresh)[0])
    return np.asarray(out, np.float64)


class Ensemble:
    """R replicas of ``world`` on one GPU.  ``replica_ids`` are the global ids of the replicas this rank owns."""

    def __init__(self, world: World, replica_ids, agent_car: int, seed: int = 0, switch_time=None, device=None,
                 write_aux: bool = False, precision: int = native.F64):
        self.world = world
        self.replica_ids = np.asarray(replica_ids, np.int64)
        R = len(self.replica_ids)
        if R < 1:
            raise ValueError("an ensemble shard needs at least one replica")
        if switch_time is None:
            all_sw = random_switch_times(world, int(self.replica_ids.max()) + 1, seed)
            switch_time = all_sw[self.replica_ids]
        self.switch_time = np.ascontiguousarray(switch_time, np.float64).reshape(R, world.n_lights)
        self.sim = native.DeviceSim(world, precision=precision, write_aux=write_aux, device=device, n_replicas=R,
                                    agent_car=agent_car, switch_time=self.switch_time)
        self.agent_car = agent_car

    def rollout(self, dt: float, max_steps: int, order: int = native.ORDER_LIGHTS_CARS):
        """Advance until every replica's agent has arrived or ``max_steps`` (the reference loop has no cap and may
        spin forever).  Returns float64 ``[R, 4]``: trip_time, steps, done, replica id."""
        # tb2_rollout: small ensembles run as ONE launch (one thread-block cluster per replica, world on chip, a replica
        # leaves when its agent has arrived); large ones as batched launches with arrived replicas frozen and a 4-byte
        # per replica readback every 256 steps
        self.sim.rollout(dt, max_steps, order)
        return self.results(max_steps)

    def results(self, steps_run: int):
        done = self.sim.download(native.DONE)
        trip_time = self.sim.download(native.TRIP_TIME)
        trip_step = self.sim.download(native.TRIP_STEP)
        if (done == 0).any():   # not arrived: report the current route time and the steps run so far
            rt = self.sim.download(native.ROUTE_TIME).reshape(len(done), -1)[:, self.agent_car]
            trip_time = np.where(done != 0, trip_time, rt)
            trip_step = np.where(done != 0, trip_step, steps_run)
        return np.stack([trip_time, trip_step.astype(np.float64), done.astype(np.float64),
                         self.replica_ids.astype(np.float64)], 1)


def gather_results(local: np.ndarray, n_replicas: int, group=None, device=None):
    """all_gather of the per-replica result rows (NCCL over NVLink on GPUs, gloo on CPU).  Shards may differ in length
    by one (round-robin), so rows are padded to ceil(R / world) and re-assembled by replica id.  Returns ``[R, 4]``."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        out = np.full((n_replicas, local.shape[1]), np.nan)
        out[local[:, 3].astype(np.int64)] = local
        return out
    world_size = dist.get_world_size(group)
    rows = -(-n_replicas // world_size)
    pad = np.full((rows, local.shape[1]), -1.0)
    pad[:len(local)] = local
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    t = torch.from_numpy(pad).to(dev)
    allt = torch.empty((world_size * rows, local.shape[1]), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(allt, t, group=group)
    allr = allt.cpu().numpy()
    allr = allr[allr[:, 3] >= 0]
    out = np.full((n_replicas, local.shape[1]), np.nan)
    out[allr[:, 3].astype(np.int64)] = allr
    return out
