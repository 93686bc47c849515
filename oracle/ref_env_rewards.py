"""TEST INFRASTRUCTURE - records what the UNMODIFIED reference's ``environment.Env.step`` (environment.py:97-152) returns as
(reward, done) for given sequences of agent trip times, as a fixture for ``traffic_b200.ensemble.reward`` /
``rewards_of``.  Only usable in the build container (needs /root/reference); the fixture it writes is committed as
tests/golden/env_rewards.json.

The rollout itself is stubbed out (``simulation_step`` returns "arrived" at once, the agent's 'route-time' is the next
value of the sequence), so exactly the reward bookkeeping lines run - on the reference's own code.

Usage: python oracle/ref_env_rewards.py
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_harness import GOLDEN_DIR, install_reference  # noqa: E402


def reference_rewards(trip_times, dt):
    install_reference()
    if "animate" not in sys.modules:   # environment.py imports animate -> matplotlib (absent here): stub the module
        m = types.ModuleType("animate")
        m.Animator = object
        sys.modules["animate"] = m
    import environment as ref_env
    env = ref_env.Env.__new__(ref_env.Env)
    env.route_times, env.dt, env.animate, env.agent = [], dt, False, 0
    env.high, env.low, env.shortest_route_thresh = 10, 2, 5
    seq = iter(trip_times)
    state = types.SimpleNamespace(loc={0: {"route-time": None}})
    env.cars_object = types.SimpleNamespace(state=state)
    env.refresh_stateview = lambda: types.SimpleNamespace(determine_state=lambda: ([True], None, None, None))
    env.simulation_step = lambda i: True
    out = []
    n = len(trip_times)
    for k in range(n):
        state.loc[0]["route-time"] = next(seq)
        _, reward, done, _ = env.step(0, (k, n))
        out.append([float(reward), bool(done)])
    return out


def main():
    rng = np.random.default_rng(7)
    cases = []
    for dt in (1e-3, 0.1):
        for n in (1, 4, 5, 12, 40):
            t = np.round(3.0 + rng.random(n) * 0.5, 3)
            if n >= 5:
                t[n // 2] = t[: n // 2].min() + 2 * dt      # within 5*dt of the running minimum -> bonus
                t[-1] = t.min()                              # a new minimum at the end
            cases.append({"dt": dt, "trip_times": t.tolist(), "expected": reference_rewards(t.tolist(), dt)})
    with open(os.path.join(GOLDEN_DIR, "env_rewards.json"), "w") as f:
        json.dump({"source": "environment.Env.step of the unmodified reference (oracle/ref_env_rewards.py)", "cases": cases}, f, indent=1)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
