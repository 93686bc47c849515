"""Randomised small worlds (not grids): random polylines with short, long, axis-aligned, duplicated and zero-length
segments, cars on and off their routes, random lights with 1-6 faces at random places, several cars and lights per bin,
cars outside the binned area - CUDA path vs the C oracle, every discrete output bit-exact at every step (NaNs included)."""
import numpy as np
import pytest

from tests.test_gpu_parity import _compare_worlds, FLOAT_RTOL
from traffic_b200.state import make_world

pytestmark = pytest.mark.gpu

X0, Y0 = 583000.0, 4507000.0


def random_world(seed):
    rng = np.random.default_rng(seed)
    size = rng.choice([400.0, 900.0, 2500.0])
    axis = (X0, X0 + size, Y0, Y0 + size)
    n = int(rng.integers(1, 400))
    pos, paths, dest = [], [], []
    for _ in range(n):
        p = np.array([X0, Y0]) + rng.uniform(-30, size + 30, 2)          # a few cars outside the binned area
        k = int(rng.integers(0, 7))
        pts, cur = [], p.copy()
        for s in range(k):
            kind = rng.integers(0, 10)
            if kind == 0:
                step = np.array([rng.uniform(1, 120), 0.0])               # axis aligned -> empty linspace
            elif kind == 1:
                step = np.zeros(2)                                        # duplicate point -> NaN angle / 0-0 velocity
            elif kind == 2:
                step = rng.uniform(-1.9, 1.9, 2)                          # n in {0, 1}
            elif kind == 3:
                step = rng.uniform(-6, 6, 2)                              # inside the crossing window
            else:
                step = rng.uniform(-150, 150, 2)
            cur = cur + step
            pts.append(cur.copy())
        if k and rng.random() < 0.3:
            p = pts[0] + rng.uniform(-3, 3, 2)                            # start next to the first route point
        pos.append(p)
        paths.append(np.array(pts).reshape(-1, 2))
        dest.append(pts[-1] if k else p + rng.uniform(-40, 40, 2))
    L = int(rng.integers(0, 60))
    lxy = np.array([X0, Y0]) + rng.uniform(0, size, (L, 2))
    # put some lights exactly on route points (that is where the reference's lights are)
    allpts = np.concatenate([p for p in paths if len(p)]) if any(len(p) for p in paths) else np.zeros((0, 2))
    for l in range(L):
        if len(allpts) and rng.random() < 0.7:
            lxy[l] = allpts[rng.integers(0, len(allpts))]
    deg = rng.integers(1, 7, L)
    face_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    F = int(face_ptr[-1])
    face_vec = rng.uniform(-100, 100, (F, 2))
    face_go = (rng.random(F) < 0.4).astype(np.uint8)
    lights = dict(xy=lxy, switch_time=np.round(rng.random(L) * deg, 2), face_ptr=face_ptr, face_vec=face_vec, face_go=face_go)
    ptr = np.concatenate([[0], np.cumsum([len(p) for p in paths])])
    path_xy = np.concatenate(paths) if ptr[-1] else np.zeros((0, 2))
    return make_world(axis, np.array(pos).reshape(-1, 2), ptr, path_xy, np.array(dest).reshape(-1, 2), lights)


@pytest.mark.parametrize("seed", range(40))
def test_random_world_matches_oracle(seed):
    from oracle import oracle
    from traffic_b200 import native
    w = random_world(seed)
    wg, wc = w.copy(), w.copy()
    sim = native.DeviceSim(wg)
    order = seed % 2
    dt = [1e-3, 0.1, 0.01][seed % 3]
    done = 0
    for n in (1, 1, 3, 10, 25):
        sim.step(dt, n, order)
        oracle.run(wc, dt, n, order)
        done += n
        sim.download_world(wg)
        # NaN positions (0/0 velocities) make the reference itself raise on the NEXT step (int(nan)); stop there
        _compare_worlds(wg, wc, FLOAT_RTOL, "seed %d step %d" % (seed, done))
        if np.isnan(wc.pos).any():
            break
