"""Torture tests for the FILTERED exact predicates of the fp64 kernel (step_kernel.cuh): inputs are constructed to fall
inside the fp32 filters' uncertainty bands - probe points within a millimetre of the linspace tolerance, positions within
centimetres of a bin edge, light faces within 1e-5 rad of the anti-parallel threshold - so that the exact fp64 routines
(on_axis_exact, digitize_exact, antiparallel_exact) decide, and every decision is compared with the CPU oracle."""
import numpy as np
import pytest

from tests.test_gpu_parity import _compare_worlds, FLOAT_RTOL
from traffic_b200.state import make_world

pytestmark = pytest.mark.gpu

X0, Y0 = 566000.0, 4186000.0
AXIS = (X0 - 50.0, X0 + 16050.0, Y0 - 50.0, Y0 + 16050.0)   # 80 x 80 bins


def _pairs_world(rng, n_pairs, make_probe):
    """n_pairs isolated 200 m bins, each with car 2k (the observer) and car 2k+1 (the probe = the observer's candidate)."""
    pos, paths, dest = [], [], []
    side = int(np.ceil(np.sqrt(n_pairs)))
    for k in range(n_pairs):
        cx = AXIS[0] + 200.0 * (k % side) + 100.0 + rng.uniform(-10, 10)     # bin centre-ish
        cy = AXIS[2] + 200.0 * (k // side) + 100.0 + rng.uniform(-10, 10)
        # short x extent -> few x samples, spaced wider than twice the x tolerance (0.57 m), so that a probe placed at
        # "tolerance +- 1.5 mm" from one sample is not within tolerance of its neighbour; |dy| > 42 m keeps the observer
        # outside its node-crossing window (navigation.py:104-106)
        tx = cx + rng.uniform(2.2, 7.5) * rng.choice([-1, 1])
        ty = cy + rng.uniform(43.0, 60.0) * rng.choice([-1, 1])
        px, py = make_probe(rng, cx, cy, tx, ty)
        pos += [[cx, cy], [px, py]]
        paths += [np.array([[tx, ty], [tx + 40.0, ty + 25.0]]), np.zeros((0, 2))]
        dest += [[tx + 40.0, ty + 25.0], [px, py]]
    ptr = np.concatenate([[0], np.cumsum([len(p) for p in paths])])
    return make_world(AXIS, np.array(pos), ptr, np.concatenate(paths), np.array(dest))


def _sample(start, stop, k):
    n = int(abs(stop - start))
    if n <= 1 or k >= n - 1:
        return stop if n > 1 else start
    return k * ((stop - start) / (n - 1)) + start


def _near_tolerance_probe(rng, cx, cy, tx, ty):
    """probe whose x lies at (tolerance +- up to 1.5 mm) from a linspace sample and whose y sits on a sample."""
    nx, ny = int(abs(tx - cx)), int(abs(ty - cy))
    kx, ky = rng.integers(0, max(nx, 1)), rng.integers(0, max(ny, 1))
    sx, sy = _sample(cx, tx, kx), _sample(cy, ty, ky)
    tol = 1e-8 + 1e-6 * abs(sx)
    off = (tol + rng.uniform(-1.5e-3, 1.5e-3)) * rng.choice([-1, 1])
    return sx + off, sy + rng.uniform(-0.5, 0.5)


def _run_both(w, steps=3):
    from oracle import oracle
    from traffic_b200 import native
    wg, wc = w.copy(), w.copy()
    sim = native.DeviceSim(wg)
    for s in range(steps):
        sim.step(1e-3, 1, 1)
        oracle.run(wc, 1e-3, 1, 1)
        sim.download_world(wg)
        _compare_worlds(wg, wc, FLOAT_RTOL, "step %d" % (s + 1))
    return wc


def test_linspace_membership_inside_the_filter_band():
    rng = np.random.default_rng(1)
    w = _pairs_world(rng, 4000, _near_tolerance_probe)
    wc = _run_both(w, steps=1)
    hits = (wc.leader[0::2] >= 0).mean()
    _run_both(w, steps=4)
    assert 0.2 < hits < 0.8, "construction should straddle the tolerance (got %.2f accepted)" % hits


def test_bin_edges_inside_the_filter_band():
    """cars within +-0.3 m of a bin edge (and exactly on edges): np.digitize must be reproduced exactly."""
    rng = np.random.default_rng(2)
    n = 6000
    ex = np.arange(AXIS[0], AXIS[1], 200.0)
    ey = np.arange(AXIS[2], AXIS[3], 200.0)
    x = ex[rng.integers(1, len(ex) - 1, n)] + rng.uniform(-0.3, 0.3, n)
    y = ey[rng.integers(1, len(ey) - 1, n)] + rng.uniform(-0.3, 0.3, n)
    x[:200] = ex[rng.integers(0, len(ex), 200)]          # exactly on an edge
    y[200:400] = ey[rng.integers(0, len(ey), 200)]
    x[400:420] = AXIS[0] - rng.uniform(0, 30, 20)        # left of the first edge -> bin 0
    y[420:440] = ey[-1] + rng.uniform(0, 30, 20)         # beyond the last edge -> bin nb
    pos = np.stack([x, y], 1)
    tgt = pos + rng.uniform(30, 60, (n, 2)) * rng.choice([-1, 1], (n, 2))
    ptr = np.arange(n + 1)
    w = make_world(AXIS, pos, ptr, tgt, tgt)
    _run_both(w, steps=4)


def test_antiparallel_faces_inside_the_filter_band():
    """a red face whose angle to the car->light vector is within 2e-5 rad of T = (pi - 0.1) / (1 + 1e-5)."""
    rng = np.random.default_rng(3)
    n = 3000
    T = (np.pi - 0.1) / (1 + 1e-5)
    side = int(np.ceil(np.sqrt(n)))
    pos, tgt, lxy, fvec = [], [], [], []
    for k in range(n):
        cx = AXIS[0] + 200.0 * (k % side) + 60.0
        cy = AXIS[2] + 200.0 * (k // side) + 60.0
        ang = rng.uniform(0.3, 1.2)
        L = rng.uniform(30.0, 60.0)
        tx, ty = cx + L * np.cos(ang), cy + L * np.sin(ang)
        pos.append([cx, cy]); tgt.append([tx, ty]); lxy.append([tx, ty])       # the light sits on the target node
        th = ang + (T + rng.uniform(-2e-5, 2e-5)) * rng.choice([-1, 1])       # face at angle ~T from the approach
        fvec.append([rng.uniform(5, 80) * np.cos(th), rng.uniform(5, 80) * np.sin(th)])
    lights = dict(xy=np.array(lxy), switch_time=np.full(n, 7.77), face_ptr=np.arange(n + 1, dtype=np.int32),
                  face_vec=np.array(fvec), face_go=np.zeros(n, np.uint8))
    ptr = np.arange(n + 1)
    w = make_world(AXIS, np.array(pos), ptr, np.array(tgt), np.array(tgt), lights)
    wc = _run_both(w, steps=2)
    red = (wc.d_light != 0).mean()
    assert 0.2 < red < 0.8, "construction should straddle the angle threshold (got %.2f red)" % red


def test_degenerate_geometry():
    """axis-aligned segments (empty linspace -> nothing is ever detected), |delta| in [1, 2) (one-sample linspace),
    co-located cars (distance 0.0 is falsy) and a car exactly on its target (0/0 -> NaN velocity)."""
    pos = np.array([[X0 + 100.0, Y0 + 100.0], [X0 + 130.0, Y0 + 100.0],      # 0 -> east, axis aligned; 1 on its line
                    [X0 + 500.0, Y0 + 500.0], [X0 + 500.7, Y0 + 500.9],      # 2: |dx|, |dy| in [1,2)
                    [X0 + 900.0, Y0 + 900.0], [X0 + 900.0, Y0 + 900.0],      # 4, 5 co-located, same route
                    [X0 + 1300.0, Y0 + 1300.0]])                              # 6 sits exactly on its only target
    paths = [np.array([[X0 + 160.0, Y0 + 100.0]]), np.zeros((0, 2)),
             np.array([[X0 + 501.5, Y0 + 501.5]]), np.zeros((0, 2)),
             np.array([[X0 + 950.0, Y0 + 930.0]]), np.array([[X0 + 950.0, Y0 + 930.0]]),
             np.array([[X0 + 1300.0, Y0 + 1300.0]])]
    dest = np.array([[X0 + 160.0, Y0 + 100.0], [X0 + 130.0, Y0 + 100.0], [X0 + 501.5, Y0 + 501.5], [X0 + 500.7, Y0 + 500.9],
                     [X0 + 950.0, Y0 + 930.0], [X0 + 950.0, Y0 + 930.0], [X0 + 1300.0, Y0 + 1300.0]])
    ptr = np.concatenate([[0], np.cumsum([len(p) for p in paths])])
    w = make_world(AXIS, pos, ptr, np.concatenate(paths), dest)
    wc = _run_both(w, steps=1)
    assert wc.leader[0] == -1 and wc.leader[4] == -1 and wc.leader[5] == -1
