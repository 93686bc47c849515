"""CPU: pins the C oracle (oracle/traffic_oracle.c) against vectors recorded from the unmodified reference
(oracle/ref_harness.py).  Discrete outputs (route cursor, bins, leader index, light faces, truthiness of the
obstacle distances) must be bit-exact at every step; floats within 1e-9 relative (observed: a few ulp)."""
import numpy as np
import pytest

from oracle import oracle
from tests import util

NAMES = util.golden_names()


def test_fixtures_present():
    assert len(NAMES) >= 4, "golden fixtures missing: run python oracle/ref_harness.py --all in the build container"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_free_running(name):
    g = util.load_golden(name)
    w = util.world_from_golden(g)
    start = w.path_start.copy()
    dt, T, order = float(g["dt"]), int(g["steps"]), util.order_code(g)
    report, worst = [], 0.0
    for t in range(T):
        oracle.run(w, dt, 1, order)
        ok, m = util.compare_step(util.snapshot(w, start), g, t, w.grid.ncx, 1e-9, report)
        worst = max(worst, m)
        if not ok:
            break
    assert not report, "\n".join(report[:10])
    print("%s: %d steps, worst float rel err %.2e" % (name, T, worst))
