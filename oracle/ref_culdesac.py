"""TEST INFRASTRUCTURE - records what the UNMODIFIED reference's ``simulation.init_culdesac_start_location``
(simulation.py:259-326) builds on a synthetic graph with dead-end stubs, as a fixture for the drop-in initialiser
``traffic_b200.init.init_culdesac_start_location``.  Build container only; writes tests/golden/culdesac_init.json.

Usage: python oracle/ref_culdesac.py
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_harness import GOLDEN_DIR, install_reference  # noqa: E402

GRAPH = dict(nx_=10, ny_=10, seed=1, culdesacs=12)
N_CARS = 9


def main():
    _, ref_sim = install_reference()
    from traffic_b200 import synth
    graph = synth.SynthGraph(**GRAPH)
    cars = ref_sim.init_culdesac_start_location(N_CARS, graph)
    out = {"source": "simulation.init_culdesac_start_location of the unmodified reference (oracle/ref_culdesac.py)",
           "graph": GRAPH, "n": N_CARS,
           "cars": [{k: (row[k] if not hasattr(row[k], "item") else row[k].item()) for k in
                     ("x", "y", "origin", "destination", "route", "xpath", "ypath", "xbin", "ybin")}
                    for _, row in cars.iterrows()]}
    with open(os.path.join(GOLDEN_DIR, "culdesac_init.json"), "w") as f:
        json.dump(out, f)
    print("wrote %d cars" % len(cars))


if __name__ == "__main__":
    main()
