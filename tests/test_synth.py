"""CPU: the synthetic-world generators used by bench.py produce what the reference's own initialisers produce on the same
graph (light geometry), valid shortest-path routes, and worlds the oracle can step."""
import numpy as np

from tests import util
from traffic_b200 import synth


def _graph(name):
    from oracle.ref_harness import CONFIGS
    return synth.SynthGraph(**CONFIGS[name]["g"]), CONFIGS[name]


def test_make_lights_geometry_equals_reference_init():
    for name in ("c1_dt0.001_p2_s1", "c2_dt0.001_p10_s0"):
        graph, cfg = _graph(name)
        g = util.load_golden(name)
        L = synth.make_lights(graph, prescale=cfg["prescale"], seed=0)
        assert np.array_equal(L["xy"], g["light_xy"])
        assert np.array_equal(L["face_ptr"], g["face_ptr"])
        assert np.array_equal(L["face_vec"], g["face_vec"])
        assert np.array_equal(L["face_go"].astype(np.uint8), g["face_go0"])
        assert ((L["switch_time"] >= 0) & (L["switch_time"] <= L["degree"])).all()


def test_make_routes_are_shortest_paths():
    from scipy.sparse.csgraph import dijkstra
    graph = synth.SynthGraph(30, 30, seed=4)
    ptr, xy, origin, dest = synth.make_routes(graph, 400, max_hops=15, n_sources=32, seed=9)
    dist = dijkstra(graph.csr(), directed=True, indices=np.unique(origin))
    row = {o: i for i, o in enumerate(np.unique(origin))}
    for k in range(400):
        seg = xy[ptr[k]:ptr[k + 1]]
        assert len(seg) >= 2 and np.array_equal(seg[0], graph.node_xy[origin[k]]) and np.array_equal(seg[-1], graph.node_xy[dest[k]])
        length = np.hypot(*(seg[1:] - seg[:-1]).T).sum()
        assert abs(length - dist[row[origin[k]], dest[k]]) < 1e-6, "route %d is not a shortest path" % k


def test_make_city_world_is_consistent_and_steps():
    from oracle import oracle
    w, graph = synth.make_city(nx_=20, ny_=20, n_cars=500, prescale=4, seed=2, max_hops=10)
    assert w.n_cars == 500 and (w.path_end - w.path_start >= 1).all()
    assert np.array_equal(w.cursor, w.path_start) and w.cell.min() >= 0 and w.cell.max() < w.grid.n_cells
    # staggered starts: every car sits on the segment between its origin node and its first route point
    assert len(np.unique(w.pos, axis=0)) > 450
    oracle.run(w, 1e-3, 50, 1)
    assert np.isfinite(w.pos).all() and (w.cursor >= w.path_start).all()
    w2, _ = synth.make_city(nx_=20, ny_=20, n_cars=500, prescale=4, seed=2, max_hops=10, spatial_order=True)
    assert (np.diff(w2.cell) >= 0).all()
