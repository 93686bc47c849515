"""Init-time routing: the GPU batched shortest-path search against the routes the reference itself
produced (golden fixtures: networkx.shortest_path inside sim.init_random_node_start_location) and against scipy."""
import numpy as np
import pytest

from tests import util
from traffic_b200 import routing, synth


def _graph_of(name):
    from oracle.ref_harness import CONFIGS   # test infrastructure; importing it does not touch /root/reference
    return synth.SynthGraph(**CONFIGS[name]["g"])


def test_polylines_drop_duplicates_and_empty_routes():
    node_xy = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 0.0], [2.0, 5.0]])
    path_len = np.array([4, 1, -1, 2], np.int32)
    path_nodes = np.array([[0, 1, 2, 3], [2, 0, 0, 0], [0, 0, 0, 0], [3, 0, 0, 0]], np.int32)
    ptr, xy = routing.polylines(node_xy, path_len, path_nodes)
    assert ptr.tolist() == [0, 3, 3, 3, 5]                       # duplicate point dropped; 1-node and no-path routes empty
    assert xy.tolist() == [[0.0, 0.0], [1.0, 0.0], [2.0, 5.0], [2.0, 5.0], [0.0, 0.0]]


def test_csr_of_keeps_parallel_edges():
    indptr, indices, w = routing.csr_of(3, [0, 0, 1, 0], [1, 2, 2, 1], [5.0, 1.0, 1.0, 2.0])
    assert indptr.tolist() == [0, 3, 4, 4] and indices.tolist() == [1, 2, 1, 2] and w.tolist() == [5.0, 1.0, 2.0, 1.0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_dt0.001_p2_s1", "dense_dt0.001_p2_s3", "c2_dt0.001_p10_s0"])
def test_gpu_routes_equal_reference_routes(name):
    g = util.load_golden(name)
    graph = _graph_of(name)
    # recover (origin, destination) node indices from the fixture's coordinates
    key = {tuple(p): i for i, p in enumerate(graph.node_xy)}
    origin = np.array([key[tuple(p)] for p in g["pos0"]], np.int32)
    dest = np.array([key[tuple(p)] for p in g["dest_xy"]], np.int32)
    indptr, indices, w = routing.csr_of(graph.n_nodes, graph.edge_u, graph.edge_v, graph.edge_len)
    path_len, path_nodes = routing.shortest_paths(indptr, indices, w, origin, dest)
    ptr, xy = routing.polylines(graph.node_xy, path_len, path_nodes)
    assert np.array_equal(ptr, g["path_ptr"]), "polyline lengths differ from the reference's routes"
    assert np.array_equal(xy, g["path_xy"]), "polyline points differ from the reference's routes"
    assert (path_len == 1).sum() == (np.diff(g["path_ptr"]) == 0).sum()   # origin == destination -> empty route


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_dt0.001_p2_s1", "c2_dt0.001_p10_s0"])
def test_device_polylines_equal_reference_routes(name):
    """tb2_route_polylines: routes and their polylines in one call, the points emitted on the device as the path-point CSR -
    identical to the polylines the reference stored for the same cars (and to the host-side rule)."""
    g = util.load_golden(name)
    graph = _graph_of(name)
    key = {tuple(p): i for i, p in enumerate(graph.node_xy)}
    origin = np.array([key[tuple(p)] for p in g["pos0"]], np.int32)
    dest = np.array([key[tuple(p)] for p in g["dest_xy"]], np.int32)
    indptr, indices, w = routing.csr_of(graph.n_nodes, graph.edge_u, graph.edge_v, graph.edge_len)
    path_len, path_nodes, ptr, xy = routing.route_polylines(indptr, indices, w, graph.node_xy, origin, dest)
    assert np.array_equal(ptr, g["path_ptr"]) and np.array_equal(xy, g["path_xy"])
    pl2, pn2 = routing.shortest_paths(indptr, indices, w, origin, dest)
    assert np.array_equal(path_len, pl2)
    ptr2, xy2 = routing.polylines(graph.node_xy, pl2, pn2)
    assert np.array_equal(ptr, ptr2) and np.array_equal(xy, xy2)


@pytest.mark.gpu
def test_device_polylines_drop_duplicate_points_and_handle_degenerate_pairs():
    """Twin nodes (same position) are dropped as models.path_decompiler drops them (the last of a run survives), a pair
    with origin == destination or without a path has no points, a too-small point buffer is retried."""
    graph = synth.SynthGraph(10, 10, seed=4)
    node_xy = np.array(graph.node_xy, np.float64)
    u, v, ln = list(graph.edge_u), list(graph.edge_v), list(graph.edge_len)
    n = graph.n_nodes
    node_xy = np.vstack([node_xy, node_xy[[5, 17, 40]], [[0.0, 0.0]]])       # three twins + one isolated node
    for t, a in enumerate((5, 17, 40)):
        u += [a, n + t]; v += [n + t, a]; ln += [0.0, 0.0]
        for b in [int(x) for x, y in zip(graph.edge_v, graph.edge_u) if y == a][:2]:
            u += [n + t, b]; v += [b, n + t]; ln += [float(np.hypot(*(node_xy[a] - node_xy[b]))) * 0.5] * 2   # a shortcut through the twin
    indptr, indices, w = routing.csr_of(n + 4, u, v, ln)
    rng = np.random.default_rng(1)
    origin = np.r_[rng.integers(0, n + 3, 400), [7, 3]].astype(np.int32)
    dest = np.r_[rng.integers(0, n + 3, 400), [7, n + 3]].astype(np.int32)  # last two: origin == destination, no path
    path_len, path_nodes, ptr, xy = routing.route_polylines(indptr, indices, w, node_xy, origin, dest)
    pl2, pn2 = routing.shortest_paths(indptr, indices, w, origin, dest)
    ptr2, xy2 = routing.polylines(node_xy, pl2, pn2)
    assert np.array_equal(path_len, pl2) and np.array_equal(ptr, ptr2) and np.array_equal(xy, xy2)
    assert path_len[-1] == -1 and ptr[-1] == ptr[-2] == ptr[-3]
    kept = np.diff(ptr)
    assert (kept < np.maximum(path_len, 0)).any(), "test graph: no duplicate point was ever dropped"


@pytest.mark.gpu
def test_gpu_routes_equal_scipy_on_random_graph():
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    rng = np.random.default_rng(5)
    n, m = 3000, 14000
    u, v = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = u != v
    u, v = u[keep], v[keep]
    wts = rng.uniform(1.0, 100.0, len(u))
    indptr, indices, w = routing.csr_of(n, u, v, wts)
    origin = rng.integers(0, 40, 500).astype(np.int32)        # 40 distinct sources
    dest = rng.integers(0, n, 500).astype(np.int32)
    path_len, path_nodes = routing.shortest_paths(indptr, indices, w, origin, dest, max_len=64)
    # scipy keeps the minimum of parallel edges when the matrix is built with duplicates summed -> build min explicitly
    best = {}
    for a, b, c in zip(u, v, wts):
        best[(a, b)] = min(best.get((a, b), np.inf), c)
    rows, cols = zip(*best.keys())
    mat = csr_matrix((list(best.values()), (rows, cols)), shape=(n, n))
    srcs = np.unique(origin)
    dist, pred = dijkstra(mat, directed=True, indices=srcs, return_predecessors=True)
    row_of = {s: i for i, s in enumerate(srcs)}
    for p in range(len(origin)):
        r = row_of[origin[p]]
        if not np.isfinite(dist[r, dest[p]]):
            assert path_len[p] == -1
            continue
        want = [dest[p]]
        while want[-1] != origin[p]:
            want.append(pred[r, want[-1]])
        want = want[::-1]
        assert path_len[p] == len(want) and path_nodes[p, :len(want)].tolist() == [int(x) for x in want]


@pytest.mark.gpu
def test_routing_limit_and_errors():
    from traffic_b200 import native
    graph = synth.SynthGraph(20, 20, seed=1)
    indptr, indices, w = routing.csr_of(graph.n_nodes, graph.edge_u, graph.edge_v, graph.edge_len)
    pl, _ = routing.shortest_paths(indptr, indices, w, [0, 0], [1, 399], limit=500.0)
    assert pl[0] == 2 and pl[1] == -1          # the far corner is outside the search radius
    with pytest.raises(native.NativeError):
        routing.shortest_paths(indptr, indices, w, [0], [400])


def test_polylines_with_edge_geometry():
    """OSMnx edges may carry a `geometry` polyline (navigation.py:823-827): it replaces the straight segment, and
    consecutive duplicates (the shared node between two edges) are dropped like models.path_decompiler does."""
    node_xy = np.array([[0.0, 0.0], [10.0, 0.0], [10.0, 10.0]])
    geom = {(0, 1): np.array([[0.0, 0.0], [4.0, 1.0], [7.0, 1.0], [10.0, 0.0]])}      # edge 1->2 has no geometry
    path_len = np.array([3, 2, 1], np.int32)
    path_nodes = np.array([[0, 1, 2], [1, 2, 0], [2, 0, 0]], np.int32)
    ptr, xy = routing.polylines_with_geometry(node_xy, path_len, path_nodes, geom)
    assert ptr.tolist() == [0, 5, 7, 7]
    assert xy[:5].tolist() == [[0.0, 0.0], [4.0, 1.0], [7.0, 1.0], [10.0, 0.0], [10.0, 10.0]]
    assert xy[5:7].tolist() == [[10.0, 0.0], [10.0, 10.0]]
    # without geometry it must agree with the vectorised version
    ptr2, xy2 = routing.polylines_with_geometry(node_xy, path_len, path_nodes, {})
    ptr3, xy3 = routing.polylines(node_xy, path_len, path_nodes)
    assert np.array_equal(ptr2, ptr3) and np.array_equal(xy2, xy3)


@pytest.mark.gpu
def test_zero_length_edges_and_self_loops_give_valid_routes():
    """OSM twin / overlapping nodes produce zero-length edges (tight in both directions) and self-loops: the predecessor
    walk must still terminate and return a shortest path (same length as networkx's, a valid edge sequence)."""
    import networkx as nx
    graph = synth.SynthGraph(12, 12, seed=2)
    u, v, ln = list(graph.edge_u), list(graph.edge_v), list(graph.edge_len)
    rng = np.random.default_rng(0)
    n = graph.n_nodes
    twins = rng.choice(n - 1, 25, replace=False)
    for a in twins:                       # zero-length edges both ways + a zero-length self-loop + a 1e-13 m edge
        u += [a, a + 1, a]; v += [a + 1, a, a]; ln += [0.0, 0.0, 0.0]
    u += [3, 4]; v += [4, 3]; ln += [1e-13, 1e-13]
    indptr, indices, w = routing.csr_of(n, u, v, ln)
    origin = rng.integers(0, n, 300).astype(np.int32)
    dest = rng.integers(0, n, 300).astype(np.int32)
    path_len, path_nodes = routing.shortest_paths(indptr, indices, w, origin, dest)
    G = nx.MultiDiGraph()
    G.add_nodes_from(range(n))
    for a, b, c in zip(u, v, ln):
        G.add_edge(int(a), int(b), length=float(c))
    wmin = {}
    for a, b, c in zip(u, v, ln):
        wmin[(int(a), int(b))] = min(wmin.get((int(a), int(b)), np.inf), float(c))
    for p in range(len(origin)):
        want = nx.shortest_path_length(G, int(origin[p]), int(dest[p]), weight="length")
        assert path_len[p] >= 1, "valid pair reported as no-path / overflow"
        nodes = path_nodes[p, :path_len[p]].tolist()
        assert nodes[0] == origin[p] and nodes[-1] == dest[p] and len(set(nodes)) == len(nodes)
        got = sum(wmin[(a, b)] for a, b in zip(nodes[:-1], nodes[1:]))
        assert abs(got - want) <= 1e-9 * max(1.0, want)


@pytest.mark.gpu
def test_routing_throughput_on_the_manhattan_grid():
    """2,000 distinct origins on the 4,096-node grid (the reference: two networkx Dijkstras per car,
    ~16 ms per car)."""
    import time
    graph = synth.SynthGraph(64, 64, seed=0)
    indptr, indices, w = routing.csr_of(graph.n_nodes, graph.edge_u, graph.edge_v, graph.edge_len)
    rng = np.random.default_rng(1)
    origin = rng.choice(graph.n_nodes, 2000, replace=False).astype(np.int32)
    dest = rng.integers(0, graph.n_nodes, 2000).astype(np.int32)
    routing.shortest_paths(indptr, indices, w, origin[:8], dest[:8], max_len=256)      # warm up
    t0 = time.perf_counter()
    path_len, _ = routing.shortest_paths(indptr, indices, w, origin, dest, max_len=256)
    el = time.perf_counter() - t0
    assert (path_len >= 1).all()
    print("routing: %d routes over %d distinct origins in %.1f ms = %.0f routes/s" % (len(origin), len(origin), el * 1e3, len(origin) / el))
    assert len(origin) / el > 50000
