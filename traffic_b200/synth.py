"""Synthetic road worlds for the vehicle-update path (no network, no osmnx).

The reference gets its map from OSMnx (``osm_request.py:5-67``): an object with ``.G`` (a networkx
MultiDiGraph whose nodes carry projected ``x``/``y`` and whose edges carry ``length``), ``.axis``
(xmin, xmax, ymin, ymax) and matplotlib ``.fig``/``.ax``.  ``SynthGraph`` is a duck-typed stand-in with the
same attributes: a rotated, jittered grid at UTM-like coordinates.  The rotation keeps every
edge off the coordinate axes (an axis-aligned edge has ``int(|dx|) == 0`` -> empty linspace -> the reference
never detects anything on it, ``models.py:168`` / ``navigation.py:437``) and the jitter removes collinear
path triples (``arccos(1 - ulp)`` razor edge).

Everything here is host-side numpy; the arrays feed ``traffic_b200.state`` (the SoA packer).  networkx is only
imported when ``.G`` is touched (the reference-facing facade and the golden-vector generator need it).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

BIN_SIZE = 200.0  # models.py:16


class SynthGraph:
    """Rotated + jittered ``nx_ x ny_`` grid, bidirectional edges, Euclidean ``length``.

    Node ids are ``node_id_base + row * nx_ + col`` (row-major), so ``list(G.nodes())[:n]`` - which is what
    ``navigation.find_nodes`` (``navigation.py:552-563``) hands to the reference initialiser - is the first n
    nodes in that order.
    """

    def __init__(self, nx_: int, ny_: int, spacing: float = 100.0, x0: float = 566000.0, y0: float = 4186000.0,
                 theta_deg: float = 29.0, jitter: float = 5.0, seed: int = 0, node_id_base: int = 1000,
                 culdesacs: int = 0):
        self.nx, self.ny = int(nx_), int(ny_)
        self.spacing = float(spacing)
        self.node_id_base = int(node_id_base)
        rng = np.random.default_rng(seed)
        cols, rows = np.meshgrid(np.arange(self.nx, dtype=np.float64), np.arange(self.ny, dtype=np.float64))
        gx = cols.ravel() * spacing
        gy = rows.ravel() * spacing
        th = math.radians(theta_deg)
        # rotate about the grid centre so the bounding box stays roughly square
        cx, cy = gx.mean(), gy.mean()
        rx = (gx - cx) * math.cos(th) - (gy - cy) * math.sin(th)
        ry = (gx - cx) * math.sin(th) + (gy - cy) * math.cos(th)
        rx += rng.uniform(-jitter, jitter, rx.shape)
        ry += rng.uniform(-jitter, jitter, ry.shape)
        rx -= rx.min()
        ry -= ry.min()
        self.node_xy = np.stack([x0 + rx, y0 + ry], axis=1)  # (n, 2) float64
        n = self.nx * self.ny
        idx = np.arange(n).reshape(self.ny, self.nx)
        right = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1)
        up = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)
        und = np.concatenate([right, up], 0)
        self.n_culdesacs = int(culdesacs)
        if culdesacs:   # dead-end stubs (one street each, navigation.find_culdesacs) hanging off random grid nodes
            host = rng.choice(n, culdesacs, replace=False)
            ang = rng.uniform(0, 2 * math.pi, culdesacs)
            stub = self.node_xy[host] + 0.45 * spacing * np.stack([np.cos(ang), np.sin(ang)], 1)
            self.node_xy = np.concatenate([self.node_xy, stub], 0)
            und = np.concatenate([und, np.stack([host, n + np.arange(culdesacs)], 1)], 0)
        self.node_ids = self.node_id_base + np.arange(len(self.node_xy), dtype=np.int64)
        # directed both ways; order: for each undirected edge (u, v) then (v, u)
        self.edge_u = np.concatenate([und[:, 0], und[:, 1]])
        self.edge_v = np.concatenate([und[:, 1], und[:, 0]])
        d = self.node_xy[self.edge_v] - self.node_xy[self.edge_u]
        self.edge_len = np.hypot(d[:, 0], d[:, 1])
        xmin, ymin = self.node_xy.min(0)
        xmax, ymax = self.node_xy.max(0)
        w, h = xmax - xmin, ymax - ymin
        self.axis = (xmin - 0.02 * w, xmax + 0.02 * w, ymin - 0.02 * h, ymax + 0.02 * h)
        self.fig = None
        self.ax = SimpleNamespace(axis=lambda: self.axis)  # environment.py:24,29 reads graph.ax.axis()
        self._G = None

    @property
    def n_nodes(self) -> int:
        return len(self.node_xy)

    @property
    def G(self):
        """networkx MultiDiGraph view (built lazily; nodes inserted in row-major order)."""
        if self._G is None:
            import networkx as nx

            G = nx.MultiDiGraph()
            for i, nid in enumerate(self.node_ids):
                G.add_node(int(nid), x=float(self.node_xy[i, 0]), y=float(self.node_xy[i, 1]))
            for u, v, ln in zip(self.edge_u, self.edge_v, self.edge_len):
                G.add_edge(int(self.node_ids[u]), int(self.node_ids[v]), length=float(ln))
            self._G = G
        return self._G

    def csr(self):
        """(indptr, indices, weights) of the directed graph, node *indices* (not ids)."""
        import scipy.sparse as sp

        n = self.n_nodes
        m = sp.csr_matrix((self.edge_len, (self.edge_u, self.edge_v)), shape=(n, n))
        return m


def grid_dims(axis, bin_size: float = BIN_SIZE):
    """Number of bin edges per axis, exactly as ``np.arange(axis[0], axis[1], 200)`` (``models.py:16``)."""
    nbx = len(np.arange(axis[0], axis[1], bin_size))
    nby = len(np.arange(axis[2], axis[3], bin_size))
    return nbx, nby


# ----------------------------------------------------------------------------------------------------------------
# Large-world generators (configs c3/c4/c5 of SURVEY.md §8(d)): same schema as the reference initialisers
# (``simulation.py:189-256, 329-381``) but vectorised - the reference's one-car-per-node, two-networkx-Dijkstras-
# per-car initialiser cannot make 50k cars on 10k nodes.
# ----------------------------------------------------------------------------------------------------------------

def make_routes(graph: SynthGraph, n_cars: int, max_hops: int = 40, n_sources: int = 256, seed: int = 0,
                max_points: int = 64, reach_hops: int | None = None):
    """True shortest paths (by ``length``) for ``n_cars`` cars, destinations uniform over the map.

    A bounded set of Dijkstra trees (``n_sources`` roots, search radius ``reach_hops``) is shared by all cars: every
    sub-path of a shortest path is a shortest path, so a car's route is the last ``l`` hops (``2 <= l <= max_hops``)
    of the tree path root -> destination, for a root at least ``l`` hops away.  Origins therefore spread over the whole
    map instead of piling up on the roots.

    Returns ``(path_ptr[int64 N+1], path_xy[float64 P,2], origin_idx, dest_idx)``.  A car's polyline is the node
    positions origin..destination - what ``navigation.shortest_path_lines_nx`` + ``models.path_decompiler``
    (``navigation.py:801-841``, ``models.py:116-137``) produce on a graph without edge ``geometry``.
    """
    from scipy.sparse.csgraph import dijkstra

    rng = np.random.default_rng(seed)
    n = graph.n_nodes
    m = graph.csr()
    n_sources = int(min(n_sources, n))
    max_l = int(min(max_hops, max_points - 1, graph.nx + graph.ny - 2))
    if reach_hops is None:
        reach_hops = max_l + 40
    sources = rng.choice(n, size=n_sources, replace=False)
    limit = reach_hops * graph.spacing * 1.25
    _, pred = dijkstra(m, directed=True, indices=sources, return_predecessors=True, limit=limit)
    pred = pred.astype(np.int32)
    srow, scol = sources // graph.nx, sources % graph.nx

    dest = rng.integers(0, n, size=n_cars)
    hops = rng.integers(2, max_l + 1, size=n_cars)
    drow, dcol = dest // graph.nx, dest % graph.nx
    src_of_car = np.full(n_cars, -1, np.int64)
    todo = np.arange(n_cars)
    for rnd in range(400):
        if len(todo) == 0:
            break
        if rnd and rnd % 25 == 0:  # unlucky destinations (map corners far from every root): redraw them
            dest[todo] = rng.integers(0, n, size=len(todo))
            hops[todo] = rng.integers(2, max(3, max_l // (1 + rnd // 100)) + 1, size=len(todo))
            drow, dcol = dest // graph.nx, dest % graph.nx
        cand = rng.integers(0, n_sources, size=len(todo))
        hd = np.abs(srow[cand] - drow[todo]) + np.abs(scol[cand] - dcol[todo])
        ok = (hd >= hops[todo]) & (hd <= reach_hops - 8) & (pred[cand, dest[todo]] >= 0)
        src_of_car[todo[ok]] = cand[ok]
        todo = todo[~ok]
    if len(todo):
        raise RuntimeError("could not route %d cars; use more sources or a larger reach" % len(todo))
    # walk `hops` predecessors from the destination towards the root (vectorised over cars)
    chain = np.full((n_cars, max_l + 1), -1, np.int32)
    cur = dest.astype(np.int32)
    chain[:, 0] = cur
    for k in range(1, max_l + 1):
        alive = (hops >= k) & (cur != sources[src_of_car])
        nxt = pred[src_of_car, cur]
        cur = np.where(alive, nxt, cur).astype(np.int32)
        chain[:, k] = np.where(alive, cur, -1)
    length = (chain >= 0).sum(1)                      # points origin..dest
    ar = np.arange(max_l + 1)[None, :]
    gather = length[:, None] - 1 - ar
    valid = gather >= 0
    fwd = np.take_along_axis(chain, np.where(valid, gather, 0), 1)
    path_ptr = np.zeros(n_cars + 1, np.int64)
    np.cumsum(length, out=path_ptr[1:])
    path_xy = graph.node_xy[fwd[valid]]
    origin = fwd[:, 0].astype(np.int64)
    return path_ptr, path_xy, origin, dest


def make_lights(graph: SynthGraph, prescale: int = 10, seed: int = 0):
    """Vectorised restatement of ``simulation.init_traffic_lights`` (``simulation.py:329-381``) on a grid graph.

    Lights sit on every ``prescale``-th node in node order (every grid node has MultiDiGraph degree > 3,
    ``navigation.py:545-547``); a light's faces are the vectors to its out-neighbours in adjacency insertion order
    (``navigation.determine_pedigree``, ``navigation.py:501-521`` - on a jittered grid the shortest path to an
    out-neighbour is the direct edge), ``switch-time = round(U(0,1) * degree, 2)`` (``models.py:31-38``) and the
    initial ``go`` pattern is ``[False, True, False, ...]`` (``simulation.py:355-356``).
    """
    rng = np.random.default_rng(seed)
    n = graph.n_nodes
    light_nodes = np.arange(0, n, prescale)
    # adjacency insertion order of SynthGraph.G: edges are added (right u->v) for all, (up u->v) for all,
    # then the reverses; out-neighbour order per node = right, up, left, down (those that exist)
    r, c = light_nodes // graph.nx, light_nodes % graph.nx
    cand = np.stack([np.where(c + 1 < graph.nx, light_nodes + 1, -1),
                     np.where(r + 1 < graph.ny, light_nodes + graph.nx, -1),
                     np.where(c - 1 >= 0, light_nodes - 1, -1),
                     np.where(r - 1 >= 0, light_nodes - graph.nx, -1)], 1)
    has = cand >= 0
    degree = has.sum(1)
    face_ptr = np.zeros(len(light_nodes) + 1, np.int32)
    np.cumsum(degree, out=face_ptr[1:])
    owner = np.repeat(np.arange(len(light_nodes)), 4).reshape(-1, 4)[has]
    nb = cand[has]
    face_vec = graph.node_xy[nb] - graph.node_xy[light_nodes[owner]]
    k_in_light = np.arange(len(nb)) - face_ptr[owner]
    face_go = (k_in_light % 2 == 1)
    switch_time = np.round(rng.random(len(light_nodes)) * degree, 2)
    return dict(node=graph.node_ids[light_nodes], xy=graph.node_xy[light_nodes].copy(), degree=degree.astype(np.int32),
                switch_time=switch_time, face_ptr=face_ptr, face_vec=face_vec, face_go=face_go)


def make_city(nx_: int, ny_: int, n_cars: int, prescale: int = 10, seed: int = 0, x0: float = 551000.0,
              y0: float = 4180000.0, max_hops: int = 40, max_points: int = 64, n_sources: int = 256,
              stagger: bool = True, spatial_order: bool = False):
    """A full synthetic ``World`` (graph + routed cars + lights) of arbitrary size, same semantics as a world packed
    from the reference's DataFrames.  With ``stagger`` each car starts at a random point of its first edge and its
    polyline starts at that edge's far end, so co-located starts with identical routes (mutually invisible forever,
    SURVEY Appendix B-5) do not occur.  Returns ``(world, graph)``.
    """
    from .state import make_world

    g = SynthGraph(nx_, ny_, x0=x0, y0=y0, seed=seed)
    path_ptr, path_xy, origin, dest = make_routes(g, n_cars, max_hops=max_hops, n_sources=n_sources, seed=seed + 1,
                                                  max_points=max_points)
    rng = np.random.default_rng(seed + 2)
    start = path_ptr[:-1]
    if stagger:
        p0 = path_xy[start]
        p1 = path_xy[start + 1]          # every route has >= 2 points (destination != origin)
        frac = rng.uniform(0.0, 0.9, n_cars)[:, None]
        pos = p0 + frac * (p1 - p0)
        keep = np.ones(len(path_xy), bool)
        keep[start] = False
        path_xy = path_xy[keep]
        path_ptr = path_ptr - np.arange(n_cars + 1)
    else:
        pos = path_xy[start].copy()
    lights = make_lights(g, prescale=prescale, seed=seed + 3)
    dest_xy = g.node_xy[dest]
    if spatial_order:  # experiment knob: number the cars in bin order of their start positions
        from .state import Grid, digitize_cells
        order = np.argsort(digitize_cells(Grid.from_axis(g.axis), pos), kind="stable")
        lens = np.diff(path_ptr)[order]
        starts = path_ptr[:-1][order]
        idx = np.concatenate([np.arange(a, a + n) for a, n in zip(starts, lens)]) if len(order) else np.zeros(0, np.int64)
        path_xy = path_xy[idx]
        path_ptr = np.concatenate([[0], np.cumsum(lens)])
        pos, dest_xy = pos[order], dest_xy[order]
    w = make_world(g.axis, pos, path_ptr, path_xy, dest_xy, lights)
    return w, g


CITY_CONFIGS = {
    # BASELINE.json configs[1..3]; sizes of the synthetic grid, cars, light prescale, UTM-like offset
    "c2_manhattan_2k": dict(nx_=64, ny_=64, n_cars=2000, prescale=10, x0=583000.0, y0=4507000.0, max_hops=40),
    "c3_sf_50k": dict(nx_=100, ny_=100, n_cars=50000, prescale=10, x0=551000.0, y0=4180000.0, max_hops=40),
    "c4_city_1m": dict(nx_=316, ny_=316, n_cars=1000000, prescale=10, x0=551000.0, y0=4180000.0, max_hops=60),
}
