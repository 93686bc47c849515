"""Init-time routing on the GPU: batched shortest paths + the car table the reference initialisers build.

``shortest_paths`` replaces ``networkx.shortest_path(G, o, d, weight='length')`` for a batch of (origin, destination)
pairs (``navigation.get_route`` / ``get_init_path`` / ``shortest_path_lines_nx``, ``navigation.py:581-609, 801-841``):
the reference runs two networkx Dijkstras per car (~16 ms per car, 32 s for 2,000 cars); here one CTA per distinct
origin runs a label-correcting search on the CSR graph (``csrc/routing.cu``).  ``route_polylines`` also emits, on the
device, the ``xpath`` / ``ypath`` polylines of ``models.path_decompiler`` (``models.py:116-137``) for graphs whose edges
carry no ``geometry`` - the node positions origin..destination with consecutive duplicates dropped - as the compact
path-point CSR the step kernels consume; ``polylines`` is the same rule on the host (numpy), kept as its cross-check and for
node sequences that did not come from the router.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native


def csr_of(n_nodes: int, edge_u, edge_v, edge_len):
    """CSR (indptr, indices, weights) in node-index space; rows keep the edge insertion order."""
    edge_u = np.asarray(edge_u, np.int64)
    order = np.argsort(edge_u, kind="stable")
    indptr = np.zeros(n_nodes + 1, np.int32)
    np.cumsum(np.bincount(edge_u, minlength=n_nodes), out=indptr[1:])
    return indptr, np.asarray(edge_v, np.int32)[order].copy(), np.asarray(edge_len, np.float64)[order].copy()


def shortest_paths(indptr, indices, weights, origin, dest, limit: float | None = None, max_len: int | None = None):
    """Node sequences of the shortest paths.  Returns ``(path_len[int32 n], path_nodes[int32 n, max_len])``;
    ``path_len`` is -1 where no path exists (the reference skips such cars, ``simulation.py:220-222``)."""
    L = native.lib()
    indptr = np.ascontiguousarray(indptr, np.int32)
    indices = np.ascontiguousarray(indices, np.int32)
    weights = np.ascontiguousarray(weights, np.float64)
    origin = np.ascontiguousarray(origin, np.int32)
    dest = np.ascontiguousarray(dest, np.int32)
    n_nodes, n_pairs = len(indptr) - 1, len(origin)
    if max_len is None:
        max_len = n_nodes
    while True:
        path_len = np.empty(n_pairs, np.int32)
        path_nodes = np.empty((n_pairs, max_len), np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = L.tb2_route_paths(n_nodes, p(indptr), p(indices), p(weights), n_pairs, p(origin), p(dest),
                               C.c_double(-1.0 if limit is None else float(limit)), max_len, p(path_len), p(path_nodes), None)
        if rc != 0:
            raise native.NativeError("tb2_route_paths failed (status %d): %s" % (rc, L.tb2_route_last_error().decode()))
        if (path_len == -2).any():
            if max_len < n_nodes:
                max_len = min(n_nodes, max_len * 4)   # a path did not fit: retry with a longer buffer
                continue
            # a simple path never has more than n_nodes nodes: this is a corrupted predecessor walk, not "no path"
            raise native.NativeError("tb2_route_paths: predecessor walk exceeded the number of nodes (%d pairs)"
                                     % int((path_len == -2).sum()))
        return path_len, path_nodes


def route_polylines(indptr, indices, weights, node_xy, origin, dest, limit: float | None = None, max_len: int | None = None):
    """``shortest_paths`` + ``polylines`` in one call, the polylines emitted on the device as the path-point CSR the step
    kernels consume (``tb2_route_polylines``).  Returns ``(path_len, path_nodes, path_ptr[int64 n+1], path_xy[float64 P,2])``."""
    L = native.lib()
    indptr = np.ascontiguousarray(indptr, np.int32)
    indices = np.ascontiguousarray(indices, np.int32)
    weights = np.ascontiguousarray(weights, np.float64)
    node_xy = np.ascontiguousarray(node_xy, np.float64).reshape(-1, 2)
    origin = np.ascontiguousarray(origin, np.int32)
    dest = np.ascontiguousarray(dest, np.int32)
    n_nodes, n_pairs = len(indptr) - 1, len(origin)
    if len(node_xy) != n_nodes:
        raise ValueError("node_xy must hold one position per node")
    if max_len is None:
        max_len = n_nodes
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    while True:
        path_len = np.empty(n_pairs, np.int32)
        path_nodes = np.empty((n_pairs, max_len), np.int32)
        path_ptr = np.zeros(n_pairs + 1, np.int64)
        xy_ptr = C.POINTER(C.c_double)()
        rc = L.tb2_route_polylines(n_nodes, p(indptr), p(indices), p(weights), p(node_xy), n_pairs, p(origin), p(dest),
                                   C.c_double(-1.0 if limit is None else float(limit)), max_len, p(path_len), p(path_nodes),
                                   p(path_ptr), C.byref(xy_ptr), None)
        if rc != 0:
            raise native.NativeError("tb2_route_polylines failed (status %d): %s" % (rc, L.tb2_route_last_error().decode()))
        if n_pairs and (path_len == -2).any():
            if max_len < n_nodes:
                max_len = min(n_nodes, max_len * 4)   # a path did not fit: retry with a longer buffer
                continue
            raise native.NativeError("tb2_route_polylines: predecessor walk exceeded the number of nodes (%d pairs)"
                                     % int((path_len == -2).sum()))
        total = int(path_ptr[-1])
        # the point array belongs to the library until the next routing call: copy it out
        path_xy = (np.ctypeslib.as_array(xy_ptr, shape=(total, 2)).copy() if total else np.zeros((0, 2), np.float64))
        return path_len, path_nodes, path_ptr, path_xy


def polylines(node_xy, path_len, path_nodes):
    """``(path_ptr[int64 n+1], path_xy[float64 P,2])``: positions of the route nodes, consecutive duplicates dropped
    (``models.path_decompiler``); a route of one node (origin == destination) yields an empty polyline, as in the
    reference (no edges -> no lines -> ``xpath == []``)."""
    n = len(path_len)
    lens = np.where(path_len >= 2, path_len, 0).astype(np.int64)
    ar = np.arange(path_nodes.shape[1])[None, :]
    valid = ar < lens[:, None]
    nodes = path_nodes[valid]
    xy = np.asarray(node_xy, np.float64)[nodes]
    owner = np.repeat(np.arange(n), lens)
    keep = np.ones(len(nodes), bool)
    if len(nodes) > 1:
        same_owner = owner[1:] == owner[:-1]
        dup = same_owner & (xy[1:] == xy[:-1]).all(1)
        keep[:-1] &= ~dup     # path_decompiler keeps the LAST of a run of identical points
    kept_len = np.bincount(owner[keep], minlength=n).astype(np.int64)
    ptr = np.zeros(n + 1, np.int64)
    np.cumsum(kept_len, out=ptr[1:])
    return ptr, xy[keep]


def polylines_with_geometry(node_xy, path_len, path_nodes, edge_geom):
    """Like ``polylines`` for graphs whose edges carry an OSMnx ``geometry`` (``navigation.py:823-827``): the polyline of
    edge (u, v) replaces the straight segment.  ``edge_geom`` maps ``(u_index, v_index)`` to an ``(k, 2)`` array (the
    geometry of the shortest parallel edge); edges without an entry contribute their two end points.  Consecutive
    duplicates are dropped as ``models.path_decompiler`` does (an element survives iff it differs from its successor;
    the last one always survives)."""
    node_xy = np.asarray(node_xy, np.float64)
    ptr = [0]
    chunks = []
    for k in range(len(path_len)):
        n = int(path_len[k])
        pts = []
        for a, b in zip(path_nodes[k, :max(n, 0) - 1].tolist() if n >= 2 else [], path_nodes[k, 1:max(n, 0)].tolist() if n >= 2 else []):
            g = edge_geom.get((a, b))
            if g is None:
                pts.append(node_xy[a]); pts.append(node_xy[b])
            else:
                pts.extend(np.asarray(g, np.float64).reshape(-1, 2))
        if pts:
            arr = np.asarray(pts, np.float64).reshape(-1, 2)
            keep = np.ones(len(arr), bool)
            keep[:-1] = (arr[:-1] != arr[1:]).any(1)
            arr = arr[keep]
        else:
            arr = np.zeros((0, 2))
        chunks.append(arr)
        ptr.append(ptr[-1] + len(arr))
    return np.asarray(ptr, np.int64), (np.concatenate(chunks) if chunks else np.zeros((0, 2))).reshape(-1, 2)
