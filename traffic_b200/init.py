"""Drop-in replacements of the reference's initialisers, with the per-car / per-face networkx Dijkstras replaced by ONE
batched GPU search (``traffic_b200.routing``).

* ``init_random_node_start_location(n, graph)``  <->  ``simulation.init_random_node_start_location``
  (``simulation.py:189-256``): n-1 cars, car i starts on the i-th node, destination ``nodes[round(random.random()*n)]``.
* ``init_traffic_lights(graph, prescale)``        <->  ``simulation.init_traffic_lights`` (``simulation.py:329-381``).

Same DataFrame schemas, same consumption order of Python's ``random`` stream, so
``random.seed(s)`` followed by these two calls yields the frames the reference yields (tests/test_init.py checks them
against the inputs recorded in the golden fixtures).  The reference spends ~16 ms per car here (32 s for 2,000 cars);
the batched search routes ~1e6 pairs per second.
"""
from __future__ import annotations

import random

import numpy as np

from . import routing


def _graph_arrays(graph):
    """Node order, positions and a CSR over node indices from an OGraph-like object (``graph.G`` networkx MultiDiGraph
    with node ``x``/``y`` and edge ``length``) or a ``SynthGraph`` (arrays already there)."""
    if hasattr(graph, "node_xy") and hasattr(graph, "edge_u"):
        ids = np.asarray(graph.node_ids)
        return ids, np.asarray(graph.node_xy, np.float64), routing.csr_of(graph.n_nodes, graph.edge_u, graph.edge_v, graph.edge_len), None, {}
    G = graph.G
    ids = np.asarray(list(G.nodes()))
    index = {nid: i for i, nid in enumerate(ids.tolist())}
    xy = np.asarray([[G.nodes[n]["x"], G.nodes[n]["y"]] for n in ids.tolist()], np.float64)
    eu, ev, el = [], [], []
    best, geom = {}, {}
    for u, v, data in G.edges(data=True):
        a, b, ln = index[u], index[v], float(data["length"])
        eu.append(a); ev.append(b); el.append(ln)
        # the shortest of parallel edges supplies the line geometry (navigation.py:820-827)
        if (a, b) not in best or ln < best[(a, b)]:
            best[(a, b)] = ln
            if "geometry" in data:
                gx, gy = data["geometry"].xy
                geom[(a, b)] = np.stack([np.asarray(gx, np.float64), np.asarray(gy, np.float64)], 1)
            else:
                geom.pop((a, b), None)
    return ids, xy, routing.csr_of(len(ids), eu, ev, el), G, geom


def _adjacency_order(graph, ids):
    """Out-neighbours of every node in the order ``graph.G[node]`` iterates them (``navigation.py:511``)."""
    if hasattr(graph, "edge_u"):
        n = len(ids)
        seen = [dict() for _ in range(n)]
        for u, v in zip(np.asarray(graph.edge_u).tolist(), np.asarray(graph.edge_v).tolist()):
            seen[u].setdefault(v, None)
        return [list(d.keys()) for d in seen]
    index = {nid: i for i, nid in enumerate(ids.tolist())}
    return [[index[v] for v in graph.G[nid]] for nid in ids.tolist()]


def _routes_and_polylines(indptr, indices, w, xy, origins, dests, geom):
    """Shortest routes and their polylines: one device call for plain graphs (tb2_route_polylines); graphs whose edges
    carry an OSMnx geometry get their polylines assembled from the edge geometries on the host."""
    if geom:
        path_len, path_nodes = routing.shortest_paths(indptr, indices, w, origins, dests)
        ptr, pxy = routing.polylines_with_geometry(xy, path_len, path_nodes, geom)
        return path_len, path_nodes, ptr, pxy
    return routing.route_polylines(indptr, indices, w, xy, origins, dests)


def init_random_node_start_location(n, graph, car_id=None, alternate_route=None):
    import pandas as pd

    ids, xy, (indptr, indices, w), _, geom = _graph_arrays(graph)
    nodes = np.arange(len(ids))[:n]                      # nav.find_nodes: the first n nodes
    origins, dests = [], []
    for i in range(n):
        if i < n - 1:
            random_index = round(random.random() * n)    # simulation.py:214
            origins.append(nodes[i])
            dests.append(nodes[random_index] if random_index != n else nodes[0])
    origins, dests = np.asarray(origins, np.int32), np.asarray(dests, np.int32)
    path_len, path_nodes, ptr, pxy = _routes_and_polylines(indptr, indices, w, xy, origins, dests, geom)
    cars_data = []
    for k in range(len(origins)):
        if path_len[k] < 0:                              # NetworkXNoPath: the reference skips the car
            print("No path between {} and {}.".format(ids[origins[k]], ids[dests[k]]))
            continue
        seg = pxy[ptr[k]:ptr[k + 1]]
        cars_data.append({"object": "car", "x": xy[origins[k], 0], "y": xy[origins[k], 1], "vx": 0, "vy": 0,
                          "route-time": 0, "origin": ids[origins[k]].item(), "destination": ids[dests[k]].item(),
                          "route": [ids[j].item() for j in path_nodes[k, :path_len[k]]],
                          "xpath": seg[:, 0].tolist(), "ypath": seg[:, 1].tolist(),
                          "distance-to-car": 0, "distance-to-node": 0, "distance-to-red-light": 0})
    if alternate_route:
        cars_data[car_id]["route"], cars_data[car_id]["xpath"], cars_data[car_id]["ypath"] = alternate_route
    cars = pd.DataFrame(cars_data)
    axis = graph.axis
    xbins, ybins = np.arange(axis[0], axis[1], 200), np.arange(axis[2], axis[3], 200)
    cars["xbin"], cars["ybin"] = pd.Series(np.digitize(cars["x"], xbins)), pd.Series(np.digitize(cars["y"], ybins))
    print("Number of cars: {}".format(len(cars)))
    return cars


def find_culdesacs(graph):
    """``navigation.find_culdesacs`` (``navigation.py:522-531``): nodes with exactly one street, in node order.  The
    reference asks ``osmnx.stats.count_streets_per_node``; on a graph without OSMnx that is the number of distinct
    neighbours of the node in the undirected sense."""
    ids, _, (indptr, indices, _w), _, _ = _graph_arrays(graph)
    n = len(ids)
    nb = [set() for _ in range(n)]
    for u in range(n):
        for v in indices[indptr[u]:indptr[u + 1]].tolist():
            if v != u:
                nb[u].add(v)
                nb[v].add(u)
    return [i for i in range(n) if len(nb[i]) == 1]


def init_culdesac_start_location(n, graph, car_id=None, alternate_route=None):
    """``simulation.init_culdesac_start_location`` (``simulation.py:259-326``): car i starts in the i-th cul-de-sac and
    drives to the (i+1)-th.  Same schema, same ValueError / IndexError behaviour (the reference indexes culdesacs[i + 1],
    so n == len(culdesacs) raises IndexError there too); all routes come from one batched GPU search."""
    import pandas as pd

    ids, xy, (indptr, indices, w), _, geom = _graph_arrays(graph)
    culdesacs = find_culdesacs(graph)
    if n > len(culdesacs):
        raise ValueError('Number of cars greater than culdesacs to place them. '
                         'Choose a number less than {}'.format(len(culdesacs)))
    origins = np.asarray([culdesacs[i] for i in range(n)], np.int32)
    dests = np.asarray([culdesacs[i + 1] for i in range(n)], np.int32)     # IndexError when n == len(culdesacs), as upstream
    path_len, path_nodes, ptr, pxy = _routes_and_polylines(indptr, indices, w, xy, origins, dests, geom)
    cars_data = []
    for k in range(n):
        if path_len[k] < 0:
            print('No path between {} and {}.'.format(ids[origins[k]], ids[dests[k]]))
            continue
        seg = pxy[ptr[k]:ptr[k + 1]]
        cars_data.append({'object': 'car', 'x': xy[origins[k], 0], 'y': xy[origins[k], 1], 'vx': 0, 'vy': 0,
                          'route-time': 0, 'origin': ids[origins[k]].item(), 'destination': ids[dests[k]].item(),
                          'route': [ids[j].item() for j in path_nodes[k, :path_len[k]]],
                          'xpath': seg[:, 0].tolist(), 'ypath': seg[:, 1].tolist(),
                          'distance-to-car': 0, 'distance-to-node': 0, 'distance-to-red-light': 0})
    if alternate_route:
        cars_data[car_id]['route'], cars_data[car_id]['xpath'], cars_data[car_id]['ypath'] = alternate_route
    cars = pd.DataFrame(cars_data)
    axis = graph.axis
    xbins, ybins = np.arange(axis[0], axis[1], 200), np.arange(axis[2], axis[3], 200)
    cars['xbin'], cars['ybin'] = pd.Series(np.digitize(cars['x'], xbins)), pd.Series(np.digitize(cars['y'], ybins))
    return cars


def init_traffic_lights(graph, prescale=10):
    import pandas as pd

    epsilon = 0.3                                         # simulation.py:337
    ids, xy, (indptr, indices, w), G, geom = _graph_arrays(graph)
    adj = _adjacency_order(graph, ids)
    if G is not None:
        degree = [d for _, d in G.degree()]
    else:                                                 # MultiDiGraph degree = in + out edges
        degree = (np.bincount(graph.edge_u, minlength=len(ids)) + np.bincount(graph.edge_v, minlength=len(ids))).tolist()
    light_nodes = [i for i in range(len(ids)) if degree[i] > 3 and not (i % prescale)]   # navigation.py:545-547
    # determine_pedigree: for every out-neighbour the 2nd point of the shortest path towards it (navigation.py:501-521)
    o = np.asarray([i for i in light_nodes for _ in adj[i]], np.int32)
    d = np.asarray([v for i in light_nodes for v in adj[i]], np.int32)
    path_len, path_nodes = routing.shortest_paths(indptr, indices, w, o, d, max_len=64) if len(o) else (np.zeros(0, np.int32), np.zeros((0, 1), np.int32))
    lights_data, k = [], 0
    for i in light_nodes:
        vectors = []
        for _ in adj[i]:
            if path_len[k] >= 2:
                first_hop = path_nodes[k, 1]
                g = geom.get((int(path_nodes[k, 0]), int(first_hop)))
                second = g[1] if g is not None and len(g) > 1 else xy[first_hop]   # lines_to_node(...)[0][1]
                vectors.append((second[0] - xy[i, 0], second[1] - xy[i, 1]))
            k += 1
        out_vectors = np.array(vectors)
        deg = len(out_vectors)
        go = ([False, True] * deg * 2)[:deg]
        light = {"object": "light", "node": ids[i].item(), "degree": deg, "x": xy[i, 0], "y": xy[i, 1], "switch-counter": 0,
                 "switch-time": round(random.random() * deg, 2)}                       # models.py:31-38
        light["out-xpositions"] = [xy[i, 0] + epsilon * out_vectors[j][0] for j in range(deg)]
        light["out-ypositions"] = [xy[i, 1] + epsilon * out_vectors[j][1] for j in range(deg)]
        light["out-xvectors"] = [out_vectors[j][0] for j in range(deg)]
        light["out-yvectors"] = [out_vectors[j][1] for j in range(deg)]
        light["go-values"] = np.array([go[j] for j in range(deg)])
        lights_data.append(light)
    lights = pd.DataFrame(lights_data)
    xbins, ybins = np.arange(graph.axis[0], graph.axis[1], 200), np.arange(graph.axis[2], graph.axis[3], 200)
    lights["xbin"], lights["ybin"] = pd.Series(np.digitize(lights["x"], xbins)), pd.Series(np.digitize(lights["y"], ybins))
    return lights
