set -x
mkdir -p gpurun_out
python -m pytest tests/test_routing.py -m gpu -x -q 2>&1 | tail -8
python - <<'PY'
import time, numpy as np, json
from traffic_b200 import routing, synth
out = {}
for name, nx_, n_pairs, limit in (("c2 64x64, 2000 cars", 64, 2000, None), ("c3 100x100, 50k cars/2048 origins", 100, 50000, None), ("c4 316x316, 1M pairs / 4096 origins, limit 7.5 km", 316, 1000000, 7500.0)):
    g = synth.SynthGraph(nx_, nx_, seed=0)
    indptr, indices, w = routing.csr_of(g.n_nodes, g.edge_u, g.edge_v, g.edge_len)
    rng = np.random.default_rng(0)
    n_src = min(n_pairs, 2000 if nx_ == 64 else (2048 if nx_ == 100 else 4096))
    srcs = rng.choice(g.n_nodes, n_src, replace=False)
    origin = srcs[rng.integers(0, n_src, n_pairs)].astype(np.int32)
    if limit is None:
        dest = rng.integers(0, g.n_nodes, n_pairs).astype(np.int32)
    else:  # destinations within ~40 hops
        r, c = origin // nx_, origin % nx_
        dest = (np.clip(r + rng.integers(-20, 21, n_pairs), 0, nx_ - 1) * nx_ + np.clip(c + rng.integers(-20, 21, n_pairs), 0, nx_ - 1)).astype(np.int32)
    routing.shortest_paths(indptr, indices, w, origin[:8], dest[:8], limit=limit, max_len=256)   # context + workspace
    t0 = time.perf_counter()
    pl, pn = routing.shortest_paths(indptr, indices, w, origin, dest, limit=limit, max_len=256)
    el = time.perf_counter() - t0
    t0 = time.perf_counter()
    ptr_h, xy_h = routing.polylines(g.node_xy, pl, pn)                      # host rule (numpy)
    el_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    pl2, pn2, ptr_d, xy_d = routing.route_polylines(indptr, indices, w, g.node_xy, origin, dest, limit=limit, max_len=256)
    el_dev = time.perf_counter() - t0
    assert np.array_equal(ptr_h, ptr_d) and np.array_equal(xy_h, xy_d)
    out[name] = {"pairs": n_pairs, "origins": int(n_src), "seconds": el, "routes_per_s": n_pairs / el, "no_path": int((pl < 0).sum()), "mean_len": float(pl[pl > 0].mean()),
                 "routes_plus_host_polylines_s": el + el_host, "routes_plus_device_polylines_s": el_dev, "polyline_points": int(ptr_d[-1])}
    print(name, out[name], flush=True)
json.dump(out, open("gpurun_out/routing_bench.json", "w"), indent=1)
PY
