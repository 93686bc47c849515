// Batched shortest-path routing on the device - the init-time producer of the hot path's inputs.
//
// Replaces, for a batch of (origin, destination) pairs, networkx.shortest_path(G, o, d, weight='length') as the
// reference calls it once per car and once per light face (navigation.get_route / get_init_path /
// shortest_path_lines_nx / lines_to_node, navigation.py:581-609, 764-841; simulation.py:217-219; about 16 ms per car in
// the reference).  One CTA per distinct origin runs a label-correcting (Bellman-Ford style) single-source search on the
// CSR graph with 64-bit atomicMin on the distance bit patterns (non-negative doubles order like their bit patterns), then
// a predecessor pass and a per-pair walk emit the node sequences.  Distances satisfy the same recurrence Dijkstra
// evaluates, dist[v] = min_u fl(dist[u] + w(u, v)), so the chosen predecessors - and hence the routes - are the ones
// networkx returns whenever the shortest path is unique (ties are measure-zero on real-valued edge lengths; the parity
// tests compare against the routes the reference itself produced).
#include "../../include/traffic_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr unsigned long long kInfBits = 0x7ff0000000000000ull;  // +inf

// One CTA per distinct origin.  SMEM = true: the distance labels, hop labels and the frontier flags of the whole graph
// live in shared memory (graphs up to ~14k nodes) and are relaxed with shared-memory atomics; otherwise they live in
// global memory.  Frontier driven: a sweep relaxes only the out-edges of nodes whose label changed in the previous sweep,
// so the work is O(edges x a small factor) instead of O(nodes x sweeps).
//
// Predecessors are made acyclic by construction: after the distances have converged, hop labels are relaxed over the
// TIGHT edges only (fl(dist[u] + w) == dist[v]) and pred[v] is the smallest u on a tight edge with hop[u] + 1 == hop[v].
// A zero-length edge (OSM twin nodes) or a tiny length lost to rounding is tight in both directions, but hop labels
// strictly decrease towards the origin, so the walk back always terminates (the predecessor graph is a tree).
template <bool SMEM>
__global__ void __launch_bounds__(256) sssp_kernel(int n_nodes, const int32_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices, const double* __restrict__ weight,
                                                   const int32_t* __restrict__ sources, double limit,
                                                   unsigned long long* dist_all, int32_t* pred_all, int32_t* hop_all,
                                                   uint8_t* flag_all) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ int changed;
    const int s = blockIdx.x;
    unsigned long long* dist_g = dist_all + (size_t)s * n_nodes;
    int32_t* pred = pred_all + (size_t)s * n_nodes;
    unsigned long long* dist = SMEM ? reinterpret_cast<unsigned long long*>(sm_raw) : dist_g;
    int32_t* hop = SMEM ? reinterpret_cast<int32_t*>(sm_raw + (size_t)n_nodes * 8) : hop_all + (size_t)s * n_nodes;
    uint8_t* flag0 = SMEM ? sm_raw + (size_t)n_nodes * 12 : flag_all + (size_t)s * 2 * n_nodes;
    uint8_t* flag1 = flag0 + n_nodes;
    const int src = sources[s];
    for (int u = threadIdx.x; u < n_nodes; u += 256) {
        dist[u] = (u == src) ? 0ull : kInfBits;
        hop[u] = (u == src) ? 0 : 0x7fffffff;
        pred[u] = 0x7fffffff;
        flag0[u] = (u == src);
        flag1[u] = 0;
    }
    __syncthreads();
    // ---- distances: frontier-driven label correcting ----
    uint8_t *cur = flag0, *nxt = flag1;
    for (int sweep = 0; sweep < n_nodes + 1; ++sweep) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        int my_changed = 0;
        for (int u = threadIdx.x; u < n_nodes; u += 256) {
            if (!cur[u]) continue;
            cur[u] = 0;
            const unsigned long long dub = *(volatile unsigned long long*)&dist[u];
            const double du = __longlong_as_double((long long)dub);
            for (int e = indptr[u]; e < indptr[u + 1]; ++e) {
                const double nd = __dadd_rn(du, weight[e]);
                if (!(nd <= limit)) continue;
                const unsigned long long nb = (unsigned long long)__double_as_longlong(nd);
                const int v = indices[e];
                if (nb < *(volatile unsigned long long*)&dist[v]) {
                    if (atomicMin(&dist[v], nb) > nb) { nxt[v] = 1; my_changed = 1; }
                }
            }
        }
        if (my_changed) changed = 1;
        __syncthreads();
        if (!changed) break;
        uint8_t* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    // ---- hop labels over the tight edges (acyclic predecessor choice) ----
    for (int sweep = 0; sweep < n_nodes + 1; ++sweep) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        int my_changed = 0;
        for (int u = threadIdx.x; u < n_nodes; u += 256) {
            const int hu = *(volatile int32_t*)&hop[u];
            if (hu == 0x7fffffff) continue;
            const double du = __longlong_as_double((long long)dist[u]);
            for (int e = indptr[u]; e < indptr[u + 1]; ++e) {
                const int v = indices[e];
                if (v == u) continue;
                const double nd = __dadd_rn(du, weight[e]);
                if ((unsigned long long)__double_as_longlong(nd) != dist[v]) continue;   // not tight
                if (hu + 1 < *(volatile int32_t*)&hop[v]) {
                    if (atomicMin(&hop[v], hu + 1) > hu + 1) my_changed = 1;
                }
            }
        }
        if (my_changed) changed = 1;
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    // predecessor of v: the smallest u on a tight edge one hop closer to the origin
    for (int u = threadIdx.x; u < n_nodes; u += 256) {
        const int hu = hop[u];
        if (hu == 0x7fffffff) continue;
        const double du = __longlong_as_double((long long)dist[u]);
        for (int e = indptr[u]; e < indptr[u + 1]; ++e) {
            const int v = indices[e];
            if (v == src || v == u) continue;
            const double nd = __dadd_rn(du, weight[e]);
            if ((unsigned long long)__double_as_longlong(nd) == dist[v] && hu + 1 == hop[v]) atomicMin(&pred[v], u);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int u = threadIdx.x; u < n_nodes; u += 256) dist_g[u] = dist[u];
    }
}

// one thread per (origin, destination) pair: walk the predecessors back from the destination, then reverse
__global__ void __launch_bounds__(128) extract_paths_kernel(int n_pairs, int n_nodes, const int32_t* __restrict__ pair_source_slot,
                                                            const int32_t* __restrict__ origin, const int32_t* __restrict__ dest,
                                                            const unsigned long long* __restrict__ dist_all,
                                                            const int32_t* __restrict__ pred_all, int max_len,
                                                            int32_t* __restrict__ path_len, int32_t* __restrict__ path_nodes) {
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_pairs) return;
    const size_t base = (size_t)pair_source_slot[p] * n_nodes;
    int32_t* out = path_nodes + (size_t)p * max_len;
    const int o = origin[p], d = dest[p];
    if (dist_all[base + d] == kInfBits) { path_len[p] = -1; return; }   // networkx.NetworkXNoPath
    int len = 0, cur = d;
    while (true) {
        if (len >= max_len) { path_len[p] = -2; return; }                // caller's buffer too short
        out[len++] = cur;
        if (cur == o) break;
        cur = pred_all[base + cur];
        if (cur == 0x7fffffff) { path_len[p] = -1; return; }
    }
    for (int a = 0, b = len - 1; a < b; ++a, --b) { const int t = out[a]; out[a] = out[b]; out[b] = t; }
    path_len[p] = len;
}

// models.path_decompiler over a node sequence (models.py:116-137) for graphs whose edges carry no geometry: the positions
// of the route nodes, a point dropped when it equals its successor (the LAST of a run of identical points survives); a
// route of fewer than two nodes has no edges, hence no points.  One thread per pair; `out` == nullptr only counts.
__global__ void __launch_bounds__(128) polyline_kernel(int n_pairs, int max_len, const int32_t* __restrict__ path_len,
                                                       const int32_t* __restrict__ path_nodes,
                                                       const double2* __restrict__ node_xy,
                                                       const long long* __restrict__ offset, int32_t* __restrict__ kept_len,
                                                       double2* __restrict__ out) {
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_pairs) return;
    const int len = path_len[p];
    const int32_t* nodes = path_nodes + (size_t)p * max_len;
    int kept = 0;
    if (len >= 2) {
        double2* dst = out ? out + offset[p] : nullptr;
        double2 cur = node_xy[nodes[0]];
        for (int q = 0; q < len; ++q) {
            const bool last = q == len - 1;
            double2 nxt = cur;
            if (!last) nxt = node_xy[nodes[q + 1]];
            if (last || cur.x != nxt.x || cur.y != nxt.y) {
                if (dst) dst[kept] = cur;
                ++kept;
            }
            cur = nxt;
        }
    }
    if (!out) kept_len[p] = kept;
}

thread_local std::string g_route_err;
thread_local std::vector<double> g_poly;   // polylines of the last tb2_route_polylines call on this thread

// Grow-only device workspace, kept between calls (an initialiser calls tb2_route_paths several times: cars, then one
// batch per light face set; a dozen cudaMalloc/cudaFree per call cost more than the searches themselves).
struct RouteWs {
    void* p[20] = {};
    size_t cap[20] = {};
    int dev = -1;
    void release() {
        for (int k = 0; k < 20; ++k) { if (p[k]) cudaFree(p[k]); p[k] = nullptr; cap[k] = 0; }
    }
};
thread_local RouteWs g_ws;

cudaError_t ws_get(int k, size_t bytes, void** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (g_ws.dev != dev) { g_ws.release(); g_ws.dev = dev; }
    if (bytes < 16) bytes = 16;
    if (g_ws.cap[k] < bytes) {
        if (g_ws.p[k]) cudaFree(g_ws.p[k]);
        g_ws.p[k] = nullptr; g_ws.cap[k] = 0;
        e = cudaMalloc(&g_ws.p[k], bytes + bytes / 4);
        if (e != cudaSuccess) return e;
        g_ws.cap[k] = bytes + bytes / 4;
    }
    *out = g_ws.p[k];
    return cudaSuccess;
}

}  // namespace

extern "C" {

const char* tb2_route_last_error(void) { return g_route_err.c_str(); }
void tb2_route_release(void) { g_ws.release(); std::vector<double>().swap(g_poly); }

// CSR graph on node indices 0..n_nodes-1 (host arrays).  For every pair p writes the node sequence origin..destination
// to path_nodes[p*max_len ...] and its length to path_len[p] (-1: no path, -2: longer than max_len).  `limit` < 0 means
// unbounded search; otherwise nodes farther than `limit` from the origin are not explored.  Synchronous.
static int route_impl(int32_t n_nodes, const int32_t* indptr, const int32_t* indices, const double* weight,
                      const double* node_xy, int32_t n_pairs, const int32_t* origin, const int32_t* dest, double limit,
                      int32_t max_len, int32_t* path_len, int32_t* path_nodes, int64_t* path_ptr, const double** path_xy_out,
                      void* stream_) {
#define RCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) { g_route_err = std::string(#call) + ": " + cudaGetErrorString(e__); rc = TB2_ERR_CUDA; goto done; } \
    } while (0)
    if (n_nodes <= 0 || n_pairs < 0 || max_len <= 0 || !indptr || !indices || !weight || !path_len ||
        (!node_xy && !path_nodes) || (node_xy && !(path_ptr && path_xy_out))) {
        g_route_err = "tb2_route_paths / tb2_route_polylines: bad argument";
        return TB2_ERR_INVALID;
    }
    if (path_ptr) path_ptr[0] = 0;
    if (path_xy_out) *path_xy_out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        g_route_err = "no CUDA device available - traffic_b200 has no CPU path";
        return TB2_ERR_NO_DEVICE;
    }
    if (n_pairs == 0) return TB2_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = TB2_OK;
    // distinct origins -> source slots
    std::vector<int32_t> slot_of_node(n_nodes, -1), sources, pair_slot(n_pairs);
    for (int p = 0; p < n_pairs; ++p) {
        const int o = origin[p];
        if (o < 0 || o >= n_nodes || dest[p] < 0 || dest[p] >= n_nodes) { g_route_err = "node index out of range"; return TB2_ERR_INVALID; }
        if (slot_of_node[o] < 0) { slot_of_node[o] = (int32_t)sources.size(); sources.push_back(o); }
        pair_slot[p] = slot_of_node[o];
    }
    const int n_edges = indptr[n_nodes];
    const double lim = limit < 0 ? __builtin_inf() : limit;
    const size_t smem_bytes = (size_t)n_nodes * 14;   // dist 8 + hop 4 + two frontier flags
    const bool use_smem = smem_bytes <= 200 * 1024;
    // chunk the sources so that dist + pred stay under ~2 GB
    const size_t per_source = (size_t)n_nodes * (use_smem ? 12 : 18);
    size_t chunk_sz = ((size_t)2 << 30) / (per_source > 0 ? per_source : 1);
    if (chunk_sz < 1) chunk_sz = 1;
    int chunk = chunk_sz > (size_t)1 << 20 ? 1 << 20 : (int)chunk_sz;
    if (chunk > (int)sources.size()) chunk = (int)sources.size();
    int32_t *d_indptr = nullptr, *d_indices = nullptr, *d_sources = nullptr, *d_pred = nullptr, *d_pair_slot = nullptr,
            *d_origin = nullptr, *d_dest = nullptr, *d_len = nullptr, *d_nodes = nullptr, *d_hop = nullptr;
    uint8_t* d_flag = nullptr;
    double* d_weight = nullptr;
    unsigned long long* d_dist = nullptr;
    std::vector<int32_t> sel_pairs, sel_slot, sel_o, sel_d, tmp_len, tmp_nodes, tmp_kept;
    std::vector<int64_t> kept_of_pair(node_xy ? n_pairs : 0, 0);
    std::vector<long long> tmp_off;
    std::vector<std::vector<double>> chunk_xy;        // polylines of each chunk, pairs in chunk order
    std::vector<std::vector<int32_t>> chunk_pairs;
    double2* d_node_xy = nullptr;
    double2* d_xy = nullptr;
    long long* d_off = nullptr;
    int32_t* d_kept = nullptr;
    if (node_xy) {
        RCU(ws_get(13, sizeof(double2) * (size_t)n_nodes, (void**)&d_node_xy));
        RCU(cudaMemcpyAsync(d_node_xy, node_xy, sizeof(double2) * (size_t)n_nodes, cudaMemcpyDefault, stream));
    }
    RCU(ws_get(0, sizeof(int32_t) * (n_nodes + 1), (void**)&d_indptr));
    RCU(ws_get(1, sizeof(int32_t) * (n_edges > 0 ? n_edges : 1), (void**)&d_indices));
    RCU(ws_get(2, sizeof(double) * (n_edges > 0 ? n_edges : 1), (void**)&d_weight));
    RCU(ws_get(3, sizeof(int32_t) * chunk, (void**)&d_sources));
    RCU(ws_get(4, sizeof(unsigned long long) * (size_t)chunk * n_nodes, (void**)&d_dist));
    RCU(ws_get(5, sizeof(int32_t) * (size_t)chunk * n_nodes, (void**)&d_pred));
    if (use_smem) {
        RCU(cudaFuncSetAttribute(sssp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    } else {
        RCU(ws_get(6, sizeof(int32_t) * (size_t)chunk * n_nodes, (void**)&d_hop));
        RCU(ws_get(7, (size_t)chunk * 2 * n_nodes, (void**)&d_flag));
    }
    RCU(cudaMemcpyAsync(d_indptr, indptr, sizeof(int32_t) * (n_nodes + 1), cudaMemcpyDefault, stream));
    RCU(cudaMemcpyAsync(d_indices, indices, sizeof(int32_t) * n_edges, cudaMemcpyDefault, stream));
    RCU(cudaMemcpyAsync(d_weight, weight, sizeof(double) * n_edges, cudaMemcpyDefault, stream));
    {
        // pairs grouped by source chunk
        std::vector<std::vector<int32_t>> by_chunk((sources.size() + chunk - 1) / chunk);
        for (int p = 0; p < n_pairs; ++p) by_chunk[pair_slot[p] / chunk].push_back(p);
        size_t max_pairs = 0;
        for (auto& v : by_chunk) max_pairs = v.size() > max_pairs ? v.size() : max_pairs;
        RCU(ws_get(8, sizeof(int32_t) * max_pairs, (void**)&d_pair_slot));
        RCU(ws_get(9, sizeof(int32_t) * max_pairs, (void**)&d_origin));
        RCU(ws_get(10, sizeof(int32_t) * max_pairs, (void**)&d_dest));
        RCU(ws_get(11, sizeof(int32_t) * max_pairs, (void**)&d_len));
        RCU(ws_get(12, sizeof(int32_t) * max_pairs * (size_t)max_len, (void**)&d_nodes));
        for (size_t c = 0; c < by_chunk.size(); ++c) {
            const int s0 = (int)c * chunk;
            const int ns = (int)sources.size() - s0 < chunk ? (int)sources.size() - s0 : chunk;
            const auto& pairs = by_chunk[c];
            if (pairs.empty()) continue;
            RCU(cudaMemcpyAsync(d_sources, sources.data() + s0, sizeof(int32_t) * ns, cudaMemcpyDefault, stream));
            if (use_smem)
                sssp_kernel<true><<<ns, 256, smem_bytes, stream>>>(n_nodes, d_indptr, d_indices, d_weight, d_sources, lim, d_dist,
                                                                   d_pred, nullptr, nullptr);
            else
                sssp_kernel<false><<<ns, 256, 0, stream>>>(n_nodes, d_indptr, d_indices, d_weight, d_sources, lim, d_dist,
                                                           d_pred, d_hop, d_flag);
            RCU(cudaGetLastError());
            const int np = (int)pairs.size();
            sel_slot.resize(np); sel_o.resize(np); sel_d.resize(np);
            for (int k = 0; k < np; ++k) {
                sel_slot[k] = pair_slot[pairs[k]] - s0; sel_o[k] = origin[pairs[k]]; sel_d[k] = dest[pairs[k]];
            }
            RCU(cudaMemcpyAsync(d_pair_slot, sel_slot.data(), sizeof(int32_t) * np, cudaMemcpyDefault, stream));
            RCU(cudaMemcpyAsync(d_origin, sel_o.data(), sizeof(int32_t) * np, cudaMemcpyDefault, stream));
            RCU(cudaMemcpyAsync(d_dest, sel_d.data(), sizeof(int32_t) * np, cudaMemcpyDefault, stream));
            extract_paths_kernel<<<(np + 127) / 128, 128, 0, stream>>>(np, n_nodes, d_pair_slot, d_origin, d_dest, d_dist, d_pred,
                                                                       max_len, d_len, d_nodes);
            RCU(cudaGetLastError());
            tmp_len.resize(np);
            RCU(cudaMemcpyAsync(tmp_len.data(), d_len, sizeof(int32_t) * np, cudaMemcpyDefault, stream));
            if (path_nodes) {
                tmp_nodes.resize((size_t)np * max_len);
                RCU(cudaMemcpyAsync(tmp_nodes.data(), d_nodes, sizeof(int32_t) * (size_t)np * max_len, cudaMemcpyDefault, stream));
            }
            if (node_xy) {   // polylines of this chunk, emitted on the device as a compact point array
                RCU(ws_get(14, sizeof(int32_t) * max_pairs, (void**)&d_kept));
                RCU(ws_get(15, sizeof(long long) * max_pairs, (void**)&d_off));
                polyline_kernel<<<(np + 127) / 128, 128, 0, stream>>>(np, max_len, d_len, d_nodes, d_node_xy, nullptr, d_kept, nullptr);
                RCU(cudaGetLastError());
                tmp_kept.resize(np);
                RCU(cudaMemcpyAsync(tmp_kept.data(), d_kept, sizeof(int32_t) * np, cudaMemcpyDefault, stream));
                RCU(cudaStreamSynchronize(stream));
                tmp_off.resize(np);
                long long total = 0;
                for (int k = 0; k < np; ++k) { tmp_off[k] = total; total += tmp_kept[k]; kept_of_pair[pairs[k]] = tmp_kept[k]; }
                chunk_pairs.push_back(pairs);
                chunk_xy.emplace_back((size_t)total * 2);
                if (total > 0) {
                    RCU(ws_get(16, sizeof(double2) * (size_t)total, (void**)&d_xy));
                    RCU(cudaMemcpyAsync(d_off, tmp_off.data(), sizeof(long long) * np, cudaMemcpyDefault, stream));
                    polyline_kernel<<<(np + 127) / 128, 128, 0, stream>>>(np, max_len, d_len, d_nodes, d_node_xy, d_off, nullptr, d_xy);
                    RCU(cudaGetLastError());
                    RCU(cudaMemcpyAsync(chunk_xy.back().data(), d_xy, sizeof(double2) * (size_t)total, cudaMemcpyDefault, stream));
                    RCU(cudaStreamSynchronize(stream));
                }
            }
            RCU(cudaStreamSynchronize(stream));
            for (int k = 0; k < np; ++k) {
                const int p = pairs[k];
                path_len[p] = tmp_len[k];
                if (!path_nodes) continue;
                const int n = tmp_len[k] > 0 ? tmp_len[k] : 0;
                for (int q = 0; q < n; ++q) path_nodes[(size_t)p * max_len + q] = tmp_nodes[(size_t)k * max_len + q];
            }
        }
    }
    if (node_xy) {   // CSR over the pairs in the caller's order
        for (int p = 0; p < n_pairs; ++p) path_ptr[p + 1] = path_ptr[p] + kept_of_pair[p];
        g_poly.resize((size_t)path_ptr[n_pairs] * 2);
        double* path_xy = g_poly.data();
        *path_xy_out = path_xy;
        for (size_t c = 0; c < chunk_pairs.size(); ++c) {
            const double* src = chunk_xy[c].data();
            for (int p : chunk_pairs[c]) {
                const int64_t n = kept_of_pair[p];
                if (n > 0) memcpy(path_xy + 2 * path_ptr[p], src, sizeof(double) * 2 * (size_t)n);
                src += 2 * n;
            }
        }
    }
done:
    // (device buffers stay in the thread's workspace for the next call; tb2_route_release() frees them)
    return rc;
#undef RCU
}

int tb2_route_paths(int32_t n_nodes, const int32_t* indptr, const int32_t* indices, const double* weight,
                    int32_t n_pairs, const int32_t* origin, const int32_t* dest, double limit, int32_t max_len,
                    int32_t* path_len, int32_t* path_nodes, void* stream) {
    return route_impl(n_nodes, indptr, indices, weight, nullptr, n_pairs, origin, dest, limit, max_len, path_len, path_nodes,
                      nullptr, nullptr, stream);
}

// tb2_route_paths + the polylines (xpath / ypath of every pair) as one compact point array, emitted on the device:
// (*path_xy_out)[2 * path_ptr[p] ... 2 * path_ptr[p + 1]) are the points of pair p.  path_nodes may be NULL.
int tb2_route_polylines(int32_t n_nodes, const int32_t* indptr, const int32_t* indices, const double* weight,
                        const double* node_xy, int32_t n_pairs, const int32_t* origin, const int32_t* dest, double limit,
                        int32_t max_len, int32_t* path_len, int32_t* path_nodes, int64_t* path_ptr,
                        const double** path_xy_out, void* stream) {
    if (!node_xy) { g_route_err = "tb2_route_polylines: node_xy is NULL"; return TB2_ERR_INVALID; }
    return route_impl(n_nodes, indptr, indices, weight, node_xy, n_pairs, origin, dest, limit, max_len, path_len, path_nodes,
                      path_ptr, path_xy_out, stream);
}

}  // extern "C"
