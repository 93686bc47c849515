/*
 * traffic_b200 - C ABI of the B200-native per-timestep vehicle update.
 *
 * This is the drop-in boundary for the hot path of donjpierce/traffic: the loop body of
 * test/profile_cars_update.py:28-34, i.e. Cars.update(dt, lights) (cars.py:61-95) and
 * TrafficLights.update(dt) (cars.py:123-133).  The reference has no FFI layer of its own (pure Python); the
 * binding a maintainer adds is a ctypes stub inside cars.py (shown in INTEGRATION.md), and the entry points below
 * are exactly what that stub binds.  Plain pointers and sizes only; no torch types.
 *
 * Memory model: one simulation = one opaque tb2_sim that owns a single device arena.  The arena is either
 * supplied by the host (a torch uint8 tensor's data_ptr, so PyTorch's caching allocator owns the memory and
 * re-creating a Cars object per episode is cheap - environment.py:50-51,85) or cudaMalloc'ed by the library when
 * NULL is passed.  Uploads/downloads accept host OR device pointers (cudaMemcpyDefault).  All work is issued on
 * the caller's stream (a cudaStream_t passed as void*); nothing synchronises unless stated.
 *
 * Every function returns 0 on success and a non-zero tb2_status on failure; tb2_last_error() gives the message.
 * There is no CPU fallback: without a CUDA device tb2_create fails.
 */
#ifndef TRAFFIC_B200_H
#define TRAFFIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB2_ABI_VERSION 2

/* order of the two reference calls inside one step */
#define TB2_ORDER_CARS_LIGHTS 0 /* test/profile_cars_update.py:29-34 */
#define TB2_ORDER_LIGHTS_CARS 1 /* animate.py:63-64, environment.py:167-168 */

/* how the per-bin (first, second) car-index table behind the leader lookup is built each step */
#define TB2_NEIGHBOR_ATOMIC 0 /* two atomicMin per car at the tail of the step kernel (default, no extra launch)      */
#define TB2_NEIGHBOR_SORT 1   /* stable radix sort of the cars by bin + warp-shuffle segment heads (same result)       */

/* arithmetic modes */
#define TB2_F64 0 /* reference arithmetic (numpy float64 evaluation order) - parity mode            */
#define TB2_F32 1 /* throughput mode: fp32 state in local-origin coordinates, tolerances from absolute */

typedef enum tb2_status {
    TB2_OK = 0,
    TB2_ERR_INVALID = 1,   /* bad argument / size mismatch */
    TB2_ERR_CUDA = 2,      /* a CUDA runtime call failed (message has the cudaError string) */
    TB2_ERR_NO_DEVICE = 3, /* no CUDA device: the library has no CPU path */
    TB2_ERR_STATE = 4      /* call sequence error (e.g. step before prepare) */
} tb2_status;

/* Fields of the device-resident structure of arrays.  Element types are for TB2_F64 / TB2_F32. */
typedef enum tb2_field {
    TB2_POS = 0,          /* [N][2] f64 | f32 (f32: relative to origin_x/origin_y)  cars 'x','y'          */
    TB2_VEL = 1,          /* [N][2] f64 | f32                                       cars 'vx','vy'        */
    TB2_ROUTE_TIME = 2,   /* [N]    f64 | f32                                       cars 'route-time'     */
    TB2_CURSOR = 3,       /* [N]    i32  absolute index of the path head in PATH_XY (popped points = cursor - start) */
    TB2_PATH_END = 4,     /* [N]    i32  one past the car's last polyline point                           */
    TB2_PATH_EFF_END = 5, /* [N]    i32  has_path <=> cursor < eff_end (navigation.py:37-39 `.any()` rule) */
    TB2_PATH_XY = 6,      /* [P][2] f64 | f32  all polylines back to back (xpath/ypath)                   */
    TB2_DEST_XY = 7,      /* [N][2] f64 | f32  destination node position (navigation.py:84,89)            */
    TB2_D_NODE = 8,       /* [N]    f64 | f32  'distance-to-node'                                         */
    TB2_D_CAR = 9,        /* [N]    f64 | f32  'distance-to-car', 0 encodes False                         */
    TB2_D_LIGHT = 10,     /* [N]    f64 | f32  'distance-to-red-light', +0 = False, -0 = None             */
    TB2_CELL = 11,        /* [N]    i32  ybin*(nbx+1)+xbin of the start-of-step position ('xbin','ybin')   */
    TB2_LEADER = 12,      /* [N]    i32  index of the car accepted as obstacle, -1 if none                */
    TB2_LIGHT_XY = 13,    /* [L][2] f64 | f32                                                             */
    TB2_SWITCH_TIME = 14, /* [L]    f64 (always f64: cars.py:131 is a round-off artefact, Appendix B-6)   */
    TB2_FACE_PTR = 15,    /* [L+1]  i32                                                                   */
    TB2_FACE_VEC = 16,    /* [F][2] f64 | f32  'out-xvectors','out-yvectors'                              */
    TB2_FACE_GO = 17,     /* [F]    u8   'go-values'                                                      */
    TB2_CELL_LIGHT_PTR = 18, /* [C+1] i32  static cell -> lights CSR (lights in DataFrame order)           */
    TB2_CELL_LIGHT_IDX = 19, /* [LC]  i32                                                                  */
    TB2_LIGHTS_TIME = 20, /* [1]    f64  TrafficLights.time_elapsed (cars.py:117,130)                     */
    /* replica ensembles (environment.Env rollouts, environment.py:97-173): per replica */
    TB2_DONE = 21,        /* [R]    i32  1 once the agent car satisfied FrontView.end_of_route (navigation.py:113-128) */
    TB2_TRIP_TIME = 22,   /* [R]    f64  the agent's 'route-time' when it arrived (environment.py:126)        */
    TB2_TRIP_STEP = 23,   /* [R]    i32  number of simulation steps executed before arrival                   */
    TB2_FIELD_COUNT = 24
} tb2_field;

typedef struct tb2_config {
    int32_t abi_version; /* must be TB2_ABI_VERSION */
    int32_t precision;   /* TB2_F64 | TB2_F32 */
    int32_t n_cars, n_lights, n_faces;
    int32_t n_cell_lights;     /* length of TB2_CELL_LIGHT_IDX */
    int64_t n_path_points;     /* P */
    /* bins: models.determine_bins (models.py:7-19).  nb = len(np.arange(lo, hi, 200)); e0,e1 = first two edges,
     * ed = e1 - e0 (numpy's arange fill rule, so that edge k = e0 + k*ed bit-for-bit). */
    int32_t nbx, nby;
    double ex0, ex1, exd;
    double ey0, ey1, eyd;
    /* simulation.py:13-16 */
    double speed_limit, stop_distance, free_distance, default_acceleration;
    /* TB2_F32 only: positions are stored as (x - origin_x, y - origin_y) */
    double origin_x, origin_y;
    int32_t write_aux;   /* 1: TB2_VEL, TB2_D_*, TB2_LEADER and TB2_CELL are valid after every tb2_cars_update and
                          *    after the LAST step of every tb2_step batch (columns nobody can observe between the steps
                          *    of a batch are not written); 0: never written (pure stepping) */
    int32_t neighbor_mode; /* TB2_NEIGHBOR_ATOMIC (default) | TB2_NEIGHBOR_SORT */
    /* Replica ensemble: n_replicas independent copies of the same world (same graph, routes and light geometry) with
     * their own car state, light timings (TB2_SWITCH_TIME [R*L]) and face states (TB2_FACE_GO [R*F]), advanced by
     * the same launches.  n_cars / n_lights / n_faces are PER REPLICA; per-car fields hold R*n_cars elements, replica
     * r at [r*n_cars, (r+1)*n_cars); TB2_LEADER holds indices local to the replica.  1 = a single simulation. */
    int32_t n_replicas;
    int32_t agent_car;   /* car whose arrival ends a rollout (learn.py:14 uses 17); -1: no agent tracking */
} tb2_config;

typedef struct tb2_sim tb2_sim;

int tb2_abi_version(void);
const char* tb2_last_error(void);

/* bytes of device memory a simulation with this config needs (the host may allocate it, e.g. with torch) */
size_t tb2_arena_bytes(const tb2_config* cfg);

/* replaces Cars.__init__ + TrafficLights.__init__ (cars.py:23-41, 110-121): allocates/carves device state.
 * arena_dev may be NULL (library allocates). */
int tb2_create(const tb2_config* cfg, void* arena_dev, size_t arena_bytes, tb2_sim** out);
int tb2_destroy(tb2_sim* sim);

/* element count / element size / current device pointer of a field (POS flips buffers every step) */
int tb2_field_info(const tb2_sim* sim, int field, size_t* n_elems, size_t* elem_bytes);
void* tb2_field_ptr(const tb2_sim* sim, int field);

/* src/dst: host or device memory; bytes must equal n_elems*elem_bytes */
int tb2_upload(tb2_sim* sim, int field, const void* src, size_t bytes, void* stream);
int tb2_download(const tb2_sim* sim, int field, void* dst, size_t bytes, void* stream);
/* elements [first, first + count) of a field - what `state.loc[agent]` (environment.py:126,161) needs is one row */
int tb2_download_range(const tb2_sim* sim, int field, size_t first, size_t count, void* dst, size_t bytes, void* stream);

/* (re)build the cell -> two-smallest-car-index table from the current positions; call after uploading POS
 * (replaces the bin assignment at the end of the reference initialisers, simulation.py:246-250) */
int tb2_prepare(tb2_sim* sim, void* stream);

/* Cars.update(dt, lights) alone (cars.py:61-95): one fused launch, lights untouched.  The light-face states it
 * reads are the device-resident TB2_FACE_GO (upload them first if the caller's lights live elsewhere). */
int tb2_cars_update(tb2_sim* sim, double dt, void* stream);
/* TrafficLights.update(dt) alone (cars.py:123-133) */
int tb2_lights_update(tb2_sim* sim, double dt, void* stream);

/* n_steps of { Cars.update(dt, lights.state) ; TrafficLights.update(dt) } in the given order (cars.py:61-95,
 * 123-133).  Asynchronous on `stream`. */
int tb2_step(tb2_sim* sim, double dt, int32_t n_steps, int32_t order, void* stream);

/* Replica ensembles (tb2_config.n_replicas, agent_car): the rollout loop of environment.Env.step / simulation_step
 * (environment.py:97-173) - every replica advances until its agent satisfies FrontView.end_of_route
 * (navigation.py:113-128, tested before each step) or max_steps is reached (the reference loop has no cap and may never
 * terminate).  Small ensembles: one launch, each replica is one thread-block cluster whose world
 * stays on chip and which leaves the launch when its agent has arrived.  Large ensembles: batched per-step launches of
 * the HBM-streaming kernels; the cars of an arrived replica are no longer loaded or stepped, the done flags are read
 * back every 256 steps - to stop as soon as every rollout is over, and to hand the kernel the list of the replicas still
 * running so that it walks only their cars.  A replica that has arrived stays frozen in later
 * tb2_rollout calls; the state of its cars is the one of the arrival step (Env.step has returned, environment.py:124)
 * and is not meant to be stepped further with tb2_step.  Results: TB2_DONE / TB2_TRIP_TIME / TB2_TRIP_STEP. */
int tb2_rollout(tb2_sim* sim, double dt, int32_t max_steps, int32_t order, void* stream);

/* End-to-end variant with HOST buffers (what Cars.update does for a caller that passes the lights DataFrame and
 * reads positions back, animate.py:63-69): optional H2D of the face states, n_steps, D2H of positions (frame_xy,
 * [N][2] in the sim's precision) and, when non-NULL, the face states.  Synchronises `stream` before returning. */
int tb2_step_host(tb2_sim* sim, double dt, int32_t n_steps, int32_t order, const uint8_t* face_go_in,
                  void* frame_xy_out, uint8_t* face_go_out, void* stream);

/* Pipelined frame extraction (animate.py-style consumers): like tb2_step_host but returns once the work
 * is enqueued; the frame is snapshotted on the device and travels on a private copy stream while the caller enqueues the
 * next batch.  Host buffers should be pinned.  The face states are copied first: tb2_frame_wait(sim, 1) returns when they
 * have landed (microseconds), tb2_frame_wait(sim, 0) when the position frame has landed too.  A new tb2_step_host_async
 * may be issued before the positions have landed (it orders itself after the pending copy); use a second position buffer
 * for it. */
int tb2_step_host_async(tb2_sim* sim, double dt, int32_t n_steps, int32_t order, const uint8_t* face_go_in,
                        void* frame_xy_out, uint8_t* face_go_out, void* stream);
int tb2_frame_wait(tb2_sim* sim, int32_t faces_only);

/* Init-time routing (the producer of the hot path's inputs): for every (origin, destination) pair the
 * shortest path by edge weight, i.e. networkx.shortest_path(G, o, d, weight='length') as navigation.get_route /
 * get_init_path / lines_to_node call it once per car and per light face (navigation.py:581-609, 764-841).  The graph is
 * a CSR over node indices 0..n_nodes-1 (host arrays; parallel edges allowed, the shortest one wins as in
 * navigation.py:820).  Writes the node sequence origin..destination of pair p to path_nodes[p*max_len ...] and its length
 * to path_len[p] (-1: no path = networkx.NetworkXNoPath, -2: longer than max_len).  limit < 0: unbounded search, else
 * nodes farther than `limit` from the origin are not explored.  Synchronous. */
int tb2_route_paths(int32_t n_nodes, const int32_t* indptr, const int32_t* indices, const double* weight,
                    int32_t n_pairs, const int32_t* origin, const int32_t* dest, double limit, int32_t max_len,
                    int32_t* path_len, int32_t* path_nodes, void* stream);
/* tb2_route_paths plus the polylines the initialisers store per car (xpath / ypath): what navigation.shortest_path_lines_nx
 * + models.path_decompiler build from the node sequence (navigation.py:801-841, models.py:116-137) for a graph whose edges
 * carry no OSMnx geometry - the positions of the route nodes, a point dropped when it equals its successor, nothing for a
 * route of fewer than two nodes.  The points are emitted on the device as one compact array: pair p owns
 * (*path_xy_out)[2*path_ptr[p] ... 2*path_ptr[p+1]) (x, y interleaved; path_ptr has n_pairs + 1 entries).  The point array
 * belongs to the library (its size is only known after the search): it stays valid until the next tb2_route_polylines /
 * tb2_route_release call on the same host thread.  node_xy: n_nodes x 2 doubles.  path_nodes may be NULL.  Synchronous. */
int tb2_route_polylines(int32_t n_nodes, const int32_t* indptr, const int32_t* indices, const double* weight,
                        const double* node_xy, int32_t n_pairs, const int32_t* origin, const int32_t* dest, double limit,
                        int32_t max_len, int32_t* path_len, int32_t* path_nodes, int64_t* path_ptr,
                        const double** path_xy_out, void* stream);
const char* tb2_route_last_error(void);
/* tb2_route_paths keeps its device workspace between calls (per host thread); this frees it */
void tb2_route_release(void);

/* Builds with -DTB2_DEBUG_BOUNDS check every gather index of the fp64 step kernels before it is dereferenced (the GPU
 * pool's compute-sanitizer is closed): bit mask of violated rules since the last call; always 0 in regular builds. */
int tb2_debug_bounds_enabled(void);
unsigned int tb2_debug_bounds_violations(void);

/* number of kernel launches issued by this sim so far (bench.py's gpu_launches) */
int64_t tb2_launch_count(const tb2_sim* sim);
int64_t tb2_step_count(const tb2_sim* sim);

#ifdef __cplusplus
}
#endif
#endif /* TRAFFIC_B200_H */
