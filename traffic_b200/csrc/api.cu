// C ABI of libtraffic_b200.so (declared in include/traffic_b200.h): arena carving, uploads/downloads, and the
// launch sequence of a simulation step.  Host-side only; the kernels live in step_kernel.cuh (instantiated for fp64 in
// step_f64.cu and for fp32 in step_f32.cu).
#include "../../include/traffic_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>

#include "step_args.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return TB2_ERR_CUDA;
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

constexpr size_t kAlign = 256;
size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

// internal buffer ids (superset of tb2_field: ping-pong partners, caches and the cell tables)
enum Buf {
    B_POS0, B_POS1, B_VEL, B_ROUTE_TIME, B_CURSOR, B_PATH_END, B_PATH_EFF_END, B_PATH_XY, B_DEST_XY,
    B_D_NODE, B_D_CAR, B_D_LIGHT, B_CELL, B_CELL_PREV, B_LEADER, B_CELL_DYN, B_LIGHT_XY, B_SWITCH_TIME, B_FACE_PTR,
    B_FACE_VEC, B_GO0, B_GO1, B_CELL_LIGHT_PTR, B_CELL_LIGHT_IDX, B_T0, B_T1, B_HEAD_XY, B_HEAD_CURV, B_PT_CURV,
    B_FACE_UNIT, B_DONE, B_TRIP_TIME, B_TRIP_STEP, B_CELL_STAT, B_LIGHT_CELLINFO, B_LOG_TABLE, B_FRAME_STAGE, B_GO_STAGE, B_SORT_KEYS_IN, B_SORT_VALS_IN, B_SORT_KEYS_OUT,
    B_SORT_VALS_OUT, B_SORT_TEMP, B_LIVE_LIST, B_DONE_V, B_COUNT
};

struct Layout {
    size_t off[B_COUNT];
    size_t bytes[B_COUNT];
    size_t total;
};

bool config_ok(const tb2_config* c, std::string* why) {
    if (!c) { *why = "null config"; return false; }
    if (c->abi_version != TB2_ABI_VERSION) { *why = "abi_version mismatch"; return false; }
    if (c->precision != TB2_F64 && c->precision != TB2_F32) { *why = "precision must be TB2_F64 or TB2_F32"; return false; }
    if (c->n_cars < 0 || c->n_lights < 0 || c->n_faces < 0 || c->n_cell_lights < 0 || c->n_path_points < 0) {
        *why = "negative size"; return false;
    }
    if (c->nbx < 0 || c->nby < 0) { *why = "negative bin count"; return false; }
    if ((int64_t)(c->nbx + 1) * (c->nby + 1) > (int64_t)1 << 28) { *why = "too many cells"; return false; }
    if (c->n_path_points >= ((int64_t)1 << 31)) { *why = "more than 2^31 path points"; return false; }
    if (c->n_replicas < 1) { *why = "n_replicas must be >= 1"; return false; }
    if ((int64_t)c->n_replicas * c->n_cars >= ((int64_t)1 << 31) - 1) { *why = "more than 2^31 cars"; return false; }
    if ((int64_t)c->n_replicas * (c->nbx + 1) * (c->nby + 1) > (int64_t)1 << 29) { *why = "too many cells x replicas"; return false; }
    if (c->agent_car >= c->n_cars) { *why = "agent_car out of range"; return false; }
    if (c->neighbor_mode != TB2_NEIGHBOR_ATOMIC && c->neighbor_mode != TB2_NEIGHBOR_SORT) { *why = "bad neighbor_mode"; return false; }
    return true;
}

int key_bits_of(const tb2_config& c) {
    const int64_t n_keys = (int64_t)c.n_replicas * (c.nbx + 1) * (c.nby + 1);
    int bits = 1;
    while (((int64_t)1 << bits) < n_keys) ++bits;
    return bits;
}

Layout make_layout(const tb2_config& c) {
    // per-car / per-replica buffers hold all replicas back to back; geometry (paths, light positions, faces) is shared
    const size_t Rep = (size_t)c.n_replicas;
    const size_t N = Rep * (size_t)c.n_cars, L = (size_t)c.n_lights, F = (size_t)c.n_faces, P = (size_t)c.n_path_points;
    const size_t C = (size_t)(c.nbx + 1) * (size_t)(c.nby + 1), LC = (size_t)c.n_cell_lights;
    const size_t R = c.precision == TB2_F64 ? sizeof(double) : sizeof(float);
    Layout l{};
    size_t b[B_COUNT];
    b[B_POS0] = b[B_POS1] = N * 2 * R;
    b[B_VEL] = N * 2 * R;
    b[B_ROUTE_TIME] = N * R;
    b[B_CURSOR] = b[B_PATH_END] = b[B_PATH_EFF_END] = N * 4;
    b[B_PATH_XY] = P * 2 * R;
    b[B_DEST_XY] = N * 2 * R;
    b[B_D_NODE] = b[B_D_CAR] = b[B_D_LIGHT] = N * R;
    b[B_CELL] = b[B_CELL_PREV] = N * 4;
    b[B_HEAD_XY] = b[B_HEAD_CURV] = N * 2 * R;
    b[B_PT_CURV] = P * 2 * R;
    b[B_FACE_UNIT] = F * 2 * R;
    b[B_LEADER] = N * 4;
    b[B_CELL_DYN] = C * Rep * sizeof(CellDynT);   // three rotating (min, second) tables + two go words per bin
    b[B_LIGHT_XY] = L * 2 * R;
    b[B_SWITCH_TIME] = Rep * L * 8;
    b[B_FACE_PTR] = (L + 1) * 4;
    b[B_FACE_VEC] = F * 2 * R;
    b[B_GO0] = b[B_GO1] = Rep * F;
    b[B_CELL_LIGHT_PTR] = (C + 1) * 4;
    b[B_CELL_LIGHT_IDX] = LC * 4;
    b[B_T0] = b[B_T1] = 8;
    b[B_DONE] = b[B_TRIP_STEP] = b[B_LIVE_LIST] = b[B_DONE_V] = Rep * 4;
    b[B_TRIP_TIME] = Rep * 8;
    b[B_CELL_STAT] = C * 32;
    b[B_LIGHT_CELLINFO] = L * 8;
    b[B_LOG_TABLE] = 128 * 32;   // LogEntry[TB2_LOG_N]
    b[B_FRAME_STAGE] = N * 2 * R;   // snapshot of the positions / faces that a pipelined frame copy reads from
    b[B_GO_STAGE] = Rep * F;
    const bool sorted = c.neighbor_mode == TB2_NEIGHBOR_SORT;
    b[B_SORT_KEYS_IN] = b[B_SORT_VALS_IN] = b[B_SORT_KEYS_OUT] = b[B_SORT_VALS_OUT] = sorted ? N * 4 : 0;
    b[B_SORT_TEMP] = sorted ? neighbor_sort_temp_bytes((int)N, key_bits_of(c)) : 0;
    size_t off = 0;
    for (int i = 0; i < B_COUNT; ++i) {
        l.off[i] = off;
        l.bytes[i] = b[i];
        off += align_up(b[i] ? b[i] : 1);
    }
    l.total = off;
    return l;
}

}  // namespace

struct tb2_sim {
    tb2_config cfg;
    Layout lay;
    char* arena = nullptr;
    bool owns_arena = false;
    int n_cells = 0;
    int cur_pos = 0, cur_tab = 0, cur_go = 0, cur_t = 0;
    bool prepared = false;
    bool freeze_done = false;       // inside tb2_rollout: arrived replicas are frozen
    int n_live = 0;                 // inside tb2_rollout, > 0: B_LIVE_LIST holds the replicas still running (walked by pipe_kernel)
    bool go_dirty = false;          // TB2_FACE_GO was uploaded: the per-bin go words must be re-derived before the next step
    bool allow_pipe = true;         // TB2_NO_PIPE=1: big worlds use the plain one-thread-per-car kernel (A/B testing)
    bool pipe_tma = false;          // TB2_PIPE_TMA=1: own-state columns travel as TMA bulk copies (measured slower, see step_kernel.cuh)
    bool force_pipe = false;        // TB2_FORCE_PIPE=1: every world goes through the pipelined kernel (test coverage)
    int n_sms = 0;
    bool allow_resident = true;     // TB2_NO_RESIDENT=1: small worlds / ensembles use the global-memory kernels (A/B testing)
    bool allow_persistent = true;   // TB2_NO_PERSISTENT=1 in the environment forces one launch per step (A/B testing)
    int64_t launches = 0, steps = 0;
    double log_free_over_stop = 0.0;
    // pipelined frame extraction (tb2_step_host_async): a private copy stream and two events
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t frame_ready = nullptr, frame_done = nullptr, faces_done = nullptr;
    bool frame_pending = false;

    template <typename T>
    T* buf(int b) const { return reinterpret_cast<T*>(arena + lay.off[b]); }
    size_t real_bytes() const { return cfg.precision == TB2_F64 ? sizeof(double) : sizeof(float); }
};

namespace {

// map a public field to (internal buffer, element bytes, element count)
int field_buf(const tb2_sim* s, int field, int* b, size_t* elem, size_t* count) {
    const tb2_config& c = s->cfg;
    const size_t Rep = (size_t)c.n_replicas;
    const size_t N = Rep * (size_t)c.n_cars, L = (size_t)c.n_lights, F = (size_t)c.n_faces, P = (size_t)c.n_path_points;
    const size_t R = s->real_bytes();
    switch (field) {
        case TB2_DONE: *b = B_DONE; *elem = 4; *count = Rep; return 0;
        case TB2_TRIP_TIME: *b = B_TRIP_TIME; *elem = 8; *count = Rep; return 0;
        case TB2_TRIP_STEP: *b = B_TRIP_STEP; *elem = 4; *count = Rep; return 0;
        case TB2_POS: *b = s->cur_pos ? B_POS1 : B_POS0; *elem = R; *count = N * 2; return 0;
        case TB2_VEL: *b = B_VEL; *elem = R; *count = N * 2; return 0;
        case TB2_ROUTE_TIME: *b = B_ROUTE_TIME; *elem = R; *count = N; return 0;
        case TB2_CURSOR: *b = B_CURSOR; *elem = 4; *count = N; return 0;
        case TB2_PATH_END: *b = B_PATH_END; *elem = 4; *count = N; return 0;
        case TB2_PATH_EFF_END: *b = B_PATH_EFF_END; *elem = 4; *count = N; return 0;
        case TB2_PATH_XY: *b = B_PATH_XY; *elem = R; *count = P * 2; return 0;
        case TB2_DEST_XY: *b = B_DEST_XY; *elem = R; *count = N * 2; return 0;
        case TB2_D_NODE: *b = B_D_NODE; *elem = R; *count = N; return 0;
        case TB2_D_CAR: *b = B_D_CAR; *elem = R; *count = N; return 0;
        case TB2_D_LIGHT: *b = B_D_LIGHT; *elem = R; *count = N; return 0;
        case TB2_CELL: *b = B_CELL_PREV; *elem = 4; *count = N; return 0;
        case TB2_LEADER: *b = B_LEADER; *elem = 4; *count = N; return 0;
        case TB2_LIGHT_XY: *b = B_LIGHT_XY; *elem = R; *count = L * 2; return 0;
        case TB2_SWITCH_TIME: *b = B_SWITCH_TIME; *elem = 8; *count = Rep * L; return 0;
        case TB2_FACE_PTR: *b = B_FACE_PTR; *elem = 4; *count = L + 1; return 0;
        case TB2_FACE_VEC: *b = B_FACE_VEC; *elem = R; *count = F * 2; return 0;
        case TB2_FACE_GO: *b = s->cur_go ? B_GO1 : B_GO0; *elem = 1; *count = Rep * F; return 0;
        case TB2_CELL_LIGHT_PTR: *b = B_CELL_LIGHT_PTR; *elem = 4; *count = (size_t)s->n_cells + 1; return 0;
        case TB2_CELL_LIGHT_IDX: *b = B_CELL_LIGHT_IDX; *elem = 4; *count = (size_t)c.n_cell_lights; return 0;
        case TB2_LIGHTS_TIME: *b = s->cur_t ? B_T1 : B_T0; *elem = 8; *count = 1; return 0;
        default: return fail(TB2_ERR_INVALID, "unknown field id");
    }
}

template <typename R>
void fill_args(const tb2_sim* s, StepArgsT<R>* a) {
    using R2 = typename real2<R>::type;
    const tb2_config& c = s->cfg;
    // float mode stores coordinates relative to the origin, so the bin edges shift by the same amount
    const double ox = std::is_same<R, float>::value ? c.origin_x : 0.0;
    const double oy = std::is_same<R, float>::value ? c.origin_y : 0.0;
    a->g.nbx = c.nbx; a->g.nby = c.nby; a->g.ncx = c.nbx + 1; a->g.n_cells = s->n_cells;
    a->g.ex0 = (R)(c.ex0 - ox); a->g.ex1 = (R)(c.ex1 - ox); a->g.exd = (R)c.exd;
    a->g.ey0 = (R)(c.ey0 - oy); a->g.ey1 = (R)(c.ey1 - oy); a->g.eyd = (R)c.eyd;
    a->g.inv_exd = (R)(1.0 / c.exd); a->g.inv_eyd = (R)(1.0 / c.eyd);
    a->g.f_inv_exd = (float)(1.0 / c.exd); a->g.f_inv_eyd = (float)(1.0 / c.eyd);
    a->g.f_nbx1 = (float)(c.nbx - 1); a->g.f_nby1 = (float)(c.nby - 1);
    a->origin_x = (R)ox; a->origin_y = (R)oy;
    a->n_replicas = c.n_replicas; a->cars_per_replica = c.n_cars > 0 ? c.n_cars : 1;
    a->lights_per_replica = c.n_lights > 0 ? c.n_lights : 1; a->faces_per_replica = c.n_faces;
    a->n_cars = c.n_replicas * c.n_cars; a->n_lights = c.n_replicas * c.n_lights;
    a->car_blocks = (a->n_cars + step_block_threads() - 1) / step_block_threads();
    a->agent_car = c.agent_car; a->step_index = (int32_t)s->steps;
    a->speed_limit = (R)c.speed_limit; a->stop_distance = (R)c.stop_distance; a->free_distance = (R)c.free_distance;
    a->default_acceleration = (R)c.default_acceleration;
    a->inv_free_distance = (R)(1.0 / c.free_distance);
    a->log_table = s->buf<char>(B_LOG_TABLE);
    a->log_scale = (R)(128.0 / (c.free_distance - c.stop_distance));
    a->log_stop = (R)std::log(c.stop_distance);
    a->inv_log_free_over_stop = (R)(1.0 / s->log_free_over_stop);
    a->vel = s->buf<R2>(B_VEL);
    a->route_time = s->buf<R>(B_ROUTE_TIME);
    a->cursor = s->buf<int32_t>(B_CURSOR);
    a->path_end = s->buf<int32_t>(B_PATH_END);
    a->path_eff_end = s->buf<int32_t>(B_PATH_EFF_END);
    a->path_xy = s->buf<R2>(B_PATH_XY);
    a->pt_curv = s->buf<R2>(B_PT_CURV);
    a->dest_xy = s->buf<R2>(B_DEST_XY);
    a->head_xy = s->buf<R2>(B_HEAD_XY);
    a->head_curv = s->buf<R2>(B_HEAD_CURV);
    a->d_node = s->buf<R>(B_D_NODE);
    a->d_car = s->buf<R>(B_D_CAR);
    a->d_light = s->buf<R>(B_D_LIGHT);
    a->cell = s->buf<int32_t>(B_CELL);
    a->cell_prev = s->buf<int32_t>(B_CELL_PREV);
    a->leader = s->buf<int32_t>(B_LEADER);
    a->light_xy = s->buf<R2>(B_LIGHT_XY);
    a->face_ptr = s->buf<int32_t>(B_FACE_PTR);
    a->face_vec = s->buf<R2>(B_FACE_VEC);
    a->face_unit = s->buf<R2>(B_FACE_UNIT);
    a->cell_light_ptr = s->buf<int32_t>(B_CELL_LIGHT_PTR);
    a->cell_light_idx = s->buf<int32_t>(B_CELL_LIGHT_IDX);
    a->cell_stat = s->buf<CellStatT<R>>(B_CELL_STAT);
    a->light_cellinfo = s->buf<int2>(B_LIGHT_CELLINFO);
    a->cell_dyn = s->buf<CellDynT>(B_CELL_DYN);
    a->tab_cur_k = s->cur_tab; a->tab_next_k = (s->cur_tab + 1) % 3; a->tab_clear_k = (s->cur_tab + 2) % 3;
    a->go_k_cur = s->cur_go;
    a->switch_time = s->buf<double>(B_SWITCH_TIME);
    a->pos_in = s->buf<R2>(s->cur_pos ? B_POS1 : B_POS0);
    a->pos_out = s->buf<R2>(s->cur_pos ? B_POS0 : B_POS1);
    a->go_cur = s->buf<uint8_t>(s->cur_go ? B_GO1 : B_GO0);
    a->go_next = s->buf<uint8_t>(s->cur_go ? B_GO0 : B_GO1);
    a->t_cur = s->buf<double>(s->cur_t ? B_T1 : B_T0);
    a->t_next = s->buf<double>(s->cur_t ? B_T0 : B_T1);
    a->done = s->buf<int32_t>(B_DONE);
    a->trip_time = s->buf<double>(B_TRIP_TIME);
    a->trip_step = s->buf<int32_t>(B_TRIP_STEP);
}

// TrafficLights.update (cars.py:123-133) as its own launch
void tick_lights(tb2_sim* s, double dt, cudaStream_t stream) {
    const tb2_config& c = s->cfg;
    // (a stale go word of the CURRENT buffer is harmless here: the tick reads the face bytes and writes the next word)
    launch_lights_tick(c.n_replicas * c.n_lights, c.n_lights > 0 ? c.n_lights : 1, c.n_faces, s->n_cells, dt,
                       s->buf<double>(B_SWITCH_TIME), s->buf<int32_t>(B_FACE_PTR),
                       s->buf<uint8_t>(s->cur_go ? B_GO1 : B_GO0), s->buf<uint8_t>(s->cur_go ? B_GO0 : B_GO1),
                       s->buf<double>(s->cur_t ? B_T1 : B_T0), s->buf<double>(s->cur_t ? B_T0 : B_T1),
                       s->buf<int2>(B_LIGHT_CELLINFO), s->buf<CellDynT>(B_CELL_DYN), s->cur_go ^ 1, stream);
    s->cur_go ^= 1; s->cur_t ^= 1; s->launches++;
}

// the caller replaced the face states (Cars.update with a foreign lights DataFrame): refresh the per-bin go words
void sync_go_words(tb2_sim* s, cudaStream_t stream) {
    if (!s->go_dirty) return;
    const tb2_config& c = s->cfg;
    launch_rebuild_go_words(c.n_replicas * c.n_lights, c.n_lights > 0 ? c.n_lights : 1, c.n_faces, s->n_cells,
                            s->buf<int32_t>(B_FACE_PTR), s->buf<uint8_t>(s->cur_go ? B_GO1 : B_GO0),
                            s->buf<int2>(B_LIGHT_CELLINFO), s->buf<CellDynT>(B_CELL_DYN), s->cur_go, stream);
    s->go_dirty = false;
    s->launches++;
}

template <typename R>
void step_cars_t(tb2_sim* s, double dt, int tick, int aux, cudaStream_t stream) {
    StepArgsT<R> a{};
    fill_args<R>(s, &a);
    a.dt = (R)dt;
    a.dt64 = dt;
    a.tick = tick;
    a.write_aux = aux;
    // big worlds (at least two tiles per resident CTA): the software-pipelined persistent kernel
    a.pipe_ctas = 0;
    a.pipe_tma = s->pipe_tma ? 1 : 0;
    if (s->allow_pipe && s->n_sms > 0) {
        const int resident = s->n_sms * pipe_ctas_per_sm();
        const int tiles = (a.n_cars + pipe_block_threads() - 1) / pipe_block_threads();
        if (tiles >= 2 * resident) a.pipe_ctas = resident;
        else if (s->force_pipe && tiles > 0) a.pipe_ctas = tiles < resident ? tiles : resident;
    }
    a.skip_insert = s->cfg.neighbor_mode == TB2_NEIGHBOR_SORT;
    a.freeze_done = s->freeze_done ? 1 : 0;
    // the live-replica list only means something to the pipelined ensemble kernel; every other form walks all replicas
    const bool walk_live = s->n_live > 0 && s->freeze_done && a.pipe_ctas > 0 && !a.pipe_tma;
    a.live_list = walk_live ? s->buf<int32_t>(B_LIVE_LIST) : nullptr;
    a.done_v = s->buf<int32_t>(B_DONE_V);
    a.n_live = walk_live ? s->n_live : 0;
    a.n_walk = walk_live ? s->n_live * s->cfg.n_cars : a.n_cars;
    launch_step(a, stream);
    if (a.skip_insert)   // north_star's pipeline: radix sort by bin + segment heads -> the same table
        s->launches += neighbor_sort_build(a.n_cars, a.cars_per_replica, s->n_cells, key_bits_of(s->cfg), a.cell,
                                           s->buf<int32_t>(B_SORT_KEYS_IN), s->buf<int32_t>(B_SORT_VALS_IN),
                                           s->buf<int32_t>(B_SORT_KEYS_OUT), s->buf<int32_t>(B_SORT_VALS_OUT),
                                           s->buf<char>(B_SORT_TEMP), s->lay.bytes[B_SORT_TEMP], a.cell_dyn, a.tab_next_k, stream);
}

// The cluster-resident form is the low-latency form; the batched HBM-streaming kernels are the throughput form (a step
// barrier per replica leaves SMs idle: 1,024 replicas x 2k cars run 2x faster batched).  Measured costs per step:
// ~6 us per wave of co-resident clusters (one CTA per SM) vs max(10 us, 50 us per million cars) batched.
bool resident_ok(const tb2_sim* s) {
    if (!(s->allow_resident && s->allow_persistent && s->cfg.neighbor_mode == TB2_NEIGHBOR_ATOMIC && s->cfg.n_cars > 0))
        return false;
    const int cl = resident_fits(s->cfg.n_cars, s->cfg.n_lights, s->n_cells, s->cfg.precision == TB2_F32);
    if (cl <= 0) return false;
    const int sms = s->n_sms > 0 ? s->n_sms : 1;
    const int64_t ctas = (int64_t)cl * s->cfg.n_replicas;
    const double waves = (double)((ctas + sms - 1) / sms);
    const double batched_us = std::fmax(10.0, 50.0e-6 * (double)s->cfg.n_cars * s->cfg.n_replicas);
    return waves * 6.0 <= batched_us;
}

// All n_steps of a batch in one launch of the persistent cluster kernel (small worlds only).
template <typename R>
cudaError_t step_persistent_t(tb2_sim* s, double dt, int n_steps, int order, cudaStream_t stream, bool resident = false,
                              bool stop_when_done = false) {
    using R2 = typename real2<R>::type;
    StepArgsT<R> a{};
    fill_args<R>(s, &a);
    a.dt = (R)dt;
    a.dt64 = dt;
    a.skip_insert = 0;
    PersistArgsT<R> pa{};
    pa.pos[0] = s->buf<R2>(B_POS0); pa.pos[1] = s->buf<R2>(B_POS1);
    pa.go[0] = s->buf<uint8_t>(B_GO0); pa.go[1] = s->buf<uint8_t>(B_GO1);
    pa.t[0] = s->buf<double>(B_T0); pa.t[1] = s->buf<double>(B_T1);
    pa.cur_pos = s->cur_pos; pa.cur_tab = s->cur_tab; pa.cur_go = s->cur_go; pa.cur_t = s->cur_t;
    pa.n_steps = n_steps; pa.order = order; pa.aux_last = s->cfg.write_aux;
    pa.stop_when_done = stop_when_done ? 1 : 0;
    return resident ? launch_resident(a, pa, stream) : launch_persistent(a, pa, stream);
}

// Cars.update (cars.py:61-95), optionally with the lights tick fused into the trailing blocks of the same launch.
// aux: also write the columns nobody can observe between steps (velocity, distances, leader, start-of-step bin).
void step_cars(tb2_sim* s, double dt, int tick, int aux, cudaStream_t stream) {
    sync_go_words(s, stream);
    if (s->cfg.precision == TB2_F64) step_cars_t<double>(s, dt, tick, aux, stream);
    else step_cars_t<float>(s, dt, tick, aux, stream);
    s->launches++;
    s->cur_pos ^= 1; s->cur_tab = (s->cur_tab + 1) % 3;
    if (tick) { s->cur_go ^= 1; s->cur_t ^= 1; }
    s->steps++;
}

}  // namespace

extern "C" {

int tb2_abi_version(void) { return TB2_ABI_VERSION; }
const char* tb2_last_error(void) { return g_err.c_str(); }

size_t tb2_arena_bytes(const tb2_config* cfg) {
    std::string why;
    if (!config_ok(cfg, &why)) { g_err = why; return 0; }
    return make_layout(*cfg).total;
}

int tb2_create(const tb2_config* cfg, void* arena_dev, size_t arena_bytes, tb2_sim** out) {
    if (!out) return fail(TB2_ERR_INVALID, "null out pointer");
    *out = nullptr;
    std::string why;
    if (!config_ok(cfg, &why)) return fail(TB2_ERR_INVALID, why);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(TB2_ERR_NO_DEVICE, "no CUDA device available - traffic_b200 has no CPU path");
    }
    tb2_sim* s = new (std::nothrow) tb2_sim();
    if (!s) return fail(TB2_ERR_INVALID, "out of host memory");
    s->cfg = *cfg;
    s->lay = make_layout(*cfg);
    s->n_cells = (cfg->nbx + 1) * (cfg->nby + 1);
    s->log_free_over_stop = std::log(cfg->free_distance / cfg->stop_distance);  // host libm, simulation.py:180
    if (const char* np = std::getenv("TB2_NO_PERSISTENT")) s->allow_persistent = !(np[0] == '1');
    if (const char* np = std::getenv("TB2_NO_PIPE")) s->allow_pipe = !(np[0] == '1');
    if (const char* np = std::getenv("TB2_NO_RESIDENT")) s->allow_resident = !(np[0] == '1');
    if (const char* np = std::getenv("TB2_FORCE_PIPE")) s->force_pipe = (np[0] == '1');
    if (const char* np = std::getenv("TB2_PIPE_TMA")) s->pipe_tma = (np[0] == '1');
    if (s->force_pipe) { s->allow_persistent = false; s->allow_resident = false; }
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&s->n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            (void)cudaGetLastError();
            s->n_sms = 0;
        }
    }
    if (arena_dev) {
        if (arena_bytes < s->lay.total) { delete s; return fail(TB2_ERR_INVALID, "arena too small (see tb2_arena_bytes)"); }
        if (reinterpret_cast<uintptr_t>(arena_dev) % kAlign) { delete s; return fail(TB2_ERR_INVALID, "arena must be 256-byte aligned"); }
        s->arena = static_cast<char*>(arena_dev);
    } else {
        e = cudaMalloc(reinterpret_cast<void**>(&s->arena), s->lay.total);
        if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMalloc(arena)"); }
        s->owns_arena = true;
    }
    *out = s;
    return TB2_OK;
}

int tb2_destroy(tb2_sim* s) {
    if (!s) return TB2_OK;
    if (s->frame_pending) cudaEventSynchronize(s->frame_done);
    if (s->frame_ready) cudaEventDestroy(s->frame_ready);
    if (s->frame_done) cudaEventDestroy(s->frame_done);
    if (s->faces_done) cudaEventDestroy(s->faces_done);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->owns_arena && s->arena) cudaFree(s->arena);
    delete s;
    return TB2_OK;
}

int tb2_field_info(const tb2_sim* s, int field, size_t* n_elems, size_t* elem_bytes) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    int b; size_t eb, cnt;
    if (int rc = field_buf(s, field, &b, &eb, &cnt)) return rc;
    if (n_elems) *n_elems = cnt;
    if (elem_bytes) *elem_bytes = eb;
    return TB2_OK;
}

void* tb2_field_ptr(const tb2_sim* s, int field) {
    if (!s) return nullptr;
    int b; size_t eb, cnt;
    if (field_buf(s, field, &b, &eb, &cnt)) return nullptr;
    return s->arena + s->lay.off[b];
}

int tb2_upload(tb2_sim* s, int field, const void* src, size_t bytes, void* stream) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    int b; size_t eb, cnt;
    if (int rc = field_buf(s, field, &b, &eb, &cnt)) return rc;
    if (bytes != eb * cnt) {
        char m[160];
        snprintf(m, sizeof m, "tb2_upload(field %d): got %zu bytes, field holds %zu", field, bytes, eb * cnt);
        return fail(TB2_ERR_INVALID, m);
    }
    if (bytes && !src) return fail(TB2_ERR_INVALID, "null source");
    if (bytes) CU(cudaMemcpyAsync(s->arena + s->lay.off[b], src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    if (field == TB2_POS || field == TB2_CURSOR || field == TB2_PATH_END || field == TB2_PATH_XY ||
        field == TB2_PATH_EFF_END || field == TB2_DEST_XY ||
        field == TB2_FACE_VEC || field == TB2_LIGHT_XY || field == TB2_FACE_PTR || field == TB2_CELL_LIGHT_PTR ||
        field == TB2_CELL_LIGHT_IDX) s->prepared = false;
    if (field == TB2_FACE_GO) s->go_dirty = true;
    return TB2_OK;
}

int tb2_download(const tb2_sim* s, int field, void* dst, size_t bytes, void* stream) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    int b; size_t eb, cnt;
    if (int rc = field_buf(s, field, &b, &eb, &cnt)) return rc;
    if (bytes != eb * cnt) {
        char m[160];
        snprintf(m, sizeof m, "tb2_download(field %d): got %zu bytes, field holds %zu", field, bytes, eb * cnt);
        return fail(TB2_ERR_INVALID, m);
    }
    if (bytes && !dst) return fail(TB2_ERR_INVALID, "null destination");
    if (bytes) CU(cudaMemcpyAsync(dst, s->arena + s->lay.off[b], bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return TB2_OK;
}

int tb2_download_range(const tb2_sim* s, int field, size_t first, size_t count, void* dst, size_t bytes, void* stream) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    int b; size_t eb, cnt;
    if (int rc = field_buf(s, field, &b, &eb, &cnt)) return rc;
    if (first > cnt || count > cnt - first) return fail(TB2_ERR_INVALID, "tb2_download_range: range outside the field");
    if (bytes != eb * count) return fail(TB2_ERR_INVALID, "tb2_download_range: byte count does not match the range");
    if (bytes && !dst) return fail(TB2_ERR_INVALID, "null destination");
    if (bytes) CU(cudaMemcpyAsync(dst, s->arena + s->lay.off[b] + first * eb, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return TB2_OK;
}

int tb2_prepare(tb2_sim* s, void* stream_) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    cudaStream_t stream = (cudaStream_t)stream_;
    CU(cudaMemsetAsync(s->buf<char>(B_DONE), 0, s->lay.bytes[B_DONE], stream));
    CU(cudaMemsetAsync(s->buf<char>(B_TRIP_TIME), 0, s->lay.bytes[B_TRIP_TIME], stream));
    CU(cudaMemsetAsync(s->buf<char>(B_TRIP_STEP), 0, s->lay.bytes[B_TRIP_STEP], stream));
    s->steps = 0;
    {   // the index arrays the caller uploaded are dereferenced by every kernel: validate them once, here
        const tb2_config& c = s->cfg;
        int* bad_dev = reinterpret_cast<int*>(s->buf<char>(B_DONE));   // (zeroed just above; zeroed again below)
        launch_validate(c.n_replicas * c.n_cars, (long long)c.n_path_points, c.n_lights, c.n_faces, s->n_cells, c.n_cell_lights,
                        s->buf<int32_t>(B_CURSOR), s->buf<int32_t>(B_PATH_END), s->buf<int32_t>(B_PATH_EFF_END),
                        s->buf<int32_t>(B_FACE_PTR), s->buf<int32_t>(B_CELL_LIGHT_PTR), s->buf<int32_t>(B_CELL_LIGHT_IDX),
                        bad_dev, stream);
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaMemsetAsync(bad_dev, 0, sizeof(int), stream));
        s->launches++;
        if (bad) {
            char m[256];
            snprintf(m, sizeof m, "tb2_prepare: inconsistent index arrays (mask 0x%x: 1 cursor/path_end range, 2 path_eff_end, "
                     "4 face_ptr, 8 more than 32767 faces per light or lights per bin, 16 cell_light_ptr, 32 cell_light_idx)", bad);
            return fail(TB2_ERR_INVALID, m);
        }
    }
    if (s->cfg.precision == TB2_F64) {
        StepArgs64 a{};
        fill_args<double>(s, &a);
        launch_prepare(a, s->buf<double2>(B_PT_CURV), s->buf<double2>(B_FACE_UNIT), s->cfg.n_faces, stream);
    } else {
        StepArgs32 a{};
        fill_args<float>(s, &a);
        launch_prepare(a, s->buf<float2>(B_PT_CURV), s->buf<float2>(B_FACE_UNIT), s->cfg.n_faces, stream);
    }
    s->launches += 3 + (s->cfg.n_cars > 0 ? 2 : 0) + (s->cfg.n_faces > 0 ? 1 : 0) + (s->cfg.n_lights > 0 ? 2 : 0);
    CU(cudaGetLastError());
    s->go_dirty = false;
    s->prepared = true;
    return TB2_OK;
}

int tb2_lights_update(tb2_sim* s, double dt, void* stream) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    tick_lights(s, dt, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return TB2_OK;
}

int tb2_cars_update(tb2_sim* s, double dt, void* stream) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    if (!s->prepared) return fail(TB2_ERR_STATE, "tb2_cars_update before tb2_prepare (state was uploaded since)");
    step_cars(s, dt, 0, s->cfg.write_aux, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return TB2_OK;
}

int tb2_step(tb2_sim* s, double dt, int32_t n_steps, int32_t order, void* stream_) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    if (n_steps < 0) return fail(TB2_ERR_INVALID, "negative n_steps");
    if (order != TB2_ORDER_CARS_LIGHTS && order != TB2_ORDER_LIGHTS_CARS) return fail(TB2_ERR_INVALID, "bad order");
    if (!s->prepared) return fail(TB2_ERR_STATE, "tb2_step before tb2_prepare (state was uploaded since)");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_steps >= 2 && resident_ok(s)) {
        // small world / replica ensemble: one cluster per replica, the world stays on chip for the whole batch
        sync_go_words(s, stream);
        const cudaError_t pe = s->cfg.precision == TB2_F64 ? step_persistent_t<double>(s, dt, n_steps, order, stream, true)
                                                           : step_persistent_t<float>(s, dt, n_steps, order, stream, true);
        if (pe != cudaSuccess) {   // cluster launch refused (shared / partitioned device): never try again
            (void)cudaGetLastError();
            s->allow_resident = false;
            return tb2_step(s, dt, n_steps, order, stream_);
        }
        s->launches++;
        s->cur_pos ^= (n_steps & 1);
        s->cur_tab = (s->cur_tab + n_steps) % 3;
        s->cur_go ^= (n_steps & 1);      // n_steps ticks in either order
        s->cur_t ^= (n_steps & 1);
        s->steps += n_steps;
        CU(cudaGetLastError());
        return TB2_OK;
    }
    if (s->allow_persistent && n_steps >= 2 && s->cfg.neighbor_mode == TB2_NEIGHBOR_ATOMIC &&
        (int64_t)s->cfg.n_cars * s->cfg.n_replicas <= persistent_max_cars(s->cfg.precision == TB2_F32)) {
        // small world: the whole batch is one launch of the persistent cluster kernel
        sync_go_words(s, stream);
        const cudaError_t pe = s->cfg.precision == TB2_F64 ? step_persistent_t<double>(s, dt, n_steps, order, stream)
                                                           : step_persistent_t<float>(s, dt, n_steps, order, stream);
        if (pe != cudaSuccess) {   // cooperative launch refused (shared / partitioned device, capacity queried for another
            (void)cudaGetLastError();   // variant, ...): fall back to one launch per step for good
            s->allow_persistent = false;
            return tb2_step(s, dt, n_steps, order, stream_);
        }
        s->launches++;
        s->cur_pos ^= (n_steps & 1);
        s->cur_tab = (s->cur_tab + n_steps) % 3;
        s->cur_go ^= (n_steps & 1);      // n_steps ticks in either order
        s->cur_t ^= (n_steps & 1);
        s->steps += n_steps;
        CU(cudaGetLastError());
        return TB2_OK;
    }
    for (int32_t k = 0; k < n_steps; ++k) {
        // lights-first order: L (C L) (C L) ... C - the first tick is its own launch, the rest ride in the step kernel
        if (order == TB2_ORDER_LIGHTS_CARS && k == 0) tick_lights(s, dt, stream);
        const int tick = (order == TB2_ORDER_CARS_LIGHTS) ? 1 : (k < n_steps - 1 ? 1 : 0);
        step_cars(s, dt, tick, (s->cfg.write_aux && k == n_steps - 1) ? 1 : 0, stream);
    }
    CU(cudaGetLastError());
    return TB2_OK;
}

int tb2_rollout(tb2_sim* s, double dt, int32_t max_steps, int32_t order, void* stream_) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    if (max_steps < 1) return fail(TB2_ERR_INVALID, "max_steps must be >= 1");
    if (order != TB2_ORDER_CARS_LIGHTS && order != TB2_ORDER_LIGHTS_CARS) return fail(TB2_ERR_INVALID, "bad order");
    if (s->cfg.agent_car < 0) return fail(TB2_ERR_INVALID, "tb2_rollout needs an agent car (tb2_config.agent_car)");
    if (!s->prepared) return fail(TB2_ERR_STATE, "tb2_rollout before tb2_prepare (state was uploaded since)");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!resident_ok(s)) {
        // large ensembles: batched launches; the cars of an arrived replica stop moving, and every `chunk` steps the done
        // flags (4 bytes per replica) are read back to leave as soon as every rollout is over
        const int R = s->cfg.n_replicas, chunk = 256;
        int32_t* done_h = nullptr;   // [0, R): done flags read back; [R, 2R): the live-replica list on its way to the device
        CU(cudaMallocHost(reinterpret_cast<void**>(&done_h), sizeof(int32_t) * 2 * (size_t)R));
        int32_t* live_h = done_h + R;
        s->freeze_done = true;
        s->n_live = 0;
        int rc = TB2_OK;
        const bool was_persistent = s->allow_persistent;
        s->allow_persistent = false;   // per-step launches (the cooperative kernel has no frozen-replica path)
        for (int32_t done_steps = 0; done_steps < max_steps && rc == TB2_OK;) {
            const int32_t n = max_steps - done_steps < chunk ? max_steps - done_steps : chunk;
            rc = tb2_step(s, dt, n, order, stream_);
            done_steps += n;
            if (rc != TB2_OK) break;
            if (cudaMemcpyAsync(done_h, s->buf<int32_t>(B_DONE), sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                cudaStreamSynchronize(stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "tb2_rollout readback"); break; }
            // the replicas still running: from now on the pipelined kernel walks only their cars (a replica that arrives
            // before the next read-back is frozen through its done_v flag)
            int n_live = 0;
            for (int r = 0; r < R; ++r) if (done_h[r] == 0) live_h[n_live++] = r;
            if (n_live == 0) break;
            if (n_live < R) {
                if (cudaMemcpyAsync(s->buf<int32_t>(B_LIVE_LIST), live_h, sizeof(int32_t) * (size_t)n_live, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
                    cudaMemsetAsync(s->buf<int32_t>(B_DONE_V), 0, sizeof(int32_t) * (size_t)n_live, stream) != cudaSuccess ||
                    cudaStreamSynchronize(stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "tb2_rollout live list"); break; }
                s->n_live = n_live;
            }
        }
        s->allow_persistent = was_persistent;
        s->freeze_done = false;
        s->n_live = 0;
        cudaFreeHost(done_h);
        return rc;
    }
    sync_go_words(s, stream);
    const cudaError_t pe = s->cfg.precision == TB2_F64 ? step_persistent_t<double>(s, dt, max_steps, order, stream, true, true)
                                                       : step_persistent_t<float>(s, dt, max_steps, order, stream, true, true);
    CU(pe);
    s->launches++;
    s->cur_pos ^= (max_steps & 1);
    s->cur_tab = (s->cur_tab + max_steps) % 3;
    s->cur_go ^= (max_steps & 1);
    s->cur_t ^= (max_steps & 1);
    s->steps += max_steps;
    CU(cudaGetLastError());
    return TB2_OK;
}

int tb2_step_host(tb2_sim* s, double dt, int32_t n_steps, int32_t order, const uint8_t* face_go_in,
                  void* frame_xy_out, uint8_t* face_go_out, void* stream) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    const size_t F = (size_t)s->cfg.n_faces * s->cfg.n_replicas, N = (size_t)s->cfg.n_cars * s->cfg.n_replicas;
    if (face_go_in) if (int rc = tb2_upload(s, TB2_FACE_GO, face_go_in, F, stream)) return rc;
    if (int rc = tb2_step(s, dt, n_steps, order, stream)) return rc;
    if (frame_xy_out) if (int rc = tb2_download(s, TB2_POS, frame_xy_out, N * 2 * s->real_bytes(), stream)) return rc;
    if (face_go_out) if (int rc = tb2_download(s, TB2_FACE_GO, face_go_out, F, stream)) return rc;
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return TB2_OK;
}

// Pipelined variant: returns as soon as the work is enqueued.  The positions / face states after the n_steps are
// snapshotted on the device and copied to the host buffers on a private copy stream, so the caller can enqueue the next
// batch of steps while the frame is in flight; tb2_frame_wait() blocks until the LAST requested frame has landed.
int tb2_step_host_async(tb2_sim* s, double dt, int32_t n_steps, int32_t order, const uint8_t* face_go_in,
                        void* frame_xy_out, uint8_t* face_go_out, void* stream_) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t F = (size_t)s->cfg.n_faces * s->cfg.n_replicas, N = (size_t)s->cfg.n_cars * s->cfg.n_replicas;
    if (!s->copy_stream) {
        CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->frame_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->frame_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->faces_done, cudaEventDisableTiming));
    }
    if (face_go_in) if (int rc = tb2_upload(s, TB2_FACE_GO, face_go_in, F, stream)) return rc;
    if (int rc = tb2_step(s, dt, n_steps, order, stream)) return rc;
    if (frame_xy_out || face_go_out) {
        // the staging buffers may still be read by the previous frame's host copy
        if (s->frame_pending) CU(cudaStreamWaitEvent(stream, s->frame_done, 0));
        if (frame_xy_out) CU(cudaMemcpyAsync(s->buf<char>(B_FRAME_STAGE), tb2_field_ptr(s, TB2_POS), N * 2 * s->real_bytes(),
                                             cudaMemcpyDeviceToDevice, stream));
        if (face_go_out && F) CU(cudaMemcpyAsync(s->buf<char>(B_GO_STAGE), tb2_field_ptr(s, TB2_FACE_GO), F,
                                                 cudaMemcpyDeviceToDevice, stream));
        CU(cudaEventRecord(s->frame_ready, stream));
        CU(cudaStreamWaitEvent(s->copy_stream, s->frame_ready, 0));
        // the (small) face states go first so that a caller that feeds them back as the next input waits microseconds
        if (face_go_out && F) CU(cudaMemcpyAsync(face_go_out, s->buf<char>(B_GO_STAGE), F, cudaMemcpyDeviceToHost,
                                                 s->copy_stream));
        CU(cudaEventRecord(s->faces_done, s->copy_stream));
        if (frame_xy_out) CU(cudaMemcpyAsync(frame_xy_out, s->buf<char>(B_FRAME_STAGE), N * 2 * s->real_bytes(),
                                             cudaMemcpyDeviceToHost, s->copy_stream));
        CU(cudaEventRecord(s->frame_done, s->copy_stream));
        s->frame_pending = true;
    }
    return TB2_OK;
}

int tb2_frame_wait(tb2_sim* s, int32_t faces_only) {
    if (!s) return fail(TB2_ERR_INVALID, "null sim");
    if (s->frame_pending) {
        if (faces_only) { CU(cudaEventSynchronize(s->faces_done)); return TB2_OK; }
        CU(cudaEventSynchronize(s->frame_done));
        s->frame_pending = false;
    }
    return TB2_OK;
}

int tb2_debug_bounds_enabled(void) { return debug_bounds_enabled(); }
unsigned int tb2_debug_bounds_violations(void) { return debug_bounds_violations(); }

int64_t tb2_launch_count(const tb2_sim* s) { return s ? s->launches : 0; }
int64_t tb2_step_count(const tb2_sim* s) { return s ? s->steps : 0; }

}  // extern "C"
