// Small worlds and replica ensembles: the whole world of ONE replica lives on chip for a whole batch of steps.
//
// One thread-block cluster per replica (<= 16 CTAs x 256 threads, so <= 4096 cars - BASELINE's 2k-car Manhattan world and
// every replica of the learn.py-style rollout ensemble, environment.py:97-173).  For the duration of the launch
//   * every car's state (position, route cursor, cached head point and curvature constants, route time, bin) stays in the
//     registers of its thread,
//   * the positions other cars look at (the candidate leader) are ping-pong buffers in the shared memory of the CTA that
//     owns the car, read through distributed shared memory (DSMEM),
//   * the per-bin records - the three rotating (first, second) car tables, the go words, the static first-light record -
//     are distributed over the cluster's shared memory by bin id (bin c lives in CTA c % cluster_size) and are read /
//     atomically updated through DSMEM,
//   * the light timers, switch times and face masks stay in the registers of one thread per light,
//   * steps are separated by ONE hardware cluster barrier (release / acquire over shared::cluster) - no global-memory
//     round trip, no L1 invalidation, no launch.
// Global memory is touched at the start (load) and at the end (write back, aux columns of the last step), by the rare
// general paths (a second light in a bin, exact predicate fall-backs, the step in which a car crosses a route point) and
// for the per-face go bytes (a handful of byte stores per switching light).  A grid of R clusters advances R replicas;
// with stop_when_done a cluster leaves as soon as its agent has arrived (FrontView.end_of_route, navigation.py:113-128),
// so rollouts of different length balance themselves over the SMs.
//
// The arithmetic is the same seg_view / seg_detect / seg_move code as every other form of the step: results are
// bit-identical to one launch per step (tests/test_gpu_parity.py, tests/test_ensemble.py).
#pragma once
#include "step_kernel.cuh"

#define TB2_RES_BLOCK 256        // largest CTA of the resident form (worlds up to 16 x 256 cars)
#define TB2_RES_MAX_CTAS 16

namespace tb2k {

namespace cg = cooperative_groups;

// operand provider: everything in registers, results stay in registers (position also goes to the local ping-pong slot)
template <typename R>
struct OwnResident {
    using R2 = typename real2<R>::type;
    const StepArgsT<R>* a;
    R2 p, head, hc_, t;
    R2 nxt, nxt_hc, dst;   // the route point after the head + its curvature constants, the destination: a crossing step
                           // (some car of the world crosses in almost every step) must not make its warp the straggler
                           // of the cluster barrier with a chain of global loads
    R rt_;
    int cur, end, eff, c, sec_;
    R2* pos_out_slot;
    const CellStatT<R>* stat_ptr;   // DSMEM address of the bin's static record
    bool arrived;
    __device__ __forceinline__ R2 target() const { return t; }
    __device__ __forceinline__ int3 route() const { return make_int3(cur, end, eff); }
    __device__ __forceinline__ R2 hc() const { return hc_; }
    __device__ __forceinline__ R rt() const { return rt_; }
    __device__ __forceinline__ int cell() const { return c; }
    __device__ __forceinline__ int sec() const { return sec_; }
    __device__ __forceinline__ int rep(int r) const { return r; }
    __device__ __forceinline__ R2 next_curv(int) const { return nxt_hc; }
    __device__ __forceinline__ R2 dest(int) const { return dst; }
    __device__ __forceinline__ R2 crossing_target_value() const { return (end - cur < 2) ? dst : nxt; }
    // issued now, consumed at the next crossing (thousands of cycles away): never waited for
    __device__ __forceinline__ void load_next() {
        if (cur + 1 < end) { nxt = a->path_xy[cur + 1]; nxt_hc = a->pt_curv[cur + 1]; }
    }
    __device__ __forceinline__ void after_cross(int, int) { load_next(); }
    __device__ __forceinline__ CellStatT<R> stat() const { return *stat_ptr; }
    __device__ __forceinline__ void store_route_time(int, R v) { rt_ = v; }
    __device__ __forceinline__ void store_cursor(int, int v) { cur = v; }
    __device__ __forceinline__ void store_head(int, R2 v) { head = v; }
    __device__ __forceinline__ void store_head_curv(int, R2 v) { hc_ = v; }
    __device__ __forceinline__ void store_pos(int, R2 v) { *pos_out_slot = v; p = v; }
    __device__ __forceinline__ void store_cell(int, int v) { c = v; }
    __device__ __forceinline__ void on_done() { arrived = true; }
};

// Dynamic shared memory of one CTA: [pos 2 x BLOCK][log table][per-slot dyn records][per-slot static records][flags]
template <typename R>
struct ResidentLayout {
    using R2 = typename real2<R>::type;
    static __host__ __device__ size_t pos_off() { return 0; }
    static __host__ __device__ size_t log_off() { return 2 * TB2_RES_BLOCK * sizeof(R2); }   // (sized for the largest CTA)
    static __host__ __device__ size_t dyn_off() { return log_off() + TB2_LOG_N * sizeof(LogEntry); }
    static __host__ __device__ size_t stat_off(int slots) { return dyn_off() + (size_t)slots * sizeof(CellDynT); }
    static __host__ __device__ size_t flag_off(int slots) { return stat_off(slots) + (size_t)slots * sizeof(CellStatT<R>); }
    static __host__ __device__ size_t bytes(int slots) { return flag_off(slots) + 16; }
};

// B = threads per CTA: 128 for worlds up to 2,048 cars (16 CTAs, one warp per SM sub-partition: the step is a latency
// chain, spreading it over more SMs shortens it - 5.4 vs 5.9 us per step on the 2k-car world), 256 beyond.
template <typename R, bool ENS, int B>
__global__ void __launch_bounds__(B, 1)
resident_kernel(const __grid_constant__ StepArgsT<R> a, const __grid_constant__ PersistArgsT<R> pa) {
    using R2 = typename real2<R>::type;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    cg::cluster_group cl = cg::this_cluster();
    const int CL = (int)cl.num_blocks();
    const int rank = (int)cl.block_rank();
    const int tid = (int)threadIdx.x;
    const int rep = (int)blockIdx.x / CL;
    const int slots = (a.g.n_cells + CL - 1) / CL;
    R2* s_pos = reinterpret_cast<R2*>(smem_raw + ResidentLayout<R>::pos_off());          // [2][B]
    LogEntry* s_log = reinterpret_cast<LogEntry*>(smem_raw + ResidentLayout<R>::log_off());
    CellDynT* s_dyn = reinterpret_cast<CellDynT*>(smem_raw + ResidentLayout<R>::dyn_off());
    CellStatT<R>* s_stat = reinterpret_cast<CellStatT<R>*>(smem_raw + ResidentLayout<R>::stat_off(slots));
    int* s_stop = reinterpret_cast<int*>(smem_raw + ResidentLayout<R>::flag_off(slots));

    // a rollout whose agent has already arrived is frozen (environment.Env.step returns at arrival)
    if (ENS && pa.stop_when_done && a.done[rep]) return;

    const int N = a.cars_per_replica;
    const int li = rank * B + tid;            // car index inside the replica
    const bool live = li < N;
    const int ibase = rep * N, cbase = rep * a.g.n_cells, fbase = rep * a.faces_per_replica;
    const int i = ibase + li;
    int cp = pa.cur_pos, ct = pa.cur_tab, cgo = pa.cur_go;
    int lp = 0;   // which local position buffer holds the current positions

    // ---- load: bin records of my slots, log table, my car, my light ----
    load_log_table(s_log, reinterpret_cast<const LogEntry*>(a.log_table));
    for (int sl = tid; sl < slots; sl += B) {
        const int c = sl * CL + rank;
        if (c < a.g.n_cells) {
            s_dyn[sl] = a.cell_dyn[cbase + c];
            s_stat[sl] = a.cell_stat[c];
        }
    }
    if (tid == 0) *s_stop = 0;
    OwnResident<R> own;
    own.a = &a;
    own.arrived = false;
    own.p = mk2(R(0), R(0)); own.head = own.p; own.hc_ = own.p; own.t = own.p; own.nxt = own.p; own.nxt_hc = own.p; own.dst = own.p;
    own.rt_ = R(0); own.cur = own.end = own.eff = own.c = 0; own.sec_ = TB2_EMPTY;
    if (live) {
        own.p = pa.pos[cp][i];
        own.head = a.head_xy[i]; own.hc_ = a.head_curv[i]; own.rt_ = a.route_time[i];
        own.cur = a.cursor[i]; own.end = a.path_end[i]; own.eff = a.path_eff_end[i]; own.c = a.cell[i];
        own.dst = a.dest_xy[i];
        own.nxt = own.dst; own.nxt_hc = own.hc_;
        own.load_next();
        s_pos[0 * B + tid] = own.p;   // local ping-pong buffer 0 holds the current positions
    }
    // one thread per light keeps the timer state in registers (TrafficLights.update, cars.py:123-133)
    const int L = a.lights_per_replica;
    const bool is_light = li < L && a.n_lights > 0;
    double l_sw = 0.0, l_t = *pa.t[pa.cur_t];
    int l_fb = 0, l_nf = 0;
    uint32_t l_mask = 0;
    int2 l_info = make_int2(-1, 0);
    if (is_light) {
        l_sw = a.switch_time[rep * L + li];
        l_fb = a.face_ptr[li];
        l_nf = a.face_ptr[li + 1] - l_fb;
        for (int f = 0; f < l_nf && f < 32; ++f) l_mask |= (pa.go[cgo][fbase + l_fb + f] ? 1u : 0u) << f;
        l_info = a.light_cellinfo[li];
    }
    auto dyn_of = [&](int c) { return cl.map_shared_rank(&s_dyn[c / CL], c % CL); };
    // The face bytes (global memory, read only by the rare general light path) and the bin's go word are double
    // buffered, so a light writes them for the two ticks that follow a switch (and at the first two ticks of the launch)
    // and stays silent otherwise: most steps leave no store in flight when the cluster barrier's release fence runs.
    int l_pending = 2;
    auto tick = [&]() {   // all threads advance the clock identically; light threads flip their faces
        l_t = __dadd_rn(l_t, a.dt64);
        if (is_light) {
            double r = fmod(l_t, l_sw);                       // numpy float remainder
            if (r != 0.0 && ((l_sw < 0.0) != (r < 0.0))) r = __dadd_rn(r, l_sw);
            const bool sw = ((fabs(r) <= __dadd_rn(1.0e-8, __dmul_rn(1.0e-4, fabs(r)))) && isfinite(r)) || (r == 0.0);
            if (sw) l_pending = 2;
            if (l_nf > 32) l_pending = 2;   // more faces than mask bits: the bytes in global memory are the state
            if (l_pending > 0) {
                --l_pending;
                uint8_t* gn = pa.go[cgo ^ 1] + fbase + l_fb;
                if (l_nf <= 32) {
                    if (sw) l_mask ^= (l_nf == 32 ? 0xffffffffu : ((1u << l_nf) - 1u));
                    for (int f = 0; f < l_nf; ++f) gn[f] = (uint8_t)((l_mask >> f) & 1u);
                } else {
                    const uint8_t* gc = pa.go[cgo] + fbase + l_fb;
                    for (int f = 0; f < l_nf; ++f) gn[f] = sw ? (uint8_t)!gc[f] : gc[f];
                    l_mask = 0;
                    for (int f = 0; f < 8; ++f) l_mask |= (gn[f] ? 1u : 0u) << f;
                }
                if (l_info.x >= 0)
                    dyn_of(l_info.x)->w[TB2_DYN_GO + (cgo ^ 1)] = (int32_t)((uint32_t)l_info.y | (l_mask & 0xffu));
            }
        }
        cgo ^= 1;
    };
    cl.sync();

    int n_done = 0;   // steps executed
    if (pa.order == TB2K_ORDER_LIGHTS_CARS) {               // L (C L) (C L) ... C
        tick();
        cl.sync();
    }
    for (int s = 0; s < pa.n_steps; ++s) {
        const bool do_tick = pa.order != TB2K_ORDER_LIGHTS_CARS || s < pa.n_steps - 1;
        // reset (locally) the table that the step after next will fill
        {
            const int k = (ct + 2) % 3;
            for (int sl = tid; sl < slots; sl += B) *reinterpret_cast<int2*>(&s_dyn[sl].w[2 * k]) = make_int2(TB2_EMPTY, TB2_EMPTY);
        }
        if (live) {
            const Rot<R> rot{nullptr, nullptr, pa.go[cgo], ct, (ct + 1) % 3, cgo, a.step_index + s};
            // the bin's records, through DSMEM
            const CellDynT* dyn = dyn_of(own.c);
            const int2 m = *reinterpret_cast<const int2*>(&dyn->w[2 * ct]);
            own.sec_ = dyn->w[2 * ((ct + 1) % 3) + 1];
            const uint32_t gw = (uint32_t)dyn->w[TB2_DYN_GO + cgo];
            own.stat_ptr = cl.map_shared_rank(&s_stat[own.c / CL], own.c % CL);
            const int j = (m.x != i) ? m.x : m.y;
            R2 o = mk2(R(0), R(0));
            TB2_BOUNDS(own.c >= 0 && own.c < a.g.n_cells, 8);
            TB2_BOUNDS(j == TB2_EMPTY || (j >= ibase && j < ibase + N && j != i), 9);
            if (j != TB2_EMPTY) {
                const int jl = j - ibase;
                o = *cl.map_shared_rank(&s_pos[lp * B + (jl % B)], jl / B);
            }
            own.pos_out_slot = &s_pos[(lp ^ 1) * B + tid];
            const bool crossed = view_crossed<R>(a, own.p, own.head, own.cur, own.eff);
            own.t = crossed ? own.crossing_target_value() : own.head;
            const View<R> v = view_build<R>(a, own.p, own.t, own.cur, own.end, own.eff, crossed);
            R d_car, d_light;
            int lead;
            seg_detect<R>(a, rot.go_cur, fbase, own.c, v, j, o, gw, own, d_car, lead, d_light);
            const bool last = s == pa.n_steps - 1;
            auto insert = [&](int rec, int ii, int sec) {
                if (ii > sec) return;
                int32_t* tn = &dyn_of(rec - cbase)->w[2 * ((ct + 1) % 3)];
                const int old = atomicMin(tn, ii);
                const int loser = old > ii ? old : ii;
                if (loser != TB2_EMPTY) atomicMin(tn + 1, loser);
            };
            if (pa.aux_last && last) seg_move<R, true, ENS>(a, rot, s_log, i, rep, v, d_car, lead, d_light, own, insert);
            else seg_move<R, false, ENS>(a, rot, s_log, i, rep, v, d_car, lead, d_light, own, insert);
            if (ENS && own.arrived) {
                for (int r = 0; r < CL; ++r) *cl.map_shared_rank(s_stop, r) = 1;
                own.arrived = false;
            }
        }
        if (do_tick) tick();
        cp ^= 1; lp ^= 1; ct = (ct + 1) % 3;
        ++n_done;
        cl.sync();
        if (ENS && pa.stop_when_done && *s_stop) break;
    }

    // ---- write back ----
    const bool complete = n_done == pa.n_steps;
    if (live) {
        if (complete) pa.pos[cp][i] = own.p;
        else { pa.pos[0][i] = own.p; pa.pos[1][i] = own.p; }   // frozen replica: both buffers
        a.head_xy[i] = own.head; a.head_curv[i] = own.hc_; a.route_time[i] = own.rt_;
        a.cursor[i] = own.cur; a.cell[i] = own.c;
    }
    for (int sl = tid; sl < slots; sl += B) {
        const int c = sl * CL + rank;
        if (c < a.g.n_cells) a.cell_dyn[cbase + c] = s_dyn[sl];
    }
    // the clock is shared by the replicas: every cluster that ran the whole batch holds the same final value
    if (complete && rank == 0 && tid == 0) { *pa.t[0] = l_t; *pa.t[1] = l_t; }
}

inline int resident_block_threads(int cars_per_replica) { return cars_per_replica <= 128 * TB2_RES_MAX_CTAS ? 128 : 256; }
// smallest cluster that holds `cars` cars, 0 if the world does not fit the resident form
inline int resident_cluster_size(int cars_per_replica) {
    const int blk = resident_block_threads(cars_per_replica);
    const int ctas = (cars_per_replica + blk - 1) / blk;
    if (ctas > TB2_RES_MAX_CTAS) return 0;
    int cl = 1;
    while (cl < ctas) cl *= 2;
    return cl;
}

template <typename R>
cudaError_t launch_resident_t(const StepArgsT<R>& a, const PersistArgsT<R>& pa, cudaStream_t stream) {
    const int cl = resident_cluster_size(a.cars_per_replica);
    if (cl == 0) return cudaErrorInvalidConfiguration;
    const int slots = (a.g.n_cells + cl - 1) / cl;
    const size_t smem = ResidentLayout<R>::bytes(slots);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    const bool ens = a.n_replicas > 1 || a.agent_car >= 0;
    const int blk = resident_block_threads(a.cars_per_replica);
    void (*kern)(const StepArgsT<R>, const PersistArgsT<R>) =
        blk == 128 ? (ens ? resident_kernel<R, true, 128> : resident_kernel<R, false, 128>)
                   : (ens ? resident_kernel<R, true, 256> : resident_kernel<R, false, 256>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cl > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl * a.n_replicas, 1, 1);
    cfg.blockDim = dim3(blk, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a, pa);
}

}  // namespace tb2k
