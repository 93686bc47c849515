// Argument block of the fused step kernel (see step_kernel.cuh), templated on the arithmetic type:
//   double = parity mode (reference arithmetic, absolute coordinates)
//   float  = throughput mode (coordinates relative to (origin_x, origin_y); the reference's relative tolerances are
//            still taken from the absolute coordinate)
// Device pointers only.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <type_traits>

#define TB2_EMPTY 0x7fffffff  // "no car" in the cell table

template <typename R> struct real2;
template <> struct real2<double> { using type = double2; };
template <> struct real2<float> { using type = float2; };

template <typename R>
struct GridArgsT {
    int32_t nbx, nby, ncx, n_cells;
    R ex0, ex1, exd, ey0, ey1, eyd;  // bin edges: e[k] = k==0 ? e0 : k==1 ? e1 : e0 + k*ed   (models.py:16)
    R inv_exd, inv_eyd;              // 1/ed, only used for the index estimate
    float f_inv_exd, f_inv_eyd;      // the same in fp32 (filters)
    float f_nbx1, f_nby1;            // (float)(nb - 1)
};

// Per-cell records.  Everything a car needs from its 200 m bin is addressable from the bin id alone and arrives in ONE
// dependent-load wave (two 32-byte sectors):
//
//  CellDynT  (per replica x cell, rewritten every step): the three rotating (smallest, second smallest) car-index tables
//            - read / being filled for the next step / being cleared - side by side, plus the double-buffered "go word"
//            of the cell's first light, so the table entry, the prefetched next-table entry (table_insert's skip test)
//            and the light-face states share a sector.
//  CellStatT (per cell, static, shared by replicas): position of the cell's first light and the unit vectors of its
//            first four faces quantised to snorm16 - enough for the filtered anti-parallel test (the exact fp64 test
//            runs only inside the filter's band and fetches face_unit[] then).
//
// go word: bits 0-7 = go state of faces 0..7 of the cell's first light, bits 8-15 = its face count (clipped to 255),
//          bits 16-31 = number of lights in the cell (clipped to 65535).  Written by the lights tick / rebuild kernel.
struct alignas(32) CellDynT {
    int32_t w[8];   // [2k], [2k+1]: table k (min, second-min);  [6], [7]: go word of face-state buffer 0 / 1
};
#define TB2_DYN_GO 6

template <typename R> struct CellStatT;
template <> struct alignas(32) CellStatT<double> {
    double2 xy0;       // first light of the cell: position
    int16_t fu[8];     // its first four faces: round(unit_vector * 32767)  (x0, y0, x1, y1, ...)
};
template <> struct alignas(32) CellStatT<float> {
    float2 xy0;
    int16_t fu[8];
    int32_t pad[2];
};

template <typename R>
struct StepArgsT {
    using R2 = typename real2<R>::type;
    GridArgsT<R> g;
    int32_t n_cars, n_lights, car_blocks, write_aux;   // n_cars / n_lights: totals over all replicas
    int32_t n_replicas, cars_per_replica, lights_per_replica, faces_per_replica;
    int32_t agent_car, step_index;
    int32_t skip_insert;
    int32_t freeze_done;         // rollouts: cars of a replica whose agent has arrived no longer move (Env.step returned)
    int32_t pipe_tma;            // 1: the TMA bulk-copy variant of the pipelined kernel
    int32_t pipe_ctas;           // > 0: launch the software-pipelined persistent kernel with this many car CTAs   // 1: the cell table is built by the sort pipeline (neighbor_sort.cu), not by atomics
    R speed_limit, stop_distance, free_distance, default_acceleration;
    R inv_free_distance;
    const void* log_table;   // LogEntry[TB2_LOG_N]: log over (stop, free] (see log_stop_free)
    R log_scale;             // TB2_LOG_N / (free - stop)
    R log_stop, inv_log_free_over_stop;  // host libm log(stop), 1/log(free/stop): simulation.obstacle_factor, simulation.py:180
    R dt;
    R origin_x, origin_y;  // float mode: absolute = stored + origin; 0 in double mode
    double dt64;           // lights always tick in fp64 (cars.py:130-133 is a round-off artefact, Appendix B-6)
    // cars
    const R2* pos_in;
    R2* pos_out;
    R2* vel;          // aux
    R* route_time;
    int32_t* cursor;
    const int32_t* path_end;
    const int32_t* path_eff_end;
    const R2* path_xy;
    const R2* pt_curv;   // per polyline point: curvature constants of the view that starts there (see prepare)
    const R2* dest_xy;
    R2* head_xy;      // per car: path_xy[cursor] (kept coalesced; refreshed when the cursor moves)
    R2* head_curv;    // per car: pt_curv[cursor]
    R* d_node;        // aux
    R* d_car;         // aux
    R* d_light;       // aux
    int32_t* cell;         // bin of the current position; updated in place by the owning thread
    int32_t* cell_prev;    // aux: bin of the start-of-step position (the reference's xbin/ybin columns)
    int32_t* leader;       // aux
    // per replica x cell dynamic record (see CellDynT): which of the three tables is read / filled / cleared this step,
    // and which go-word slot belongs to go_cur / go_next
    CellDynT* cell_dyn;
    int32_t tab_cur_k, tab_next_k, tab_clear_k;
    int32_t go_k_cur;
    // lights
    const R2* light_xy;
    const int32_t* face_ptr;
    const R2* face_vec;
    const R2* face_unit;  // face_vec / |face_vec| (prepare)
    const uint8_t* go_cur;
    const int32_t* cell_light_ptr;
    const int32_t* cell_light_idx;
    CellStatT<R>* cell_stat;       // [n_cells] derived from the CSR + light arrays at prepare time
    int2* light_cellinfo;          // [L] (cell id if the light is the first of its cell else -1, static bits of the go word)
    // fused TrafficLights.update (cars.py:123-133), runs in the trailing blocks when tick != 0
    int32_t tick, pad;
    const double* switch_time;
    uint8_t* go_next;
    const double* t_cur;
    double* t_next;
    // replica-ensemble bookkeeping (environment.py:120-126, 154-173)
    int32_t* done;
    // long sweeps (tb2_rollout, pipelined kernel): the replicas that were still running at the last read-back; the kernel
    // walks only their cars.  done_v: arrival flags indexed like live_list.  nullptr / 0: every replica is walked
    const int32_t* live_list;
    int32_t* done_v;
    int32_t n_live;
    int32_t n_walk;   // car positions the pipelined kernel walks: n_live * cars_per_replica with a list, else n_cars
    double* trip_time;
    int32_t* trip_step;
};

// Extra arguments of the persistent small-world kernel: every rotating buffer and where the rotation stands.
template <typename R>
struct PersistArgsT {
    typename real2<R>::type* pos[2];
    uint8_t* go[2];
    double* t[2];
    int32_t cur_pos, cur_tab, cur_go, cur_t;
    int32_t n_steps, order, aux_last;
    int32_t stop_when_done;   // resident ensembles: a replica whose agent has arrived leaves the launch
};

using StepArgs64 = StepArgsT<double>;
using StepArgs32 = StepArgsT<float>;

unsigned int debug_bounds_violations();
int debug_bounds_enabled();
int step_block_threads();
int pipe_block_threads();
int pipe_ctas_per_sm();
void launch_step(const StepArgs64& a, cudaStream_t stream);
void launch_step(const StepArgs32& a, cudaStream_t stream);
int persistent_max_cars(int precision_f32);   // largest world the persistent (one launch per batch) kernels can take
int persistent_grid_capacity_f32();
cudaError_t launch_persistent(const StepArgs64& a, const PersistArgsT<double>& pa, cudaStream_t stream);
cudaError_t launch_persistent(const StepArgs32& a, const PersistArgsT<float>& pa, cudaStream_t stream);
// one thread-block cluster per replica, world resident on chip (resident_kernel.cuh); cudaErrorInvalidConfiguration when
// the world does not fit that form
cudaError_t launch_resident(const StepArgs64& a, const PersistArgsT<double>& pa, cudaStream_t stream);
cudaError_t launch_resident(const StepArgs32& a, const PersistArgsT<float>& pa, cudaStream_t stream);
int resident_fits(int cars_per_replica, int lights_per_replica, int n_cells, int precision_f32);
// (re)derive everything the step kernel keeps cached from the uploaded state: per-point curvature constants, per-car
// head point + constants, unit face vectors, bins, and the cell table `cur_table` of the three in tab3 (all reset).
void launch_prepare(const StepArgs64& a, double2* pt_curv, double2* face_unit, int n_faces, cudaStream_t stream);
void launch_prepare(const StepArgs32& a, float2* pt_curv, float2* face_unit, int n_faces, cudaStream_t stream);
// range / monotonicity checks of the caller's index arrays (n_cars etc. are per replica except cursor: all replicas);
// *bad_dev (device int, zeroed by the caller) receives a bit mask of violated rules
void launch_validate(int n_cars_total, long long n_points, int n_lights, int n_faces, int n_cells, int n_cell_lights,
                     const int32_t* cursor, const int32_t* path_end, const int32_t* path_eff_end, const int32_t* face_ptr,
                     const int32_t* cell_light_ptr, const int32_t* cell_light_idx, int* bad_dev, cudaStream_t stream);
// re-derive the go words of the CURRENT face-state buffer (after tb2_upload(TB2_FACE_GO) / prepare)
void launch_rebuild_go_words(int n_lights_total, int lights_per_replica, int faces_per_replica, int n_cells,
                             const int32_t* face_ptr, const uint8_t* go_cur, const int2* light_cellinfo, CellDynT* cell_dyn,
                             int go_k, cudaStream_t stream);
static_assert(sizeof(CellStatT<double>) == 32 && sizeof(CellStatT<float>) == 32 && sizeof(CellDynT) == 32, "cell record layout");
// sort-based neighbour source (neighbor_sort.cu)
size_t neighbor_sort_temp_bytes(int n_cars, int key_bits);
int neighbor_sort_build(int n_cars, int cars_per_replica, int n_cells, int key_bits, const int32_t* cell, int32_t* keys_in,
                        int32_t* vals_in, int32_t* keys_out, int32_t* vals_out, void* temp, size_t temp_bytes,
                        CellDynT* cell_dyn, int table_k, cudaStream_t stream);
// TrafficLights.update alone (fp64 regardless of the car arithmetic)
void launch_lights_tick(int n_lights_total, int lights_per_replica, int faces_per_replica, int n_cells, double dt,
                        const double* switch_time, const int32_t* face_ptr, const uint8_t* go_cur, uint8_t* go_next,
                        const double* t_cur, double* t_next, const int2* light_cellinfo, CellDynT* cell_dyn, int go_k_next,
                        cudaStream_t stream);
