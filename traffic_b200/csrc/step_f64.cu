// fp64 ("parity mode") instantiation of the fused step kernel.  Compiled with --fmad=false: every multiply/add must
// round separately, as numpy's ufuncs do (see step_kernel.cuh).
#include "step_kernel.cuh"
#include "resident_kernel.cuh"

namespace {
__global__ void __launch_bounds__(256) lights_tick_kernel(int n_lights, int lights_per_replica, int faces_per_replica,
                                                          int n_cells, double dt, const double* switch_time,
                                                          const int32_t* face_ptr, const uint8_t* go_cur, uint8_t* go_next,
                                                          const double* t_cur, double* t_next, const int2* light_cellinfo,
                                                          CellDynT* cell_dyn, int go_k_next) {
    const int l = (int)blockIdx.x * 256 + (int)threadIdx.x;
    tb2k::lights_tick_body(l, n_lights, lights_per_replica, faces_per_replica, n_cells, dt, switch_time, face_ptr, go_cur,
                           go_next, t_cur, t_next, light_cellinfo, cell_dyn, go_k_next);
}
}  // namespace

int step_block_threads() { return TB2_BLOCK; }
int pipe_block_threads() { return TB2_PIPE_BLOCK; }
int pipe_ctas_per_sm() {
    // resident CTAs per SM of the pipelined kernel, asked of the occupancy calculator (registers and staging memory
    // both bound it); cached per process
    static int cached = -1;
    if (cached < 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tb2k::pipe_kernel<double, false, false>, TB2_PIPE_BLOCK, 0) !=
                cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = TB2_PIPE_MIN_BLOCKS;
        }
        cached = n;
    }
    return cached;
}

void launch_step(const StepArgs64& a, cudaStream_t stream) { tb2k::launch_step_t<double>(a, stream); }

int persistent_max_cars(int precision_f32) {
    // the cooperative variant's capacity depends on the device and the kernel variant; it is queried once per process for
    // the current device and only decides which path is TRIED - a refused launch falls back to per-step launches (api.cu)
    static int cap64 = -1, cap32 = -1;
    const int cluster_cap = TB2_PERSIST_BLOCK * TB2_PERSIST_MAX_CTAS;
    int& cap = precision_f32 ? cap32 : cap64;
    if (cap < 0) cap = precision_f32 ? persistent_grid_capacity_f32() : tb2k::persistent_grid_capacity_t<double>();
    return cap > cluster_cap ? cap : cluster_cap;
}
cudaError_t launch_persistent(const StepArgs64& a, const PersistArgsT<double>& pa, cudaStream_t stream) {
    return tb2k::launch_persistent_t<double>(a, pa, stream);
}

void launch_prepare(const StepArgs64& a, double2* pt_curv, double2* face_unit, int n_faces, cudaStream_t stream) {
    tb2k::launch_prepare_t<double>(a, pt_curv, face_unit, n_faces, stream);
}

void launch_rebuild_go_words(int n_lights_total, int lights_per_replica, int faces_per_replica, int n_cells,
                             const int32_t* face_ptr, const uint8_t* go_cur, const int2* light_cellinfo, CellDynT* cell_dyn,
                             int go_k, cudaStream_t stream) {
    if (n_lights_total <= 0) return;
    tb2k::rebuild_go_words_kernel<<<(n_lights_total + 255) / 256, 256, 0, stream>>>(
        n_lights_total, lights_per_replica, faces_per_replica, n_cells, face_ptr, go_cur, light_cellinfo, cell_dyn, go_k);
}

void launch_lights_tick(int n_lights_total, int lights_per_replica, int faces_per_replica, int n_cells, double dt,
                        const double* switch_time, const int32_t* face_ptr, const uint8_t* go_cur, uint8_t* go_next,
                        const double* t_cur, double* t_next, const int2* light_cellinfo, CellDynT* cell_dyn, int go_k_next,
                        cudaStream_t stream) {
    const int blocks = (n_lights_total + 255) / 256 > 0 ? (n_lights_total + 255) / 256 : 1;
    lights_tick_kernel<<<blocks, 256, 0, stream>>>(n_lights_total, lights_per_replica, faces_per_replica, n_cells, dt,
                                                   switch_time, face_ptr, go_cur, go_next, t_cur, t_next, light_cellinfo,
                                                   cell_dyn, go_k_next);
}

cudaError_t launch_resident(const StepArgs64& a, const PersistArgsT<double>& pa, cudaStream_t stream) {
    return tb2k::launch_resident_t<double>(a, pa, stream);
}
int resident_fits(int cars_per_replica, int lights_per_replica, int n_cells, int precision_f32) {
    const int cl = tb2k::resident_cluster_size(cars_per_replica);
    if (cl == 0 || lights_per_replica > cl * tb2k::resident_block_threads(cars_per_replica)) return 0;
    const int slots = (n_cells + cl - 1) / cl;
    const size_t smem = precision_f32 ? tb2k::ResidentLayout<float>::bytes(slots) : tb2k::ResidentLayout<double>::bytes(slots);
    return smem <= 200 * 1024 ? cl : 0;   // the cluster size, 0 = does not fit
}

void launch_validate(int n_cars_total, long long n_points, int n_lights, int n_faces, int n_cells, int n_cell_lights,
                     const int32_t* cursor, const int32_t* path_end, const int32_t* path_eff_end, const int32_t* face_ptr,
                     const int32_t* cell_light_ptr, const int32_t* cell_light_idx, int* bad_dev, cudaStream_t stream) {
    int n = n_cars_total;
    if (n_lights > n) n = n_lights;
    if (n_cells > n) n = n_cells;
    if (n_cell_lights > n) n = n_cell_lights;
    if (n < 1) n = 1;
    tb2k::validate_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n_cars_total, n_points, n_lights, n_faces, n_cells, n_cell_lights,
                                                               cursor, path_end, path_eff_end, face_ptr, cell_light_ptr,
                                                               cell_light_idx, bad_dev);
}

// -DTB2_DEBUG_BOUNDS builds: bit mask of the bounds rules violated since the last call (0 in regular builds)
unsigned int debug_bounds_violations() {
#ifdef TB2_DEBUG_BOUNDS
    unsigned int v = 0, z = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&v, g_bounds_violations, sizeof v);
    cudaMemcpyToSymbol(g_bounds_violations, &z, sizeof z);
    return v;
#else
    return 0;
#endif
}
int debug_bounds_enabled() {
#ifdef TB2_DEBUG_BOUNDS
    return 1;
#else
    return 0;
#endif
}
