// Sort-based neighbour source (the pipeline BASELINE.json's north_star sketches): per step, a stable radix sort of the
// cars by bin id (CUB DeviceRadixSort, key = replica * cells + bin, value = car index; stable, so cars of a bin stay in
// index order), then one thread per sorted slot finds the segment boundaries with warp shuffles and writes the bin's
// first two car indices - the only ones navigation.car_obstacles (navigation.py:438-452) can ever look at - into the
// same (min, second-min) table the default atomicMin path fills.  Results are bit-identical to the atomic path
// (tests/test_gpu_parity.py::test_sort_neighbour_source_equals_atomic_table); it exists to measure the design choice:
// the sort moves ~40 B/car-step in 3+ extra launches to produce two integers per bin.
#include "step_args.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace {

__global__ void __launch_bounds__(256) sort_keys_kernel(int n_cars, int cars_per_replica, int n_cells,
                                                        const int32_t* __restrict__ cell, int32_t* __restrict__ keys,
                                                        int32_t* __restrict__ vals) {
    const int i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (i >= n_cars) return;
    keys[i] = (i / cars_per_replica) * n_cells + cell[i];
    vals[i] = i;
}

// one thread per sorted slot; the left neighbour's key comes from a warp shuffle (lane 0 reads it from memory)
__global__ void __launch_bounds__(256) segment_heads_kernel(int n_cars, const int32_t* __restrict__ keys,
                                                            const int32_t* __restrict__ vals, CellDynT* __restrict__ dyn,
                                                            int table_k) {
    const int p = (int)blockIdx.x * 256 + (int)threadIdx.x;
    const bool in = p < n_cars;
    const int key = in ? keys[p] : -1;
    int left = __shfl_up_sync(0xffffffffu, key, 1);
    if ((threadIdx.x & 31) == 0) left = (p > 0 && in) ? keys[p - 1] : -2;
    int right_key = __shfl_down_sync(0xffffffffu, key, 1);
    int right_val = __shfl_down_sync(0xffffffffu, in ? vals[p] : TB2_EMPTY, 1);
    if ((threadIdx.x & 31) == 31) {
        right_key = (p + 1 < n_cars) ? keys[p + 1] : -3;
        right_val = (p + 1 < n_cars) ? vals[p + 1] : TB2_EMPTY;
    }
    if (in && key != left) {   // head of a segment: first car of the bin, and the next slot if it is the same bin
        dyn[key].w[2 * table_k] = vals[p];
        dyn[key].w[2 * table_k + 1] = (right_key == key) ? right_val : TB2_EMPTY;
    }
}

}  // namespace

size_t neighbor_sort_temp_bytes(int n_cars, int key_bits) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, n_cars, 0, key_bits);
    return bytes;
}

// table_k of cell_dyn must already be cleared (all TB2_EMPTY).  Returns the number of kernel launches issued (CUB's included, estimated).
int neighbor_sort_build(int n_cars, int cars_per_replica, int n_cells, int key_bits, const int32_t* cell,
                        int32_t* keys_in, int32_t* vals_in, int32_t* keys_out, int32_t* vals_out, void* temp,
                        size_t temp_bytes, CellDynT* cell_dyn, int table_k, cudaStream_t stream) {
    if (n_cars <= 0) return 0;
    const int blocks = (n_cars + 255) / 256;
    sort_keys_kernel<<<blocks, 256, 0, stream>>>(n_cars, cars_per_replica, n_cells, cell, keys_in, vals_in);
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n_cars, 0, key_bits, stream);
    segment_heads_kernel<<<blocks, 256, 0, stream>>>(n_cars, keys_out, vals_out, cell_dyn, table_k);
    return 2 + 3;
}
