// Argument block of the fused fp64 step kernel (see step_f64.cu).  Device pointers only.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define TB2_EMPTY 0x7fffffff  // "no car" in the cell table

struct GridArgs {
    int32_t nbx, nby, ncx, n_cells;
    double ex0, ex1, exd, ey0, ey1, eyd;  // bin edges: e[k] = k==0 ? e0 : k==1 ? e1 : e0 + k*ed   (models.py:16)
    double inv_exd, inv_eyd;              // 1/ed, only used for the index guess (the result is verified exactly)
};

struct StepArgs64 {
    GridArgs g;
    int32_t n_cars, n_lights, car_blocks, write_aux;   // n_cars / n_lights: totals over all replicas
    int32_t n_replicas, cars_per_replica, lights_per_replica, faces_per_replica;
    int32_t agent_car, step_index;
    double speed_limit, stop_distance, free_distance, default_acceleration;
    double log_free_over_stop;  // host libm log(free/stop), simulation.py:180
    double dt;
    // cars
    const double2* pos_in;
    double2* pos_out;
    double2* vel;          // aux
    double* route_time;
    int32_t* cursor;
    const int32_t* path_end;
    const int32_t* path_eff_end;
    const double2* path_xy;
    const double2* pt_curv;   // per polyline point: curvature constants of the view that starts there (see prepare)
    const double2* dest_xy;
    double2* head_xy;      // per car: path_xy[cursor] (kept coalesced; refreshed when the cursor moves)
    double2* head_curv;    // per car: pt_curv[cursor]
    double* d_node;        // aux
    double* d_car;         // aux
    double* d_light;       // aux
    int32_t* cell;         // bin of the current position; updated in place by the owning thread
    int32_t* cell_prev;    // aux: bin of the start-of-step position (the reference's xbin/ybin columns)
    int32_t* leader;       // aux
    // cell -> (smallest, second smallest) car index; three rotating tables
    const int2* tab_cur;
    int32_t* tab_next;
    int32_t* tab_clear;
    // lights
    const double2* light_xy;
    const int32_t* face_ptr;
    const double2* face_vec;
    const double2* face_unit;  // face_vec / |face_vec| (prepare)
    const uint8_t* go_cur;
    const int32_t* cell_light_ptr;
    const int32_t* cell_light_idx;
    // fused TrafficLights.update (cars.py:123-133), runs in the trailing blocks when tick != 0
    int32_t tick, pad;
    const double* switch_time;
    uint8_t* go_next;
    const double* t_cur;
    double* t_next;
    // replica-ensemble bookkeeping (environment.py:120-126, 154-173)
    int32_t* done;
    double* trip_time;
    int32_t* trip_step;
};

void launch_step_f64(const StepArgs64& a, cudaStream_t stream);
// (re)derive everything the step kernel keeps cached from the uploaded state: per-point curvature constants, per-car
// head point + constants, bins, and the cell table `cur_table` of the three in tab3 (all three are reset).
void launch_prepare_f64(const StepArgs64& a, double2* pt_curv, double2* face_unit, int n_faces, int32_t* tab3,
                        int cur_table, cudaStream_t stream);
void launch_lights_tick(const StepArgs64& a, cudaStream_t stream);
