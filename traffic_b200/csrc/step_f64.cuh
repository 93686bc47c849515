// Argument block of the fused fp64 step kernel (see step_f64.cu).  Device pointers only.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define TB2_EMPTY 0x7fffffff  // "no car" in the cell table

struct StepArgs64 {
    int32_t n_cars, n_lights, n_cells, car_blocks;
    int32_t nbx, nby, ncx, write_aux;
    double ex0, ex1, exd, ey0, ey1, eyd;
    double speed_limit, stop_distance, free_distance, default_acceleration;
    double log_free_over_stop;  // host libm log(free/stop), simulation.py:180
    double dt;
    // cars
    const double2* pos_in;
    double2* pos_out;
    double2* vel;
    double* route_time;
    int32_t* cursor;
    const int32_t* path_end;
    const int32_t* path_eff_end;
    const double2* path_xy;
    const double2* dest_xy;
    double* d_node;
    double* d_car;
    double* d_light;
    const int32_t* cell_in;   // cell of pos_in (written by the previous step / prepare)
    int32_t* cell_out;        // cell of pos_out
    int32_t* leader;
    // cell -> (smallest, second smallest) car index; three rotating tables
    const int2* tab_cur;
    int32_t* tab_next;
    int32_t* tab_clear;
    // lights
    const double2* light_xy;
    const int32_t* face_ptr;
    const double2* face_vec;
    const uint8_t* go_cur;
    const int32_t* cell_light_ptr;
    const int32_t* cell_light_idx;
    // fused TrafficLights.update (cars.py:123-133), runs in the trailing blocks when tick != 0
    int32_t tick, pad;
    const double* switch_time;
    uint8_t* go_next;
    const double* t_cur;
    double* t_next;
};

void launch_step_f64(const StepArgs64& a, cudaStream_t stream);
void launch_prepare_f64(int n_cars, int n_cells, const double2* pos, int32_t* cell, int32_t* tab3 /*3 tables*/,
                        int cur_table, const StepArgs64& grid, cudaStream_t stream);
void launch_lights_tick(int n_lights, double dt, const double* switch_time, const int32_t* face_ptr,
                        const uint8_t* go_cur, uint8_t* go_next, const double* t_cur, double* t_next,
                        cudaStream_t stream);
