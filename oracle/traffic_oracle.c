/*
 * TEST INFRASTRUCTURE - CPU restatement (plain C, fp64) of the reference's per-timestep vehicle update.
 *
 * This file is the ORACLE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (traffic_b200/csrc) never calls into it.
 *
 * Parity status: PINNED - tests/test_oracle_golden.py checks this restatement against per-step vectors recorded
 * from the unmodified reference (oracle/ref_harness.py -> tests/golden/ (npz files)): discrete outputs bit-exact,
 * floats to <= a few ulp (np.arccos is an SVML vector routine in the numpy build that made the fixtures, libm
 * acos here).
 *
 * Every function cites the reference lines it follows (paths relative to donjpierce/traffic).  Arithmetic
 * follows numpy's evaluation order; compile with -ffp-contract=off so the compiler adds no FMAs of its own.
 * The only fused operation is the explicit fma() in dot2(): OpenBLAS ddot (which np.dot / np.linalg.norm
 * reach for 2-vectors) evaluates x0*y0 then fma(x1, y1, .) - measured on the fixture machine, 200000/200000.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* number of worker threads for the per-car loops (cars are independent within a step: every read of another car
 * uses start-of-step values).  1 = the reference's own single-threaded execution. */
static int g_threads = 1;
void to_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int to_get_threads(void) { return g_threads; }

typedef struct {
    int32_t n_cars, n_lights, n_faces;
    int32_t nbx, nby;            /* number of bin edges per axis: len(np.arange(a0, a1, 200))  models.py:16 */
    double ex0, ex1, exd;        /* e[0], e[1], e[1]-e[0] of the x edges (numpy arange fill rule) */
    double ey0, ey1, eyd;
    double speed_limit, stop_distance, free_distance, default_acceleration; /* simulation.py:13-16 */
} to_params;

/* numpy.isclose(a, b, rtol, atol): |a-b| <= atol + rtol*|b| and isfinite(b), or a == b   numeric.py:2496-2498 */
static inline int isclose_(double a, double b, double rtol, double atol) {
    return ((fabs(a - b) <= atol + rtol * fabs(b)) && isfinite(b)) || (a == b);
}

/* np.dot of two 2-vectors as OpenBLAS evaluates it */
static inline double dot2(double ax, double ay, double bx, double by) { return fma(ay, by, ax * bx); }
/* np.linalg.norm of a 2-vector = sqrt(dot(v, v))   models.py:69-71 */
static inline double norm2(double x, double y) { return sqrt(dot2(x, y, x, y)); }

static inline double edge_(int k, double e0, double e1, double d) {
    return k == 0 ? e0 : (k == 1 ? e1 : e0 + (double)k * d);
}

/* np.digitize(x, np.arange(a0, a1, 200)) : number of edges <= x; NaN -> nb   models.py:16-17 */
static int digitize_(double x, int nb, double e0, double e1, double d) {
    if (nb <= 0) return 0;
    if (isnan(x)) return nb;
    double q = floor((x - e0) / d);
    int k = q < -1.0 ? -1 : (q > (double)(nb - 1) ? nb - 1 : (int)q);
    while (k + 1 < nb && edge_(k + 1, e0, e1, d) <= x) k++;
    while (k >= 0 && edge_(k, e0, e1, d) > x) k--;
    return k + 1;
}

/* k-th element of np.linspace(start, stop, n)   numpy/_core/function_base.py:127-175 */
static inline double linspace_at(double start, double stop, int n, double step, int k) {
    if (n == 1) return 0.0 * (stop - start) + start;
    if (k == n - 1) return stop;
    return (double)k * step + start;
}

/* np.isclose(np.linspace(start, stop, n)[kmin:], v, rtol=1e-6).any()   navigation.py:443-444, 478-479
 * evaluated at the nearest sample (and its neighbours) instead of materialising the array. */
static int on_linspace(double start, double stop, int n, int kmin, double v) {
    if (n - kmin <= 0) return 0;
    if (n == 1) return isclose_(linspace_at(start, stop, 1, 0.0, 0), v, 1.0e-6, 1.0e-8);
    double step = (stop - start) / (double)(n - 1);
    double kf = rint((v - start) / step);
    int k0;
    if (!(kf >= (double)kmin)) k0 = kmin; else if (kf > (double)(n - 1)) k0 = n - 1; else k0 = (int)kf;
    for (int k = k0 - 1; k <= k0 + 1; ++k) {
        if (k < kmin || k > n - 1) continue;
        if (isclose_(linspace_at(start, stop, n, step, k), v, 1.0e-6, 1.0e-8)) return 1;
    }
    return 0;
}

/* int(abs(delta)) of models.upcoming_linspace (models.py:165-168); NaN/inf would raise in Python -> 0 here */
static inline int trunc_len(double delta) {
    double a = fabs(delta);
    if (!(a < 2.0e9)) return 0;
    return (int)a;
}
/* `linspace(...).any()`  (navigation.py:437, 473): non-empty and some element non-zero */
static inline int space_any(double start, double stop, int n) {
    if (n <= 0) return 0;
    if (n == 1) return !(linspace_at(start, stop, 1, 0.0, 0) == 0.0);
    return 1; /* |stop-start| >= 2 -> an endpoint is non-zero */
}

/* models.angle_between on unit vectors of u, w   models.py:74-87 */
static double angle_between_(double ux, double uy, double wx, double wy) {
    double nu = norm2(ux, uy), nw = norm2(wx, wy);
    double ax = ux / nu, ay = uy / nu, bx = wx / nw, by = wy / nw;
    double c = dot2(ax, ay, bx, by);
    if (c < -1.0) c = -1.0; else if (c > 1.0) c = 1.0;   /* np.clip keeps NaN */
    double a = acos(c);
    if (a > M_PI / 2) a = a - M_PI / 2;
    return a;
}

/* models.get_angles / upcoming_vectors on a 3-point view (models.py:174-208): vectors between consecutive
 * points, each divided by sqrt(dot(p_i, p_{i+1})) (sic). */
static double view_angle(const double* p0, const double* p1, const double* p2) {
    double s01 = sqrt(dot2(p0[0], p0[1], p1[0], p1[1]));
    double s12 = sqrt(dot2(p1[0], p1[1], p2[0], p2[1]));
    double ux = (p1[0] - p0[0]) / s01, uy = (p1[1] - p0[1]) / s01;
    double wx = (p2[0] - p1[0]) / s12, wy = (p2[1] - p1[1]) / s12;
    return angle_between_(ux, uy, wx, wy);
}

/* simulation.obstacle_factor   simulation.py:167-186 */
static double obstacle_factor(const to_params* P, double d) {
    if (P->stop_distance < d && d <= P->free_distance)
        return log(d / P->stop_distance) / log(P->free_distance / P->stop_distance);
    if (d <= P->stop_distance) return 0.0;
    return 1.0;
}

/* simulation.road_curvature_factor   simulation.py:135-164 */
static double curvature_factor(const to_params* P, int path_len, double angle, double d) {
    double theta = (path_len == 1) ? M_PI / 2 : angle;
    if (isclose_(theta, 0.0, 1.0e-1, 1.0e-8)) return 1.0;
    if (P->stop_distance < d && d <= P->free_distance) {
        double s = P->stop_distance * 2 * theta / M_PI;
        return log(d / s) / log(P->free_distance / s);
    }
    return 1.0;
}

/* Build the per-cell table of the two smallest car indices: the reference's candidate is the first *other* car in
 * DataFrame order inside my 200 m bin (navigation.py:438-442). */
static void build_cell_min2(int n, const int32_t* cell, int n_cells, int32_t* min12) {
    for (int c = 0; c < 2 * n_cells; ++c) min12[c] = INT32_MAX;
    for (int i = 0; i < n; ++i) {
        int c = cell[i];
        if (min12[2 * c] == INT32_MAX) min12[2 * c] = i;
        else if (min12[2 * c + 1] == INT32_MAX) min12[2 * c + 1] = i;
    }
}

/* One Cars.update(dt, lights)   cars.py:61-95 (+ find_obstacles 98-106, simulation.update_cars 20-80).
 * pos_new is scratch (N,2); min12 is scratch (2*n_cells). */
typedef struct {
    const to_params* P; double dt;
    double* pos; double* vel; double* route_time; int32_t* cursor;
    const int32_t* path_end; const int32_t* path_eff_end; const double* path_xy; const double* dest_xy;
    double* d_node_o; double* d_car_o; double* d_light_o; int32_t* cell; int32_t* leader;
    const double* light_xy; const int32_t* face_ptr; const double* face_vec; const uint8_t* face_go;
    const int32_t* cell_light_ptr; const int32_t* cell_light_idx; double* pos_new; int32_t* min12;
    int phase, lo, hi;
} car_job;

static void cars_range(const car_job* J)
{
    const to_params* P = J->P; const double dt = J->dt;
    double* pos = J->pos; double* vel = J->vel; double* route_time = J->route_time; int32_t* cursor = J->cursor;
    const int32_t* path_end = J->path_end; const int32_t* path_eff_end = J->path_eff_end;
    const double* path_xy = J->path_xy; const double* dest_xy = J->dest_xy;
    double* d_node_o = J->d_node_o; double* d_car_o = J->d_car_o; double* d_light_o = J->d_light_o;
    int32_t* cell = J->cell; int32_t* leader = J->leader;
    const double* light_xy = J->light_xy; const int32_t* face_ptr = J->face_ptr; const double* face_vec = J->face_vec;
    const uint8_t* face_go = J->face_go; const int32_t* cell_light_ptr = J->cell_light_ptr;
    const int32_t* cell_light_idx = J->cell_light_idx; double* pos_new = J->pos_new; const int32_t* min12 = J->min12;
    const int ncx = P->nbx + 1;
    if (J->phase == 0) {
        /* models.determine_bins   models.py:7-19, cars.py:78 */
        for (int i = J->lo; i < J->hi; ++i) {
            int xb = digitize_(pos[2 * i], P->nbx, P->ex0, P->ex1, P->exd);
            int yb = digitize_(pos[2 * i + 1], P->nby, P->ey0, P->ey1, P->eyd);
            cell[i] = yb * ncx + xb;
        }
        return;
    }
    for (int i = J->lo; i < J->hi; ++i) {
        const double x = pos[2 * i], y = pos[2 * i + 1];
        const int cur = cursor[i], end = path_end[i];
        const int has_path = cur < path_eff_end[i];        /* navigation.py:37-39 */
        const int path_len = has_path ? end - cur : 0;
        const double* p0 = path_xy + 2 * (size_t)cur;
        /* FrontView.crossed_node_event   navigation.py:91-111 */
        int crossed = 0;
        if (has_path) crossed = isclose_(p0[0], x, 1.0e-5, 1.0e-8) && isclose_(p0[1], y, 1.0e-5, 1.0e-8);
        /* FrontView.upcoming_node_position   navigation.py:73-89 */
        double tx, ty;
        if (has_path && !crossed) { tx = p0[0]; ty = p0[1]; }
        else if (has_path && path_len >= 2) { tx = p0[2]; ty = p0[3]; }
        else { tx = dest_xy[2 * i]; ty = dest_xy[2 * i + 1]; }
        /* FrontView.distance_to_node   navigation.py:62-71 */
        const double dx = tx - x, dy = ty - y;
        const double d_node = norm2(dx, dy);
        /* models.upcoming_linspace   models.py:155-171 */
        const int nx = trunc_len(dx), ny = trunc_len(dy);
        const int spaces = space_any(x, tx, nx) && space_any(y, ty, ny);
        /* navigation.car_obstacles   navigation.py:422-456 */
        double d_car = 0.0; int lead = -1;
        if (spaces) {
            int c = cell[i];
            int j = min12[2 * c] != i ? min12[2 * c] : min12[2 * c + 1];
            if (j != INT32_MAX) {
                double ox = pos[2 * j], oy = pos[2 * j + 1];
                if (on_linspace(x, tx, nx, 0, ox) && on_linspace(y, ty, ny, 0, oy)) {
                    d_car = norm2(ox - x, oy - y);
                    if (d_car != 0.0) lead = j;
                }
            }
        }
        /* navigation.light_obstacles   navigation.py:459-498 */
        double d_light = 0.0;
        if (spaces) {
            int c = cell[i];
            int b = cell_light_ptr[c], e = cell_light_ptr[c + 1];
            if (b < e) {
                d_light = -0.0;  /* loop exhausted -> implicit None */
                for (int q = b; q < e; ++q) {
                    int l = cell_light_idx[q];
                    double lx = light_xy[2 * l], ly = light_xy[2 * l + 1];
                    if (on_linspace(x, tx, nx, 1, lx) && on_linspace(y, ty, ny, 1, ly)) {
                        double cvx = lx - x, cvy = ly - y;
                        double ncv = norm2(cvx, cvy);
                        int hit = 0;
                        for (int f = face_ptr[l]; f < face_ptr[l + 1]; ++f) {
                            if (face_go[f]) continue;
                            /* models.determine_anti_parallel_vectors   models.py:90-97 */
                            double fx = face_vec[2 * f], fy = face_vec[2 * f + 1];
                            double nf = norm2(fx, fy);
                            double cc = dot2(cvx / ncv, cvy / ncv, fx / nf, fy / nf);
                            if (cc < -1.0) cc = -1.0; else if (cc > 1.0) cc = 1.0;
                            double ang = acos(cc);
                            if (isclose_(M_PI, ang, 1.0e-5, 0.1)) { hit = 1; break; }
                        }
                        if (hit) { d_light = ncv; break; }
                    } else { d_light = 0.0; break; }   /* light off my linspace -> return False */
                }
            }
        }
        d_node_o[i] = d_node; d_car_o[i] = d_car; d_light_o[i] = d_light; leader[i] = lead;

        /* simulation.update_cars   simulation.py:37-74 */
        double vx = 0.0, vy = 0.0;
        if (has_path) {
            route_time[i] = route_time[i] + dt;
            /* update_speed_factor   simulation.py:99-132 */
            double angle = 0.0; /* `False` */
            if (path_len >= 3) angle = view_angle(p0, p0 + 2, p0 + 4);
            double curv = curvature_factor(P, path_len, angle, d_node);
            const int c_t = d_car != 0.0, r_t = d_light != 0.0;   /* Python truthiness (NaN is truthy) */
            double f;
            if (c_t && r_t) f = (d_car <= d_light) ? obstacle_factor(P, d_car) : obstacle_factor(P, d_light);
            else if (c_t) {
                double cf = obstacle_factor(P, d_car);
                /* models.weigh_factors   models.py:41-66 */
                if (d_car > d_node) f = cf * cos(d_car / P->free_distance) + curv * sin(d_node / P->free_distance);
                else f = cf;
            } else if (r_t) f = obstacle_factor(P, d_light);
            else f = curv;
            f = fabs(f);
            double ux = dx / d_node, uy = dy / d_node;          /* models.unit_vector   models.py:74-76 */
            vx = ux * P->speed_limit * f; vy = uy * P->speed_limit * f;
            /* stall push   simulation.py:61-62, accelerate 83-96 */
            if (isclose_(0.0, vx, 1.0e-5, 0.1) && isclose_(0.0, vy, 1.0e-5, 0.1)) {
                int acc = !r_t && (!c_t || d_car > P->stop_distance);
                if (acc) { vx += P->default_acceleration; vy += P->default_acceleration; }
            }
            if (crossed) cursor[i] = cur + 1;
        } else {
            cursor[i] = end;   /* path becomes [] */
        }
        vel[2 * i] = vx; vel[2 * i + 1] = vy;
        /* Euler   cars.py:89-90 */
        pos_new[2 * i] = x + vx * dt; pos_new[2 * i + 1] = y + vy * dt;
    }
}

static void* cars_thread(void* arg) { cars_range((const car_job*)arg); return NULL; }

static void fan_out(car_job* proto, int phase, int n)
{
    int T = g_threads;
    if (T > n / 256 + 1) T = n / 256 + 1;
    car_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < T; ++t) {
        jobs[t] = *proto; jobs[t].phase = phase;
        jobs[t].lo = (int)((int64_t)n * t / T); jobs[t].hi = (int)((int64_t)n * (t + 1) / T);
    }
    for (int t = 1; t < T; ++t) pthread_create(&th[t], NULL, cars_thread, &jobs[t]);
    cars_range(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
}

void to_cars_update(const to_params* P, double dt,
                    double* pos, double* vel, double* route_time, int32_t* cursor,
                    const int32_t* path_end, const int32_t* path_eff_end,
                    const double* path_xy, const double* dest_xy,
                    double* d_node_o, double* d_car_o, double* d_light_o, int32_t* cell, int32_t* leader,
                    const double* light_xy, const int32_t* face_ptr, const double* face_vec, const uint8_t* face_go,
                    const int32_t* cell_light_ptr, const int32_t* cell_light_idx,
                    double* pos_new, int32_t* min12)
{
    const int n = P->n_cars;
    const int ncx = P->nbx + 1, ncy = P->nby + 1;
    car_job J = { P, dt, pos, vel, route_time, cursor, path_end, path_eff_end, path_xy, dest_xy, d_node_o, d_car_o,
                  d_light_o, cell, leader, light_xy, face_ptr, face_vec, face_go, cell_light_ptr, cell_light_idx,
                  pos_new, min12, 0, 0, 0 };
    fan_out(&J, 0, n);
    build_cell_min2(n, cell, ncx * ncy, min12);
    fan_out(&J, 1, n);
    memcpy(pos, pos_new, sizeof(double) * 2 * (size_t)n);
}

/* TrafficLights.update(dt)   cars.py:123-133 ; *t is time_elapsed */
void to_lights_update(int n_lights, double* t, double dt, const double* switch_time,
                      const int32_t* face_ptr, uint8_t* face_go)
{
    *t = *t + dt;
    for (int l = 0; l < n_lights; ++l) {
        double r = fmod(*t, switch_time[l]);             /* numpy remainder == fmod for positive operands */
        if (r != 0.0 && (switch_time[l] < 0) != (r < 0)) r += switch_time[l];
        if (isclose_(0.0, r, 1.0e-4, 1.0e-8))
            for (int f = face_ptr[l]; f < face_ptr[l + 1]; ++f) face_go[f] = !face_go[f];
    }
}

/* n_steps of the driver loop. order 0: cars then lights (test/profile_cars_update.py:28-34);
 * order 1: lights then cars (animate.py:63-64, environment.py:167-168). */
void to_run(const to_params* P, double dt, int n_steps, int order, double* lights_time,
            double* pos, double* vel, double* route_time, int32_t* cursor,
            const int32_t* path_end, const int32_t* path_eff_end,
            const double* path_xy, const double* dest_xy,
            double* d_node, double* d_car, double* d_light, int32_t* cell, int32_t* leader,
            const double* light_xy, const double* switch_time, const int32_t* face_ptr, const double* face_vec,
            uint8_t* face_go, const int32_t* cell_light_ptr, const int32_t* cell_light_idx)
{
    const int ncells = (P->nbx + 1) * (P->nby + 1);
    double* pos_new = (double*)malloc(sizeof(double) * 2 * (size_t)(P->n_cars > 0 ? P->n_cars : 1));
    int32_t* min12 = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)ncells);
    for (int s = 0; s < n_steps; ++s) {
        if (order == 1) to_lights_update(P->n_lights, lights_time, dt, switch_time, face_ptr, face_go);
        to_cars_update(P, dt, pos, vel, route_time, cursor, path_end, path_eff_end, path_xy, dest_xy,
                       d_node, d_car, d_light, cell, leader, light_xy, face_ptr, face_vec, face_go,
                       cell_light_ptr, cell_light_idx, pos_new, min12);
        if (order == 0) to_lights_update(P->n_lights, lights_time, dt, switch_time, face_ptr, face_go);
    }
    free(pos_new); free(min12);
}

int to_abi_version(void) { return 1; }
