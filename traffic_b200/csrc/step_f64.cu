// Fused fp64 ("parity mode") step kernel for sm_100a: one thread per car, one launch per simulation step.
//
// What one launch does (reference lines in parentheses):
//   * FrontView: view of <= 3 polyline points, crossed-node test, target point (navigation.py:31-42, 73-111)
//   * distance-to-node (navigation.py:62-71)
//   * leader lookup: first other car in my 200 m bin, accepted iff it lies on my pos->target linspace in x and
//     in y (navigation.py:422-456) - the candidate comes from a per-cell (min, 2nd-min car index) table that the
//     PREVIOUS launch built with two atomicMin per car from the positions it wrote
//   * red-light scan over the static cell -> lights CSR (navigation.py:459-498, models.py:90-97)
//   * speed factor, velocity law, stall push (simulation.py:37-74, 83-186; models.py:41-66, 174-208)
//   * Euler step (cars.py:89-90), route-cursor advance (simulation.py:47-52)
//   * bin of the new position (models.py:7-19) + insertion into the next step's cell table
//   * TrafficLights.update (cars.py:123-133) in the trailing blocks of the same grid (double-buffered faces)
//
// Arithmetic: this translation unit is compiled with --fmad=false so that every multiply/add rounds separately,
// as numpy's ufuncs do.  The only fused operation is the explicit fma() in dot2(), which is what OpenBLAS' ddot
// (reached by np.dot / np.linalg.norm on 2-vectors) does on the machine that recorded the golden vectors.
// Double-precision '/', sqrt, fmod, floor, rint are IEEE-exact on the device; log/sin/cos/acos are libdevice
// (<= 1-2 ulp from glibc), which is why float outputs are compared with a tolerance and discrete ones exactly.
#include "step_f64.cuh"

#include <math_constants.h>

namespace {

constexpr double kPi = 3.141592653589793;          // math.pi
constexpr double kHalfPi = 1.5707963267948966;     // math.pi / 2

// numpy.isclose(a, b, rtol, atol)  (numeric.py:2496-2498): asymmetric in b
__device__ __forceinline__ bool isclose_(double a, double b, double rtol, double atol) {
    return ((fabs(a - b) <= atol + rtol * fabs(b)) && isfinite(b)) || (a == b);
}
__device__ __forceinline__ double dot2(double ax, double ay, double bx, double by) { return fma(ay, by, ax * bx); }
__device__ __forceinline__ double norm2(double x, double y) { return sqrt(dot2(x, y, x, y)); }

__device__ __forceinline__ double edge_(int k, double e0, double e1, double d) {
    return k == 0 ? e0 : (k == 1 ? e1 : e0 + (double)k * d);
}
// np.digitize(x, np.arange(lo, hi, 200)): number of edges <= x, NaN -> nb  (models.py:16-17)
__device__ __forceinline__ int digitize_(double x, int nb, double e0, double e1, double d) {
    if (nb <= 0) return 0;
    if (isnan(x)) return nb;
    double q = floor((x - e0) / d);
    int k = q < -1.0 ? -1 : (q > (double)(nb - 1) ? nb - 1 : (int)q);
    while (k + 1 < nb && edge_(k + 1, e0, e1, d) <= x) k++;
    while (k >= 0 && edge_(k, e0, e1, d) > x) k--;
    return k + 1;
}

// element k of np.linspace(start, stop, n)  (numpy/_core/function_base.py:127-175)
__device__ __forceinline__ double linspace_at(double start, double stop, int n, double step, int k) {
    if (n == 1) return 0.0 * (stop - start) + start;
    if (k == n - 1) return stop;
    return (double)k * step + start;
}
// np.isclose(np.linspace(start, stop, n)[kmin:], v, rtol=1e-6).any() without materialising the array: the nearest
// sample (and its two neighbours, to absorb rounding of the index estimate) decides.  navigation.py:443-444,478-479
__device__ __forceinline__ bool on_linspace(double start, double stop, int n, int kmin, double v) {
    if (n - kmin <= 0) return false;
    if (n == 1) return isclose_(linspace_at(start, stop, 1, 0.0, 0), v, 1.0e-6, 1.0e-8);
    const double step = (stop - start) / (double)(n - 1);
    const double kf = rint((v - start) / step);
    int k0;
    if (!(kf >= (double)kmin)) k0 = kmin; else if (kf > (double)(n - 1)) k0 = n - 1; else k0 = (int)kf;
    bool hit = false;
#pragma unroll
    for (int dk = -1; dk <= 1; ++dk) {
        const int k = k0 + dk;
        if (k >= kmin && k <= n - 1) hit |= isclose_(linspace_at(start, stop, n, step, k), v, 1.0e-6, 1.0e-8);
    }
    return hit;
}
// int(abs(delta))  (models.py:165-168); Python raises on NaN/inf - we return 0
__device__ __forceinline__ int trunc_len(double delta) {
    const double a = fabs(delta);
    return (a < 2.0e9) ? (int)a : 0;
}
// `linspace(...).any()`  (navigation.py:437, 473)
__device__ __forceinline__ bool space_any(double start, double stop, int n) {
    if (n <= 0) return false;
    if (n == 1) return !(linspace_at(start, stop, 1, 0.0, 0) == 0.0);
    return true;
}

__device__ __forceinline__ double clip1(double c) { return c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c); }

// models.get_angles on a 3-point view  (models.py:174-208, 74-87)
__device__ __forceinline__ double view_angle(double2 p0, double2 p1, double2 p2) {
    const double s01 = sqrt(dot2(p0.x, p0.y, p1.x, p1.y));
    const double s12 = sqrt(dot2(p1.x, p1.y, p2.x, p2.y));
    const double ux = (p1.x - p0.x) / s01, uy = (p1.y - p0.y) / s01;
    const double wx = (p2.x - p1.x) / s12, wy = (p2.y - p1.y) / s12;
    const double nu = norm2(ux, uy), nw = norm2(wx, wy);
    const double c = clip1(dot2(ux / nu, uy / nu, wx / nw, wy / nw));
    double a = acos(c);
    if (a > kHalfPi) a = a - kHalfPi;
    return a;
}

// simulation.obstacle_factor  (simulation.py:167-186)
__device__ __forceinline__ double obstacle_factor(const StepArgs64& a, double d) {
    if (a.stop_distance < d && d <= a.free_distance) return log(d / a.stop_distance) / a.log_free_over_stop;
    if (d <= a.stop_distance) return 0.0;
    return 1.0;
}
// simulation.road_curvature_factor  (simulation.py:135-164)
__device__ __forceinline__ double curvature_factor(const StepArgs64& a, int path_len, double angle, double d) {
    const double theta = (path_len == 1) ? kHalfPi : angle;
    if (isclose_(theta, 0.0, 1.0e-1, 1.0e-8)) return 1.0;
    if (a.stop_distance < d && d <= a.free_distance) {
        const double s = a.stop_distance * 2 * theta / kPi;
        return log(d / s) / log(a.free_distance / s);
    }
    return 1.0;
}

__device__ __forceinline__ void lights_tick_body(int l, int n_lights, double dt, const double* __restrict__ switch_time,
                                                 const int32_t* __restrict__ face_ptr, const uint8_t* __restrict__ go_cur,
                                                 uint8_t* __restrict__ go_next, const double* t_cur, double* t_next) {
    // TrafficLights.update  (cars.py:130-133)
    const double t = *t_cur + dt;
    if (l == 0) *t_next = t;
    if (l >= n_lights) return;
    const double sw = switch_time[l];
    double r = fmod(t, sw);                       // numpy float remainder
    if (r != 0.0 && ((sw < 0.0) != (r < 0.0))) r += sw;
    const bool s = isclose_(0.0, r, 1.0e-4, 1.0e-8);
    for (int f = face_ptr[l]; f < face_ptr[l + 1]; ++f) go_next[f] = s ? (uint8_t)!go_cur[f] : go_cur[f];
}

__global__ void __launch_bounds__(256) step_f64_kernel(const __grid_constant__ StepArgs64 a) {
    if ((int)blockIdx.x >= a.car_blocks) {
        const int l = ((int)blockIdx.x - a.car_blocks) * 256 + (int)threadIdx.x;
        lights_tick_body(l, a.n_lights, a.dt, a.switch_time, a.face_ptr, a.go_cur, a.go_next, a.t_cur, a.t_next);
        return;
    }
    const int i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    // reset the table that the launch after next will fill
    for (int k = i; k < 2 * a.n_cells; k += a.car_blocks * 256) a.tab_clear[k] = TB2_EMPTY;
    if (i >= a.n_cars) return;

    const double2 p = a.pos_in[i];
    const double x = p.x, y = p.y;
    const int cur = a.cursor[i], end = a.path_end[i];
    const bool has_path = cur < a.path_eff_end[i];
    const int path_len = has_path ? end - cur : 0;
    double2 p0 = make_double2(0.0, 0.0), p1 = p0, p2 = p0;
    if (path_len >= 1) p0 = a.path_xy[cur];
    if (path_len >= 2) p1 = a.path_xy[cur + 1];
    if (path_len >= 3) p2 = a.path_xy[cur + 2];

    bool crossed = false;
    if (has_path) crossed = isclose_(p0.x, x, 1.0e-5, 1.0e-8) && isclose_(p0.y, y, 1.0e-5, 1.0e-8);
    double tx, ty;
    if (has_path && !crossed) { tx = p0.x; ty = p0.y; }
    else if (has_path && path_len >= 2) { tx = p1.x; ty = p1.y; }
    else { const double2 d = a.dest_xy[i]; tx = d.x; ty = d.y; }

    const double dx = tx - x, dy = ty - y;
    const double d_node = norm2(dx, dy);
    const int nx = trunc_len(dx), ny = trunc_len(dy);
    const bool spaces = space_any(x, tx, nx) && space_any(y, ty, ny);
    const int c = a.cell_in[i];

    // ---- leader lookup (navigation.car_obstacles) ----
    double d_car = 0.0;
    int lead = -1;
    if (spaces) {
        const int2 m = a.tab_cur[c];
        const int j = (m.x != i) ? m.x : m.y;
        if (j != TB2_EMPTY) {
            const double2 o = a.pos_in[j];
            if (on_linspace(x, tx, nx, 0, o.x) && on_linspace(y, ty, ny, 0, o.y)) {
                d_car = norm2(o.x - x, o.y - y);
                if (d_car != 0.0) lead = j;
            }
        }
    }
    // ---- red-light scan (navigation.light_obstacles) ----
    double d_light = 0.0;
    if (spaces) {
        const int b = a.cell_light_ptr[c], e = a.cell_light_ptr[c + 1];
        if (b < e) {
            d_light = -0.0;  // loop exhausted -> None
            for (int q = b; q < e; ++q) {
                const int l = a.cell_light_idx[q];
                const double2 lp = a.light_xy[l];
                if (on_linspace(x, tx, nx, 1, lp.x) && on_linspace(y, ty, ny, 1, lp.y)) {
                    const double cvx = lp.x - x, cvy = lp.y - y;
                    const double ncv = norm2(cvx, cvy);
                    bool hit = false;
                    for (int f = a.face_ptr[l]; f < a.face_ptr[l + 1]; ++f) {
                        if (a.go_cur[f]) continue;
                        const double2 fv = a.face_vec[f];
                        const double nf = norm2(fv.x, fv.y);
                        const double ang = acos(clip1(dot2(cvx / ncv, cvy / ncv, fv.x / nf, fv.y / nf)));
                        if (isclose_(kPi, ang, 1.0e-5, 0.1)) { hit = true; break; }
                    }
                    if (hit) { d_light = ncv; break; }
                } else { d_light = 0.0; break; }
            }
        }
    }
    a.d_node[i] = d_node;
    a.d_car[i] = d_car;
    a.d_light[i] = d_light;
    if (a.write_aux) a.leader[i] = lead;

    // ---- velocity law + route advance (simulation.update_cars) ----
    double vx = 0.0, vy = 0.0;
    if (has_path) {
        a.route_time[i] = a.route_time[i] + a.dt;
        double angle = 0.0;  // `False`
        if (path_len >= 3) angle = view_angle(p0, p1, p2);
        const double curv = curvature_factor(a, path_len, angle, d_node);
        const bool c_t = d_car != 0.0, r_t = d_light != 0.0;  // Python truthiness (NaN truthy)
        double f;
        if (c_t && r_t) f = (d_car <= d_light) ? obstacle_factor(a, d_car) : obstacle_factor(a, d_light);
        else if (c_t) {
            const double cf = obstacle_factor(a, d_car);
            if (d_car > d_node) f = cf * cos(d_car / a.free_distance) + curv * sin(d_node / a.free_distance);
            else f = cf;
        } else if (r_t) f = obstacle_factor(a, d_light);
        else f = curv;
        f = fabs(f);
        vx = dx / d_node * a.speed_limit * f;
        vy = dy / d_node * a.speed_limit * f;
        if (isclose_(0.0, vx, 1.0e-5, 0.1) && isclose_(0.0, vy, 1.0e-5, 0.1)) {
            const bool acc = !r_t && (!c_t || d_car > a.stop_distance);
            if (acc) { vx += a.default_acceleration; vy += a.default_acceleration; }
        }
        if (crossed) a.cursor[i] = cur + 1;
    } else if (cur != end) {
        a.cursor[i] = end;  // path becomes []
    }
    a.vel[i] = make_double2(vx, vy);
    const double xn = x + vx * a.dt, yn = y + vy * a.dt;  // cars.py:89-90
    a.pos_out[i] = make_double2(xn, yn);

    // ---- bin of the new position + insertion into the next step's table ----
    const int xb = digitize_(xn, a.nbx, a.ex0, a.ex1, a.exd);
    const int yb = digitize_(yn, a.nby, a.ey0, a.ey1, a.eyd);
    const int cn = yb * a.ncx + xb;
    a.cell_out[i] = cn;
    const int old = atomicMin(&a.tab_next[2 * cn], i);
    const int loser = old > i ? old : i;   // whoever is not the cell minimum competes for second place
    if (loser != TB2_EMPTY) atomicMin(&a.tab_next[2 * cn + 1], loser);
}

__global__ void __launch_bounds__(256) prepare_f64_kernel(int n_cars, int n_cells, const double2* __restrict__ pos,
                                                          int32_t* __restrict__ cell, int32_t* tab, int nbx, int nby,
                                                          int ncx, double ex0, double ex1, double exd, double ey0,
                                                          double ey1, double eyd) {
    const int i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (i >= n_cars) return;
    const double2 p = pos[i];
    const int cn = digitize_(p.y, nby, ey0, ey1, eyd) * ncx + digitize_(p.x, nbx, ex0, ex1, exd);
    cell[i] = cn;
    const int old = atomicMin(&tab[2 * cn], i);
    const int loser = old > i ? old : i;
    if (loser != TB2_EMPTY) atomicMin(&tab[2 * cn + 1], loser);
}

__global__ void fill_i32_kernel(int32_t* p, size_t n, int32_t v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

__global__ void __launch_bounds__(256) lights_tick_kernel(int n_lights, double dt, const double* switch_time,
                                                          const int32_t* face_ptr, const uint8_t* go_cur, uint8_t* go_next,
                                                          const double* t_cur, double* t_next) {
    const int l = (int)blockIdx.x * 256 + (int)threadIdx.x;
    lights_tick_body(l, n_lights, dt, switch_time, face_ptr, go_cur, go_next, t_cur, t_next);
}

}  // namespace

void launch_step_f64(const StepArgs64& a, cudaStream_t stream) {
    const int light_blocks = a.tick ? ((a.n_lights + 255) / 256 > 0 ? (a.n_lights + 255) / 256 : 1) : 0;
    const int grid = a.car_blocks + light_blocks;
    if (grid <= 0) return;
    step_f64_kernel<<<grid, 256, 0, stream>>>(a);
}

void launch_prepare_f64(int n_cars, int n_cells, const double2* pos, int32_t* cell, int32_t* tab3, int cur_table,
                        const StepArgs64& g, cudaStream_t stream) {
    const size_t n = (size_t)6 * n_cells;
    fill_i32_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, stream>>>(tab3, n, TB2_EMPTY);
    if (n_cars > 0)
        prepare_f64_kernel<<<(n_cars + 255) / 256, 256, 0, stream>>>(n_cars, n_cells, pos, cell,
                                                                     tab3 + (size_t)2 * n_cells * cur_table, g.nbx, g.nby,
                                                                     g.ncx, g.ex0, g.ex1, g.exd, g.ey0, g.ey1, g.eyd);
}

void launch_lights_tick(int n_lights, double dt, const double* switch_time, const int32_t* face_ptr,
                        const uint8_t* go_cur, uint8_t* go_next, const double* t_cur, double* t_next,
                        cudaStream_t stream) {
    const int blocks = (n_lights + 255) / 256 > 0 ? (n_lights + 255) / 256 : 1;
    lights_tick_kernel<<<blocks, 256, 0, stream>>>(n_lights, dt, switch_time, face_ptr, go_cur, go_next, t_cur, t_next);
}
