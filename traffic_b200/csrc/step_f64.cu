// Fused fp64 ("parity mode") step kernel for sm_100a: one thread per car, one launch per simulation step.
//
// What one launch does (reference lines in parentheses):
//   * FrontView: crossed-node test against the path head, target point (navigation.py:31-42, 73-111)
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
//
// What keeps it off the fp64 pipe (the first version spent 2,200 instructions per warp, see profiles/r1_a_*):
//   * everything that depends only on the route cursor is cached per car and refreshed when the cursor moves:
//     the path head point (coalesced 16 B instead of a 48 B gather) and the curvature constants of the 3-point view
//     (models.get_angles + the two logarithm arguments of simulation.road_curvature_factor), which are precomputed
//     per polyline point by the prepare kernel with the very same expressions;
//   * "is v on np.linspace(start, stop, n)" first runs a conservative bounding-interval test, then evaluates numpy's
//     samples only at the nearest index, located with an fp32 estimate (any estimate within +-1 works because the
//     three neighbouring samples are evaluated exactly);
//   * np.digitize uses a multiply for the index estimate and verifies it against the exact edges;
//   * per-step columns nobody can observe between steps (velocity, the three distances, leader, start-of-step bin)
//     are only written by launches with write_aux set (the last step of a tb2_step batch, every tb2_cars_update).
#include "step_f64.cuh"

#include <math_constants.h>

#ifndef TB2_MIN_BLOCKS
#define TB2_MIN_BLOCKS 4
#endif

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
// np.digitize(x, np.arange(lo, hi, 200)): number of edges <= x, NaN -> nb  (models.py:16-17).  The multiply only
// seeds the search; the two loops compare against the exact edges, so the result does not depend on inv_d's rounding.
__device__ __forceinline__ int digitize_(double x, int nb, double e0, double e1, double d, double inv_d) {
    if (nb <= 0) return 0;
    if (isnan(x)) return nb;
    const double q = floor((x - e0) * inv_d);
    int k = q < -1.0 ? -1 : (q > (double)(nb - 1) ? nb - 1 : (int)q);
    while (k + 1 < nb && edge_(k + 1, e0, e1, d) <= x) k++;
    while (k >= 0 && edge_(k, e0, e1, d) > x) k--;
    return k + 1;
}
// Filtered predicate: an fp32 estimate of the fractional bin index decides when x is at least 0.4 m (2e-3 bins) away
// from every edge - the estimate's error is < 1e-4 bins for maps up to ~100 km - and the exact fp64 routine runs only
// for positions inside that band (or outside the edge range, or non-finite).
__device__ __forceinline__ int digitize_f(double x, int nb, double e0, double e1, double d, double inv_d) {
    const float fx = __double2float_rn(x - e0) * __double2float_rn(inv_d);
    const float k = floorf(fx);
    const float fr = fx - k;
    if (fr > 2.0e-3f && fr < 1.0f - 2.0e-3f && k >= 0.0f && k < (float)(nb - 1) && nb < 30000) return (int)k + 1;
    return digitize_(x, nb, e0, e1, d, inv_d);
}
__device__ __forceinline__ int cell_of(const GridArgs& g, double x, double y) {
    return digitize_f(y, g.nby, g.ey0, g.ey1, g.eyd, g.inv_eyd) * g.ncx + digitize_f(x, g.nbx, g.ex0, g.ex1, g.exd, g.inv_exd);
}

// One axis of models.upcoming_linspace (models.py:155-171): np.linspace(start, stop, n), n = int(|stop - start|).
struct Axis {
    double start, stop, delta;
    int n;
};
__device__ __forceinline__ Axis make_axis(double start, double stop) {
    Axis a;
    a.start = start; a.stop = stop; a.delta = stop - start;
    const double m = fabs(a.delta);
    a.n = (m < 2.0e9) ? (int)m : 0;                     // int(abs(.)); Python raises on NaN/inf, we yield 0
    return a;
}
// `linspace(...).any()`  (navigation.py:437, 473): non-empty and some element non-zero
__device__ __forceinline__ bool axis_any(const Axis& a) {
    if (a.n <= 0) return false;
    if (a.n == 1) return !((0.0 * a.delta + a.start) == 0.0);
    return true;
}
// np.isclose(np.linspace(start, stop, n)[kmin:], v, rtol=1e-6).any()   (navigation.py:443-444, 478-479), exact:
// numpy's samples L_k = k*step + start (L_{n-1} = stop; n == 1: [0*delta + start]; function_base.py:127-175) are
// evaluated at the nearest index and its two neighbours - the nearest sample decides, whatever the tolerance.
__device__ __noinline__ bool on_axis_exact(double start, double stop, int n, int kmin, double v) {
    const double delta = stop - start;
    const double tol = 1.0e-8 + 1.0e-6 * fabs(v);
    const double step = (n > 1) ? delta / (double)(n - 1) : 0.0;
    double kf = (n > 1) ? rint((v - start) / step) : 0.0;
    kf = fmin(fmax(kf, (double)kmin), (double)(n - 1));   // NaN -> kmin
    const int k0 = (int)kf;
    const bool vfin = isfinite(v);
    bool hit = false;
#pragma unroll 1
    for (int k = k0 - 1; k <= k0 + 1; ++k) {
        if (k < kmin || k > n - 1) continue;
        const double Lk = (n == 1) ? 0.0 * delta + start : ((k == n - 1) ? stop : (double)k * step + start);
        hit |= ((fabs(Lk - v) <= tol) && vfin) || (Lk == v);
    }
    return hit;
}
// Filtered predicate in front of it: the distance r from v to the nearest sample is estimated in fp32 (error
// < 2e-3 m + 4e-7 * extent for n < 1e5); the exact routine only runs when r is within that error of the tolerance.
__device__ __forceinline__ bool on_axis(const Axis& a, int kmin, double v) {
    if (a.n - kmin <= 0) return false;
    if (a.n >= 2 && a.n < 100000) {
        const float w = __double2float_rn(v - a.start);
        const float fd = __double2float_rn(a.delta);
        const float fs = __fdividef(fd, (float)(a.n - 1));
        float k = rintf(__fdividef(w, fs));
        k = fminf(fmaxf(k, (float)kmin), (float)(a.n - 1));
        const float r = fabsf(w - k * fs);
        const float tf = __double2float_rn(1.0e-8 + 1.0e-6 * fabs(v));
        const float eps = 2.0e-3f + 4.0e-7f * (fabsf(w) + fabsf(fd));
        if (r + eps < tf) return true;
        if (r - eps > tf) return false;
    }
    return on_axis_exact(a.start, a.stop, a.n, kmin, v);
}

__device__ __forceinline__ double clip1(double c) { return c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c); }

// models.get_angles on a 3-point view  (models.py:174-208, 74-87)
__device__ __noinline__ double view_angle(double2 p0, double2 p1, double2 p2) {
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

// Curvature constants of the view whose head is polyline point q of a car whose polyline ends at `end`
// (simulation.road_curvature_factor, simulation.py:135-164): returns (s, L) with
//   s = stop*2*theta/pi and L = log(free/s), or s = 0 when the factor is identically 1 (theta isclose 0).
__device__ __forceinline__ double2 curv_constants(const double2* __restrict__ path_xy, int q, int end, double stop_d,
                                                  double free_d) {
    const int len = end - q;
    double theta;
    if (len == 1) theta = kHalfPi;
    else if (len >= 3) theta = view_angle(path_xy[q], path_xy[q + 1], path_xy[q + 2]);
    else theta = 0.0;  // models.get_angles returns False
    if (isclose_(theta, 0.0, 1.0e-1, 1.0e-8)) return make_double2(0.0, 0.0);
    const double s = stop_d * 2 * theta / kPi;
    return make_double2(s, log(free_d / s));
}

// models.determine_anti_parallel_vectors over the red faces of light l  (navigation.py:481-489, models.py:90-97).
// face_unit holds unit_vector(face) = face / norm(face), computed once by the prepare kernel with the reference's
// expression.  np.isclose(pi, angle, atol=0.1) is true iff angle >= T = (pi - 0.1) / (1 + 1e-5); acos is monotone, so
// the cosine decides: an fp32 cosine when it is further than 1e-4 from cos(T), the reference's fp64 expression
// otherwise, and acos itself only inside a +-1e-6 band.
constexpr double kCosT = -0.9950011283222804;  // cos(3.041562237967413)
__device__ __noinline__ bool antiparallel_exact(double cvx, double cvy, double2 fu) {
    const double ncv = norm2(cvx, cvy);
    const double c = clip1(dot2(cvx / ncv, cvy / ncv, fu.x, fu.y));
    if (c < kCosT - 1.0e-6) return true;
    if (c > kCosT + 1.0e-6) return false;
    return isclose_(kPi, acos(c), 1.0e-5, 0.1);
}
__device__ __forceinline__ bool red_face_ahead(const StepArgs64& a, int l, int fbase, double cvx, double cvy) {
    const float fx = __double2float_rn(cvx), fy = __double2float_rn(cvy);
    const float inv = rsqrtf(fx * fx + fy * fy);
    bool found = false;
    for (int f = a.face_ptr[l]; f < a.face_ptr[l + 1] && !found; ++f) {
        if (a.go_cur[fbase + f]) continue;
        const double2 fu = a.face_unit[f];
        const float c = (fx * __double2float_rn(fu.x) + fy * __double2float_rn(fu.y)) * inv;
        if (c < (float)kCosT - 1.0e-4f) found = true;
        else if (!(c > (float)kCosT + 1.0e-4f)) found = antiparallel_exact(cvx, cvy, fu);
    }
    return found;
}

// TrafficLights.update  (cars.py:130-133); lg runs over replicas x lights, geometry (face_ptr) is shared
__device__ __forceinline__ void lights_tick_body(int lg, int n_lights_total, int lights_per_replica, int faces_per_replica,
                                                 double dt, const double* __restrict__ switch_time,
                                                 const int32_t* __restrict__ face_ptr, const uint8_t* __restrict__ go_cur,
                                                 uint8_t* __restrict__ go_next, const double* t_cur, double* t_next) {
    const double t = *t_cur + dt;
    if (lg == 0) *t_next = t;
    if (lg >= n_lights_total) return;
    const int rep = lg / lights_per_replica, l = lg - rep * lights_per_replica;
    const int fbase = rep * faces_per_replica;
    const double sw = switch_time[lg];
    double r = fmod(t, sw);                       // numpy float remainder
    if (r != 0.0 && ((sw < 0.0) != (r < 0.0))) r += sw;
    const bool s = isclose_(0.0, r, 1.0e-4, 1.0e-8);
    for (int f = face_ptr[l]; f < face_ptr[l + 1]; ++f) go_next[fbase + f] = s ? (uint8_t)!go_cur[fbase + f] : go_cur[fbase + f];
}

__device__ __forceinline__ void table_insert(int32_t* tab, int cell, int i) {
    const int old = atomicMin(&tab[2 * cell], i);
    const int loser = old > i ? old : i;   // whoever is not the cell minimum competes for second place
    if (loser != TB2_EMPTY) atomicMin(&tab[2 * cell + 1], loser);
}

// AUX: also write the observation-only columns.  ENS: replica ensemble (per-replica cell tables / face states, agent
// arrival bookkeeping); a single simulation compiles these out.
template <bool AUX, bool ENS>
__global__ void __launch_bounds__(256, TB2_MIN_BLOCKS) step_f64_kernel(const __grid_constant__ StepArgs64 a) {
    if ((int)blockIdx.x >= a.car_blocks) {
        const int l = ((int)blockIdx.x - a.car_blocks) * 256 + (int)threadIdx.x;
        lights_tick_body(l, a.n_lights, a.lights_per_replica, a.faces_per_replica, a.dt, a.switch_time, a.face_ptr,
                         a.go_cur, a.go_next, a.t_cur, a.t_next);
        return;
    }
    const int i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    // reset the table that the launch after next will fill
    for (int k = i; k < 2 * a.g.n_cells * a.n_replicas; k += a.car_blocks * 256) a.tab_clear[k] = TB2_EMPTY;
    if (i >= a.n_cars) return;
    // replica of this car (1 for a single simulation): offsets into the per-replica cell tables and face states
    const int rep = ENS ? i / a.cars_per_replica : 0;
    const int cbase = rep * a.g.n_cells, fbase = rep * a.faces_per_replica, ibase = rep * a.cars_per_replica;

    // ---- independent loads first ----
    const double2 p = a.pos_in[i];
    const int cur = a.cursor[i];
    const int end = a.path_end[i];
    const int eff = a.path_eff_end[i];
    const int c = a.cell[i];
    const double2 head = a.head_xy[i];
    const double2 hc = a.head_curv[i];
    const double rt = a.route_time[i];
    // ---- dependent on the bin ----
    const int2 m = a.tab_cur[cbase + c];
    const int lb = a.cell_light_ptr[c], le = a.cell_light_ptr[c + 1];

    const double x = p.x, y = p.y;
    const bool has_path = cur < eff;
    const int path_len = has_path ? end - cur : 0;
    const bool crossed = has_path && isclose_(head.x, x, 1.0e-5, 1.0e-8) && isclose_(head.y, y, 1.0e-5, 1.0e-8);
    double2 t = head;
    if (!has_path || (crossed && path_len < 2)) t = a.dest_xy[i];
    else if (crossed) t = a.path_xy[cur + 1];

    const double dx = t.x - x, dy = t.y - y;
    const double d_node = norm2(dx, dy);
    const Axis ax = make_axis(x, t.x), ay = make_axis(y, t.y);
    const bool spaces = axis_any(ax) && axis_any(ay);

    // ---- leader lookup (navigation.car_obstacles) ----
    double d_car = 0.0;
    int lead = -1;
    const int j = (m.x != i) ? m.x : m.y;
    if (spaces && j != TB2_EMPTY) {
        const double2 o = a.pos_in[j];
        if (on_axis(ax, 0, o.x) && on_axis(ay, 0, o.y)) {
            d_car = norm2(o.x - x, o.y - y);
            if (d_car != 0.0) lead = j;
        }
    }
    // ---- red-light scan (navigation.light_obstacles) ----
    double d_light = 0.0;
    if (spaces && lb < le) {
        d_light = -0.0;  // loop exhausted -> None
        for (int q = lb; q < le; ++q) {
            const int l = a.cell_light_idx[q];
            const double2 lp = a.light_xy[l];
            if (on_axis(ax, 1, lp.x) && on_axis(ay, 1, lp.y)) {
                const double cvx = lp.x - x, cvy = lp.y - y;
                if (red_face_ahead(a, l, fbase, cvx, cvy)) { d_light = norm2(cvx, cvy); break; }
            } else { d_light = 0.0; break; }
        }
    }

    // ---- velocity law + route advance (simulation.update_cars) ----
    double vx = 0.0, vy = 0.0;
    if (has_path) {
        a.route_time[i] = rt + a.dt;
        // Both speed factors have the form log(d / den) / L on (stop, free]:
        //   simulation.road_curvature_factor (simulation.py:135-164): den = stop*2*theta/pi, L = log(free/den) (cached)
        //   simulation.obstacle_factor       (simulation.py:167-186): den = stop,            L = log(free/stop)
        // so they share one code path.  Pass 0 evaluates the factor the car actually uses (obstacle factor when it
        // has an obstacle, curvature factor otherwise); pass 1 only runs for cars that blend both (models.weigh_factors).
        const bool c_t = d_car != 0.0, r_t = d_light != 0.0;  // Python truthiness (NaN truthy)
        const bool obs = c_t || r_t;
        const bool weigh = c_t && !r_t && d_car > d_node;      // simulation.py:120-126
        const double d_obs = (c_t && r_t) ? ((d_car <= d_light) ? d_car : d_light) : (c_t ? d_car : d_light);
        double r0 = 1.0, r1 = 1.0;
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
            const bool as_obs = (it == 0) && obs;
            const bool wanted = (it == 0) ? (obs || hc.x != 0.0) : (weigh && hc.x != 0.0);
            const double d = as_obs ? d_obs : d_node;
            const double den = as_obs ? a.stop_distance : hc.x;
            const double L = as_obs ? a.log_free_over_stop : hc.y;
            double r = 1.0;
            if (wanted) {
                if (a.stop_distance < d && d <= a.free_distance) r = log(d / den) / L;
                else if (as_obs && d <= a.stop_distance) r = 0.0;
            }
            if (it == 0) r0 = r; else r1 = r;
        }
        double f = r0;   // curvature factor without an obstacle, obstacle factor with one
        if (weigh) f = r0 * cos(d_car / a.free_distance) + r1 * sin(d_node / a.free_distance);  // models.py:41-66
        f = fabs(f);
        vx = dx / d_node * a.speed_limit * f;
        vy = dy / d_node * a.speed_limit * f;
        if (isclose_(0.0, vx, 1.0e-5, 0.1) && isclose_(0.0, vy, 1.0e-5, 0.1)) {
            const bool acc = !r_t && (!c_t || d_car > a.stop_distance);
            if (acc) { vx += a.default_acceleration; vy += a.default_acceleration; }
        }
        if (crossed) {
            a.cursor[i] = cur + 1;
            if (cur + 1 < end) {          // new head: the point we just steered to
                a.head_xy[i] = t;
                a.head_curv[i] = a.pt_curv[cur + 1];
            }
        }
    } else if (cur != end) {
        a.cursor[i] = end;  // path becomes []
    }
    const double xn = x + vx * a.dt, yn = y + vy * a.dt;  // cars.py:89-90
    a.pos_out[i] = make_double2(xn, yn);
    if (AUX) {
        a.vel[i] = make_double2(vx, vy);
        a.d_node[i] = d_node;
        a.d_car[i] = d_car;
        a.d_light[i] = d_light;
        a.leader[i] = lead >= 0 ? lead - ibase : -1;
        a.cell_prev[i] = c;
    }
    // ---- bin of the new position + insertion into the next step's table ----
    const int cn = cell_of(a.g, xn, yn);
    if (cn != c) a.cell[i] = cn;
    table_insert(a.tab_next, cbase + cn, i);
    // rollout bookkeeping: Env.simulation_step tests FrontView.end_of_route on the agent BEFORE stepping
    // (environment.py:161-171, navigation.py:113-128; FrontView's default stop_distance is 5)
    if (ENS && a.agent_car >= 0 && i - ibase == a.agent_car && !a.done[rep]) {
        const double2 d = a.dest_xy[i];
        if (isclose_(0.0, d.x - x, 1.0e-5, 5.0) && isclose_(0.0, d.y - y, 1.0e-5, 5.0)) {
            a.done[rep] = 1;
            a.trip_time[rep] = rt;
            a.trip_step[rep] = a.step_index;
        }
    }
}

// prepare, pass 1: curvature constants of every polyline point still ahead of its car
__global__ void __launch_bounds__(128) prepare_points_kernel(int n_cars, const int32_t* __restrict__ cursor,
                                                             const int32_t* __restrict__ path_end,
                                                             const double2* __restrict__ path_xy, double2* pt_curv,
                                                             double stop_d, double free_d) {
    const int i = (int)blockIdx.x * 128 + (int)threadIdx.x;
    if (i >= n_cars) return;
    const int end = path_end[i];
    for (int q = cursor[i]; q < end; ++q) pt_curv[q] = curv_constants(path_xy, q, end, stop_d, free_d);
}

// prepare, pass 2: per-car head cache, bin, cell-table insertion
__global__ void __launch_bounds__(256) prepare_cars_kernel(int n_cars, GridArgs g, const double2* __restrict__ pos,
                                                           const int32_t* __restrict__ cursor,
                                                           const int32_t* __restrict__ path_end,
                                                           const double2* __restrict__ path_xy,
                                                           const double2* __restrict__ pt_curv, double2* head_xy,
                                                           double2* head_curv, int32_t* cell, int32_t* cell_prev,
                                                           int32_t* tab, int cars_per_replica) {
    const int i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (i >= n_cars) return;
    const int cur = cursor[i];
    double2 h = make_double2(0.0, 0.0), hc = h;
    if (cur < path_end[i]) { h = path_xy[cur]; hc = pt_curv[cur]; }
    head_xy[i] = h;
    head_curv[i] = hc;
    const double2 p = pos[i];
    const int cn = cell_of(g, p.x, p.y);
    cell[i] = cn;
    cell_prev[i] = cn;
    table_insert(tab, (i / cars_per_replica) * g.n_cells + cn, i);
}

// prepare: unit vectors of the light faces (models.unit_vector, models.py:74-76)
__global__ void __launch_bounds__(256) prepare_faces_kernel(int n_faces, const double2* __restrict__ face_vec,
                                                            double2* __restrict__ face_unit) {
    const int f = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (f >= n_faces) return;
    const double2 fv = face_vec[f];
    const double nf = norm2(fv.x, fv.y);
    face_unit[f] = make_double2(fv.x / nf, fv.y / nf);
}

__global__ void fill_i32_kernel(int32_t* p, size_t n, int32_t v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

__global__ void __launch_bounds__(256) lights_tick_kernel(int n_lights, int lights_per_replica, int faces_per_replica,
                                                          double dt, const double* switch_time,
                                                          const int32_t* face_ptr, const uint8_t* go_cur, uint8_t* go_next,
                                                          const double* t_cur, double* t_next) {
    const int l = (int)blockIdx.x * 256 + (int)threadIdx.x;
    lights_tick_body(l, n_lights, lights_per_replica, faces_per_replica, dt, switch_time, face_ptr, go_cur, go_next, t_cur,
                     t_next);
}

}  // namespace

void launch_step_f64(const StepArgs64& a, cudaStream_t stream) {
    const int light_blocks = a.tick ? ((a.n_lights + 255) / 256 > 0 ? (a.n_lights + 255) / 256 : 1) : 0;
    const int grid = a.car_blocks + light_blocks;
    if (grid <= 0) return;
    const bool ens = a.n_replicas > 1 || a.agent_car >= 0;
    if (ens) {
        if (a.write_aux) step_f64_kernel<true, true><<<grid, 256, 0, stream>>>(a);
        else step_f64_kernel<false, true><<<grid, 256, 0, stream>>>(a);
    } else {
        if (a.write_aux) step_f64_kernel<true, false><<<grid, 256, 0, stream>>>(a);
        else step_f64_kernel<false, false><<<grid, 256, 0, stream>>>(a);
    }
}

void launch_prepare_f64(const StepArgs64& a, double2* pt_curv, double2* face_unit, int n_faces, int32_t* tab3,
                        int cur_table, cudaStream_t stream) {
    const size_t n = (size_t)6 * a.g.n_cells * a.n_replicas;
    const size_t blocks = (n + 255) / 256;
    fill_i32_kernel<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, stream>>>(tab3, n, TB2_EMPTY);
    if (n_faces > 0) prepare_faces_kernel<<<(n_faces + 255) / 256, 256, 0, stream>>>(n_faces, a.face_vec, face_unit);
    if (a.n_cars > 0) {
        prepare_points_kernel<<<(a.n_cars + 127) / 128, 128, 0, stream>>>(a.n_cars, a.cursor, a.path_end, a.path_xy,
                                                                            pt_curv, a.stop_distance, a.free_distance);
        prepare_cars_kernel<<<(a.n_cars + 255) / 256, 256, 0, stream>>>(
            a.n_cars, a.g, a.pos_in, a.cursor, a.path_end, a.path_xy, pt_curv, a.head_xy, a.head_curv, a.cell,
            a.cell_prev, tab3 + (size_t)2 * a.g.n_cells * a.n_replicas * cur_table, a.cars_per_replica);
    }
}

void launch_lights_tick(const StepArgs64& a, cudaStream_t stream) {
    const int blocks = (a.n_lights + 255) / 256 > 0 ? (a.n_lights + 255) / 256 : 1;
    lights_tick_kernel<<<blocks, 256, 0, stream>>>(a.n_lights, a.lights_per_replica, a.faces_per_replica, a.dt,
                                                   a.switch_time, a.face_ptr, a.go_cur, a.go_next, a.t_cur, a.t_next);
}
