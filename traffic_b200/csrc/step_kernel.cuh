// Fused step kernel for sm_100a, templated on the arithmetic type: one thread per car, ONE launch per simulation step.
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
//   * TrafficLights.update (cars.py:123-133) in the trailing blocks of the same grid (double-buffered faces, fp64)
//
// R = double ("parity mode", step_f64.cu, compiled with --fmad=false): every multiply/add rounds separately, as
// numpy's ufuncs do.  The only fused operation is the explicit fma() in dot2(), which is what OpenBLAS' ddot (reached
// by np.dot / np.linalg.norm on 2-vectors) does on the machine that recorded the golden vectors.  '/', sqrt, fmod, floor,
// rint are IEEE-exact on the device; log/sin/cos/acos are libdevice (<= 1-2 ulp from glibc), which is why float outputs
// are compared with a tolerance and discrete ones exactly.  The reference's predicates (np.isclose over np.linspace,
// np.digitize, the anti-parallel face test) are FILTERED: an fp32 evaluation with an explicit error bound decides unless
// the instance lies inside the bound's band, in which case the exact fp64 routine runs - same results, ~1 % of the
// lanes pay for fp64 division sequences.
//
// R = float ("throughput mode", step_f32.cu): the same algorithm in fp32 on coordinates relative to the map origin; the
// predicates are evaluated in fp32 only, the reference's relative tolerances still use the absolute coordinate.
//
// Everything that depends only on the route cursor is cached per car and refreshed when the cursor moves: the path
// head point (coalesced instead of a 3-point gather) and the curvature constants of the 3-point view, precomputed per
// polyline point by the prepare kernel with the same expressions.  Columns nobody can observe between the steps of a
// batch (velocity, the three distances, leader, start-of-step bin) are only written by AUX launches.
#pragma once
#include "step_args.cuh"

#include <cooperative_groups.h>
#include <math_constants.h>
#include <type_traits>
#include <cstdlib>

#ifndef TB2_MIN_BLOCKS
#define TB2_MIN_BLOCKS 8
#endif
#ifndef TB2_BLOCK
#define TB2_BLOCK 128   // threads per CTA of the step kernel (cars per CTA)
#endif
#ifndef TB2_PIPE_BLOCK
#define TB2_PIPE_BLOCK 128         // threads per CTA of the pipelined big-world kernel (cars per tile)
#endif
#ifndef TB2_PIPE_MIN_BLOCKS
#define TB2_PIPE_MIN_BLOCKS 5      // resident CTAs per SM (96 registers per thread, one 4-byte spill; ~42 KB of staging per CTA)
#endif
#ifndef TB2_PIPE_L2_AHEAD
#define TB2_PIPE_L2_AHEAD 3        // own-state columns are bulk-prefetched into L2 this many tiles ahead
#endif
#define TB2_PERSIST_BLOCK 256      // threads per CTA of the persistent small-world kernel
#define TB2_PERSIST_MAX_CTAS 16    // one thread-block cluster (16 needs the non-portable-size opt-in)
#define TB2K_ORDER_LIGHTS_CARS 1

// compute-sanitizer is closed on the GPU pool: the library carries its own bounds checks instead.  A build with
// -DTB2_DEBUG_BOUNDS records every out-of-range index the step kernels are about to dereference (bit mask in
// g_bounds_violations, read and reset by tb2_debug_bounds_violations()); scripts/gpu_bounds.sh runs the GPU suite on it.
#ifdef TB2_DEBUG_BOUNDS
__device__ unsigned int g_bounds_violations = 0;
#define TB2_BOUNDS(cond, bit) do { if (!(cond)) atomicOr(&g_bounds_violations, 1u << (bit)); } while (0)
#else
#define TB2_BOUNDS(cond, bit) do { } while (0)
#endif

namespace tb2k {

template <typename R> constexpr bool is_f64 = std::is_same<R, double>::value;

template <typename R> struct K;
template <> struct K<double> {
    static constexpr double pi = 3.141592653589793;        // math.pi
    static constexpr double half_pi = 1.5707963267948966;  // math.pi / 2
    static constexpr double cosT = -0.9950011283222804;    // cos((pi - 0.1) / (1 + 1e-5))
};
template <> struct K<float> {
    static constexpr float pi = 3.14159265358979f;
    static constexpr float half_pi = 1.57079632679490f;
    static constexpr float cosT = -0.9950011283222804f;
};

__device__ __forceinline__ double rsqrt_(double v) { return rsqrt(v); }
__device__ __forceinline__ float rsqrt_(float v) { return rsqrtf(v); }
__device__ __forceinline__ double rcp_(double v) { return __drcp_rn(v); }
__device__ __forceinline__ float rcp_(float v) { return 1.0f / v; }
__device__ __forceinline__ double log_(double v) { return log(v); }
__device__ __forceinline__ float log_(float v) { return logf(v); }
// sin(x + q*pi/2) for the arguments models.weigh_factors feeds it - distances over free_distance, a few radians, never
// above 1e5: two-constant Cody-Waite reduction (exact product with the 33-bit head of pi/2) and the fdlibm kernel
// polynomials, <= 1 ulp (2.2e-16 absolute) on [0, 100]; anything else takes libdevice.  ~20 instructions instead of ~45.
// (rare paths are kept out of line: the pipelined kernel is sensitive to the instruction footprint of its loop)
static __device__ __noinline__ double trig_libdevice(double x, int q) { return q ? cos(x) : sin(x); }
__device__ __forceinline__ double sin_quadrant(double x, int q) {
    if (!(fabs(x) < 1.0e5)) return trig_libdevice(x, q);
    const double kf = rint(x * 0.6366197723675814);
    double r = fma(-kf, 1.57079632673412561417e+00, x);
    r = fma(-kf, 6.07710050650619224932e-11, r);
    const int n = (int)kf + q;
    const double z = r * r;
    double v;
    if (n & 1) {
        double p = -1.13596475577881948265e-11;
        p = fma(p, z, 2.08757232129817482790e-09);
        p = fma(p, z, -2.75573143513906633035e-07);
        p = fma(p, z, 2.48015872894767294178e-05);
        p = fma(p, z, -1.38888888888741095749e-03);
        p = fma(p, z, 4.16666666666666019037e-02);
        v = fma(z * z, p, fma(-0.5, z, 1.0));
    } else {
        double p = 1.58969099521155010221e-10;
        p = fma(p, z, -2.50507602534068634195e-08);
        p = fma(p, z, 2.75573137070700676789e-06);
        p = fma(p, z, -1.98412698298579493134e-04);
        p = fma(p, z, 8.33333333332248946124e-03);
        p = fma(p, z, -1.66666666666666324348e-01);
        v = fma(p * z, r, r);
    }
    return (n & 2) ? -v : v;
}
#ifdef TB2_LIBDEVICE_TRIG
__device__ __forceinline__ double cos_(double v) { return cos(v); }
__device__ __forceinline__ double sin_(double v) { return sin(v); }
#else
__device__ __forceinline__ double cos_(double v) { return sin_quadrant(v, 1); }
__device__ __forceinline__ double sin_(double v) { return sin_quadrant(v, 0); }
#endif
__device__ __forceinline__ float cos_(float v) { return cosf(v); }
__device__ __forceinline__ float sin_(float v) { return sinf(v); }
__device__ __forceinline__ double acos_(double v) { return acos(v); }
__device__ __forceinline__ float acos_(float v) { return acosf(v); }
__device__ __forceinline__ double2 mk2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ float2 mk2(float x, float y) { return make_float2(x, y); }

// numpy.isclose(a, b, rtol, atol)  (numeric.py:2496-2498): asymmetric in b.  b_abs is the ABSOLUTE value of b (equal
// to b in double mode; b + origin in float mode where coordinates are stored relative to the map origin).
template <typename R>
__device__ __forceinline__ bool isclose_abs(R a, R b, R b_abs, R rtol, R atol) {
    return ((fabs(a - b) <= atol + rtol * fabs(b_abs)) & (bool)isfinite(b)) | (a == b);   // no short-circuit: branch-free
}
template <typename R>
__device__ __forceinline__ bool isclose_(R a, R b, R rtol, R atol) { return isclose_abs<R>(a, b, b, rtol, atol); }

// isclose(0, b): |0 - b| <= atol + rtol*|b| for finite b (b == 0 satisfies it: no separate equality term)
template <typename R>
__device__ __forceinline__ bool isclose0_(R b, R rtol, R atol) { return (fabs(b) <= atol + rtol * fabs(b)) & (bool)isfinite(b); }

template <typename R>
__device__ __forceinline__ R dot2(R ax, R ay, R bx, R by) { return fma(ay, by, ax * bx); }
template <typename R>
__device__ __forceinline__ R norm2(R x, R y) { return sqrt(dot2<R>(x, y, x, y)); }
// distance-to-node and its reciprocal.  The distance itself must be the correctly rounded sqrt (np.linalg.norm): a leader
// parked exactly on my target node makes `d_car > d_node` (simulation.py:120) an exact tie, which a 1-ulp shortcut
// (d = dd * rsqrt(dd)) would break.
__device__ __forceinline__ void norm2_inv(double x, double y, double& d, double& inv) {
    d = sqrt(dot2<double>(x, y, x, y));
    inv = __drcp_rn(d);
}
__device__ __forceinline__ void norm2_inv(float x, float y, float& d, float& inv) {
    d = sqrtf(dot2<float>(x, y, x, y));
    inv = 1.0f / d;
}
template <typename R>
__device__ __forceinline__ R clip1(R c) { return c < R(-1) ? R(-1) : (c > R(1) ? R(1) : c); }

// ---------------------------------------------------------------------------------------------------------------
// np.digitize(x, np.arange(lo, hi, 200)): number of edges <= x, NaN -> nb  (models.py:16-17)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double edge_(int k, double e0, double e1, double d) {
    return k == 0 ? e0 : (k == 1 ? e1 : e0 + (double)k * d);
}
// exact: the multiply only seeds the search; the two loops compare against the exact edges
static __device__ __noinline__ int digitize_exact(double x, int nb, double e0, double e1, double d, double inv_d) {
    if (nb <= 0) return 0;
    if (isnan(x)) return nb;
    const double q = floor((x - e0) * inv_d);
    int k = q < -1.0 ? -1 : (q > (double)(nb - 1) ? nb - 1 : (int)q);
    while (k + 1 < nb && edge_(k + 1, e0, e1, d) <= x) k++;
    while (k >= 0 && edge_(k, e0, e1, d) > x) k--;
    return k + 1;
}
// Filtered predicate: an fp32 estimate of the fractional bin index decides when x is at least 0.4 m (2e-3 bins) away
// from every edge (the estimate's error is < 1.3e-3 bins up to 8192 bins); inside that band one comparison with the
// exact fp64 value of the nearest edge decides; the general exact routine only runs for positions outside the edge
// range, non-finite ones, or huge maps.
__device__ __forceinline__ int digitize_(double x, int nb, double e0, double e1, double d, double inv_d, float f_inv_d,
                                         float f_nb1) {
    const float fx = __double2float_rn(x - e0) * f_inv_d;
    const float k = floorf(fx);
    const float fr = fx - k;
    if (k >= 0.0f && k < f_nb1 && nb <= 8192) {   // (up to 8192 bins the estimate is good to 1.3e-3 bins)
        if (fr > 2.0e-3f && fr < 1.0f - 2.0e-3f) return (int)k + 1;
        // inside the band: x lies strictly between the two neighbours of the nearest edge, so that edge alone decides
        const int ke = (int)k + (fr >= 0.5f ? 1 : 0);
        return ke + (x >= edge_(ke, e0, e1, d) ? 1 : 0);
    }
    return digitize_exact(x, nb, e0, e1, d, inv_d);
}
// float mode: closed form on the (uniform) local edges
__device__ __forceinline__ int digitize_(float x, int nb, float e0, float, float, float inv_d, float, float) {
    if (nb <= 0) return 0;
    if (isnan(x)) return nb;
    const float q = floorf((x - e0) * inv_d) + 1.0f;
    return q < 0.0f ? 0 : (q > (float)nb ? nb : (int)q);
}
template <typename R>
__device__ __forceinline__ int cell_of(const GridArgsT<R>& g, R x, R y) {
    return digitize_(y, g.nby, g.ey0, g.ey1, g.eyd, g.inv_eyd, g.f_inv_eyd, g.f_nby1) * g.ncx +
           digitize_(x, g.nbx, g.ex0, g.ex1, g.exd, g.inv_exd, g.f_inv_exd, g.f_nbx1);
}

// ---------------------------------------------------------------------------------------------------------------
// One axis of models.upcoming_linspace (models.py:155-171): np.linspace(start, stop, n), n = int(|stop - start|).
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
struct Axis {
    R start, stop, delta;
    int n;
};
template <typename R>
__device__ __forceinline__ Axis<R> make_axis(R start, R stop) {
    Axis<R> a;
    a.start = start; a.stop = stop; a.delta = stop - start;
    const R m = fabs(a.delta);
    a.n = (m < R(2.0e9)) ? (int)m : 0;                  // int(abs(.)); Python raises on NaN/inf, we yield 0
    return a;
}
// `linspace(...).any()`  (navigation.py:437, 473): non-empty and some element non-zero (absolute coordinates)
template <typename R>
__device__ __forceinline__ bool axis_any(const Axis<R>& a, R origin) {
    // n == 1: the single sample is 0*delta + start; delta is finite here (1 <= |delta| < 2), so in double mode
    // (origin 0) the sample is zero iff start is
    if constexpr (is_f64<R>) return (a.n >= 2) | ((a.n == 1) & (a.start != R(0)));
    else {
        if (a.n <= 0) return false;
        if (a.n == 1) return !((R(0) * a.delta + (a.start + origin)) == R(0));
        return true;
    }
}
// np.isclose(np.linspace(start, stop, n)[kmin:], v, rtol=1e-6).any()   (navigation.py:443-444, 478-479), exact:
// numpy's samples L_k = k*step + start (L_{n-1} = stop; n == 1: [0*delta + start]; function_base.py:127-175) are
// evaluated at the nearest index and its two neighbours - the nearest sample decides, whatever the tolerance.
static __device__ __noinline__ bool on_axis_exact(double start, double stop, int n, int kmin, double v) {
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
// fp32 distance from v to the nearest sample with index in [kmin, n-1]  (n >= 2)
__device__ __forceinline__ float nearest_sample_dist(float w /* v - start */, float fd /* delta */, int n, int kmin) {
    const float fs = __fdividef(fd, (float)(n - 1));
    float k = rintf(__fdividef(w, fs));
    k = fminf(fmaxf(k, (float)kmin), (float)(n - 1));
    return fabsf(w - k * fs);
}
// double mode: filtered predicate - the fp32 distance (error < 2e-3 m + 4e-7 * extent for n < 1e5) decides unless it is
// within that error of the tolerance, then the exact routine runs.
__device__ __forceinline__ bool on_axis(const Axis<double>& a, int kmin, double v, double /*origin*/) {
    if (a.n - kmin <= 0) return false;
    if (a.n >= 2 && a.n < 100000) {
        const float w = __double2float_rn(v - a.start);
        const float fd = __double2float_rn(a.delta);
        const float r = nearest_sample_dist(w, fd, a.n, kmin);
        const float tf = __double2float_rn(1.0e-8 + 1.0e-6 * fabs(v));
        const float eps = 2.0e-3f + 4.0e-7f * (fabsf(w) + fabsf(fd));
        if (r + eps < tf) return true;
        if (r - eps > tf) return false;
    }
    return on_axis_exact(a.start, a.stop, a.n, kmin, v);
}
// The same filter with everything that depends only on the axis hoisted out of the per-candidate test: sample spacing,
// its reciprocal, and the tolerance taken at the axis START instead of at the candidate (|tol(v) - tol(start)| =
// 1e-6 |v - start| is folded into the band).  One car tests up to four candidates coordinates per step (leader x/y, light
// x/y) against the same two axes.
struct AxisFilter {
    float fs, inv_fs, nm1, tf0, epsc;
    bool ok;
};
__device__ __forceinline__ float rcp_approx(float v) {   // one MUFU.RCP (<= 1 ulp); callers pass normal, non-zero values
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float v) {   // one MUFU.RSQ
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ AxisFilter axis_filter(const Axis<double>& a) {
    AxisFilter f;
    f.ok = a.n >= 2 && a.n < 100000;
    const float fd = __double2float_rn(a.delta);
    f.nm1 = (float)(a.n - 1);
    // spacing and its reciprocal, each within 2 ulp (f.ok: nm1 >= 1 and |fd| >= 2, so both reciprocals are of normal
    // numbers); the band below (2e-6 relative to the extent) is 16 ulp wide
    f.fs = fd * rcp_approx(f.nm1);
    f.inv_fs = f.nm1 * rcp_approx(fd);
    f.tf0 = __double2float_rn(1.0e-8 + 1.0e-6 * fabs(a.start));
    f.epsc = __fmaf_rn(2.0e-6f, fabsf(fd), 2.5e-3f);
    return f;
}
// Outcome of the fp32 filter for one coordinate of an axis with at least kmin + 1 samples: `yes` = certainly on the axis,
// `no` = certainly off it, neither = the exact routine decides.  Branch-free, so that the up to four tests of a car
// (leader x/y, light x/y) issue back to back.  An axis the filter does not apply to (f.ok false) has a NaN spacing:
// every comparison fails and the coordinate stays undecided; so does a NaN coordinate.
struct AxisHit {
    bool yes, no;
};
__device__ __forceinline__ AxisFilter axis_filter_nan(const Axis<double>& a) {
    AxisFilter f = axis_filter(a);
    f.fs = f.ok ? f.fs : __int_as_float(0x7fc00000);
    return f;
}
__device__ __forceinline__ AxisHit on_axis_quick(const Axis<double>& a, const AxisFilter& f, int kmin, double v) {
    const float w = __double2float_rn(v - a.start);
    float k = rintf(w * f.inv_fs);
    k = fminf(fmaxf(k, (float)kmin), f.nm1);              // NaN -> kmin
    const float r = fabsf(__fmaf_rn(-k, f.fs, w));         // distance to the nearest sample with index in [kmin, n-1]
    const float eps = __fmaf_rn(2.0e-6f, fabsf(w), f.epsc);
    AxisHit h;
    h.yes = r + eps < f.tf0;
    h.no = r - eps > f.tf0;
    return h;
}
template <typename StopFn>
__device__ __forceinline__ bool on_axis(const Axis<double>& a, const AxisFilter& f, int kmin, double v, StopFn stop) {
    if (a.n - kmin <= 0) return false;
    if (f.ok) {
        const float w = __double2float_rn(v - a.start);
        float k = rintf(w * f.inv_fs);
        k = fminf(fmaxf(k, (float)kmin), f.nm1);              // NaN -> kmin
        const float r = fabsf(__fmaf_rn(-k, f.fs, w));         // distance to the nearest sample with index in [kmin, n-1]
        const float eps = __fmaf_rn(2.0e-6f, fabsf(w), f.epsc);
        if (r + eps < f.tf0) return true;
        if (r - eps > f.tf0) return false;
    }
    return on_axis_exact(a.start, stop(), a.n, kmin, v);   // the linspace end point is only needed here: fetched lazily
}
// float mode: the fp32 distance decides (tolerance from the absolute coordinate)
__device__ __forceinline__ bool on_axis(const Axis<float>& a, int kmin, float v, float origin) {
    if (a.n - kmin <= 0) return false;
    const float w = v - a.start;
    const float r = (a.n >= 2) ? nearest_sample_dist(w, a.delta, a.n, kmin) : fabsf(w);
    return r <= 1.0e-8f + 1.0e-6f * fabsf(v + origin);
}

// models.get_angles on a 3-point view  (models.py:174-208, 74-87).  The reference divides each difference vector by
// sqrt(dot(p_i, p_{i+1})) of the ABSOLUTE position vectors (sic) before normalising it; double mode reproduces that,
// float mode (relative coordinates) skips the scale, which cancels in unit_vector.
template <typename R>
__device__ __noinline__ R view_angle(typename real2<R>::type p0, typename real2<R>::type p1, typename real2<R>::type p2) {
    R ux = p1.x - p0.x, uy = p1.y - p0.y, wx = p2.x - p1.x, wy = p2.y - p1.y;
    if constexpr (is_f64<R>) {
        const R s01 = sqrt(dot2<R>(p0.x, p0.y, p1.x, p1.y));
        const R s12 = sqrt(dot2<R>(p1.x, p1.y, p2.x, p2.y));
        ux = ux / s01; uy = uy / s01; wx = wx / s12; wy = wy / s12;
    }
    const R nu = norm2<R>(ux, uy), nw = norm2<R>(wx, wy);
    const R c = clip1<R>(dot2<R>(ux / nu, uy / nu, wx / nw, wy / nw));
    R a = acos_(c);
    if (a > K<R>::half_pi) a = a - K<R>::half_pi;
    return a;
}

// Curvature constants of the view whose head is polyline point q of a car whose polyline ends at `end`
// (simulation.road_curvature_factor, simulation.py:135-164): with s = stop*2*theta/pi the factor on (stop, free] is
// log(d/s) / log(free/s); cached as (log s, 1/log(free/s)) so that the step evaluates (log d - log s) * inv_L - one
// log, no division.  (0, 0) when the factor is identically 1 (theta isclose 0).
template <typename R>
__device__ __forceinline__ typename real2<R>::type curv_constants(const typename real2<R>::type* __restrict__ path_xy,
                                                                  int q, int end, R stop_d, R free_d) {
    const int len = end - q;
    R theta;
    if (len == 1) theta = K<R>::half_pi;
    else if (len >= 3) theta = view_angle<R>(path_xy[q], path_xy[q + 1], path_xy[q + 2]);
    else theta = R(0);  // models.get_angles returns False
    if (isclose_<R>(theta, R(0), R(1.0e-1), R(1.0e-8))) return mk2(R(0), R(0));
    const R s = stop_d * 2 * theta / K<R>::pi;
    return mk2(log_(s), R(1) / log_(free_d / s));
}

// models.determine_anti_parallel_vectors  (navigation.py:481-489, models.py:90-97).
// face_unit holds unit_vector(face) = face / norm(face), computed once by the prepare kernel with the reference's
// expression.  np.isclose(pi, angle, atol=0.1) is true iff angle >= T = (pi - 0.1) / (1 + 1e-5); acos is monotone, so
// the cosine decides: an fp32 cosine when it is further than the filter band from cos(T), the reference's expression
// otherwise, and acos itself only inside a +-1e-6 band.
template <typename R>
static __device__ __noinline__ bool antiparallel_exact(R cvx, R cvy, typename real2<R>::type fu) {
    const R ncv = norm2<R>(cvx, cvy);
    const R c = clip1<R>(dot2<R>(cvx / ncv, cvy / ncv, fu.x, fu.y));
    if (c < K<R>::cosT - R(1.0e-6)) return true;
    if (c > K<R>::cosT + R(1.0e-6)) return false;
    return isclose_<R>(K<R>::pi, acos_(c), R(1.0e-5), R(0.1));
}
// Band of the fp32 cosine filter.  Error budget: snorm16 face units (<= 2.2e-5), fp32 rounding of the light-car vector
// at UTM-scale differences, rsqrtf and the dot product (< 1e-5 together).
#define TB2_COS_BAND 1.5e-4f

// Red faces of light l (general path: per-face bytes + fp64/fp32 unit vectors from memory)
template <typename R>
__device__ __forceinline__ bool red_face_ahead(const StepArgsT<R>& a, const uint8_t* __restrict__ go_cur, int fb, int fe,
                                               int fbase, R cvx, R cvy) {
    const float fx = (float)cvx, fy = (float)cvy;
    const float inv = rsqrtf(fx * fx + fy * fy);
    bool found = false;
    for (int f = fb; f < fe && !found; ++f) {
        if (go_cur[fbase + f]) continue;
        const typename real2<R>::type fu = a.face_unit[f];
        const float c = (fx * (float)fu.x + fy * (float)fu.y) * inv;
        if (c < (float)K<double>::cosT - TB2_COS_BAND) found = true;
        else if (!(c > (float)K<double>::cosT + TB2_COS_BAND)) found = antiparallel_exact<R>(cvx, cvy, fu);
    }
    return found;
}

// navigation.light_obstacles (navigation.py:459-498) over the lights first, first+1, ... of bin c through the static
// cell -> lights CSR: the rare general path (a second light in the bin, or a first light with more than four faces).
// Returns the SQUARED distance to the first red anti-parallel face's light, +0 for False (a light off the linspace), -0
// for None (lights exhausted).
template <typename R>
static __device__ __noinline__ R light_scan_general(const StepArgsT<R>& a, const uint8_t* __restrict__ go_cur, int fbase,
                                                    int c, int first, R x, R tx, R y, R ty) {
    const Axis<R> ax = make_axis<R>(x, tx), ay = make_axis<R>(y, ty);
    const int lb = a.cell_light_ptr[c], le = a.cell_light_ptr[c + 1];
    for (int q = lb + first; q < le; ++q) {
        const int l = a.cell_light_idx[q];
        TB2_BOUNDS(l >= 0 && l < a.lights_per_replica, 7);
        const typename real2<R>::type lp = a.light_xy[l];
        if (!(on_axis(ax, 1, lp.x, a.origin_x) && on_axis(ay, 1, lp.y, a.origin_y))) return R(0);
        const R cvx = lp.x - x, cvy = lp.y - y;
        if (red_face_ahead<R>(a, go_cur, a.face_ptr[l], a.face_ptr[l + 1], fbase, cvx, cvy)) return dot2<R>(cvx, cvy, cvx, cvy);
    }
    return -R(0);
}

// exact re-evaluation of the first light's faces that fell into the filter band (bit k of `band` = face k)
template <typename R>
static __device__ __noinline__ bool first_light_band_exact(const StepArgsT<R>& a, int c, uint32_t band, R cvx, R cvy) {
    const int fb0 = a.face_ptr[a.cell_light_idx[a.cell_light_ptr[c]]];
    bool found = false;
    for (int k = 0; k < 4; ++k)
        if ((band >> k) & 1u) found = found || antiparallel_exact<R>(cvx, cvy, a.face_unit[fb0 + k]);
    return found;
}

// go word of a light's faces (see step_args.cuh): bit k = go state of face k, k < 8
__device__ __forceinline__ uint32_t go_mask_of(const uint8_t* __restrict__ go, int fb, int fe) {
    uint32_t m = 0;
    for (int f = fb, k = 0; f < fe && k < 8; ++f, ++k) m |= (go[f] ? 1u : 0u) << k;
    return m;
}

// TrafficLights.update  (cars.py:130-133); lg runs over replicas x lights, geometry (face_ptr) is shared.  Always fp64.
// Also refreshes the go word of the bin whose first light this is (next face-state buffer).
__device__ __forceinline__ void lights_tick_body(int lg, int n_lights_total, int lights_per_replica, int faces_per_replica,
                                                 int n_cells, double dt, const double* __restrict__ switch_time,
                                                 const int32_t* __restrict__ face_ptr, const uint8_t* __restrict__ go_cur,
                                                 uint8_t* __restrict__ go_next, const double* t_cur, double* t_next,
                                                 const int2* __restrict__ light_cellinfo, CellDynT* cell_dyn, int go_k_next,
                                                 const int32_t* __restrict__ frozen = nullptr) {
    const double t = __dadd_rn(*t_cur, dt);
    if (lg == 0) *t_next = t;
    if (lg >= n_lights_total) return;
    const int rep = lg / lights_per_replica, l = lg - rep * lights_per_replica;
    if (frozen && frozen[rep]) return;   // rollouts: nobody looks at the lights of a replica whose agent has arrived
    const int fbase = rep * faces_per_replica;
    const double sw = switch_time[lg];
    double r = fmod(t, sw);                       // numpy float remainder
    if (r != 0.0 && ((sw < 0.0) != (r < 0.0))) r = __dadd_rn(r, sw);
    // isclose(0, r, rtol=1e-4): |0 - r| <= 1e-8 + 1e-4*|r|
    const bool s = ((fabs(r) <= __dadd_rn(1.0e-8, __dmul_rn(1.0e-4, fabs(r)))) && isfinite(r)) || (r == 0.0);
    uint32_t mask = 0;
    const int fb = face_ptr[l], fe = face_ptr[l + 1];
    for (int f = fb; f < fe; ++f) {
        const uint8_t g = s ? (uint8_t)!go_cur[fbase + f] : go_cur[fbase + f];
        go_next[fbase + f] = g;
        if (f - fb < 8) mask |= (g ? 1u : 0u) << (f - fb);
    }
    const int2 info = light_cellinfo[l];
    if (info.x >= 0) cell_dyn[(size_t)rep * n_cells + info.x].w[TB2_DYN_GO + go_k_next] = (int32_t)((uint32_t)info.y | mask);
}

// Insert car i into table slot `t` (= &cell_dyn[bin].w[2k]) of the step being prepared.
// Entries only ever decrease while a table is being filled, so a car whose index already exceeds the bin's current
// second-smallest entry can never end up among the first two: it skips both atomics.  CTAs run roughly in index order,
// so after the first couple of cars of a bin almost everybody skips - no same-address serialisation in L2 and no wait
// for the atomic's return value.  `sec` is a (possibly stale, hence only larger = conservative) copy of t[1]; the step
// kernel prefetches it together with the bin's current-table entry, long before the new bin is known (a car stays in
// its bin in ~99 % of the steps).
__device__ __forceinline__ void table_insert(int32_t* t, int i, int sec) {
    if (i > sec) return;
    const int old = atomicMin(t, i);
    const int loser = old > i ? old : i;   // whoever is not the cell minimum competes for second place
    if (loser != TB2_EMPTY) atomicMin(t + 1, loser);
}

// What rotates from step to step: ping-pong positions, which of the three tables is read / filled, face-state buffer.
template <typename R>
struct Rot {
    const typename real2<R>::type* pos_in;
    typename real2<R>::type* pos_out;
    const uint8_t* go_cur;
    int32_t tab_cur_k, tab_next_k, go_k_cur;
    int32_t step_index;
};

// ---------------------------------------------------------------------------------------------------------------
// One car, one step, in three segments so that the same arithmetic serves the one-launch-per-step kernel, the
// persistent small-world kernels (registers filled by plain loads) and the software-pipelined big-world kernel
// (operands staged in shared memory by cp.async, see pipe_kernel):
//   seg_view   : FrontView - crossed-node test, target, distance-to-node, the two linspace axes
//   seg_detect : leader lookup + red-light scan
//   seg_move   : speed factor, velocity law, stall push, route advance, Euler step, bin + table insertion, bookkeeping
// AUX: also write the observation-only columns.  ENS: replica ensemble (per-replica cell records / face states, agent
// arrival bookkeeping); a single simulation compiles these out.
// ---------------------------------------------------------------------------------------------------------------
// Operand providers of the segments: plain registers (one-launch-per-step and persistent small-world kernels) ...
template <typename R>
struct OwnRegs {
    using R2 = typename real2<R>::type;
    const StepArgsT<R>* a;
    R2 t, hc_;
    R rt_;
    int cur, end, eff, c, sec_;
    __device__ __forceinline__ R2 target() const { return t; }
    __device__ __forceinline__ int3 route() const { return make_int3(cur, end, eff); }
    __device__ __forceinline__ R2 hc() const { return hc_; }
    __device__ __forceinline__ R rt() const { return rt_; }
    __device__ __forceinline__ int cell() const { return c; }
    __device__ __forceinline__ int sec() const { return sec_; }
    __device__ __forceinline__ int rep(int r) const { return r; }   // replica of the car (ensembles)
    __device__ __forceinline__ R2 next_curv(int q) const { return a->pt_curv[q]; }
    __device__ __forceinline__ R2 dest(int i) const { return a->dest_xy[i]; }
    __device__ __forceinline__ void after_cross(int, int) {}
    __device__ __forceinline__ CellStatT<R> stat() const { return a->cell_stat[c]; }
    // results go straight to the global columns
    R2* pos_out;
    __device__ __forceinline__ void store_route_time(int i, R v) { a->route_time[i] = v; }
    __device__ __forceinline__ void store_cursor(int i, int v) { a->cursor[i] = v; }
    __device__ __forceinline__ void store_head(int i, R2 v) { a->head_xy[i] = v; }
    __device__ __forceinline__ void store_head_curv(int i, R2 v) { a->head_curv[i] = v; }
    __device__ __forceinline__ void store_pos(int i, R2 v) { pos_out[i] = v; }
    __device__ __forceinline__ void store_cell(int i, int v) { a->cell[i] = v; }
    __device__ __forceinline__ void on_done() {}
};

template <typename R>
struct View {
    R x, y, dx, dy, d_node;   // (the target itself is re-fetched by the rare paths that need it: fewer live registers)
    R inv_d;                  // 1 / d_node
    int nx, ny;          // int(|dx|), int(|dy|): lengths of the two linspaces
    bool has_path, crossed, spaces;
    int path_len;
};

// FrontView.crossed_node_event (navigation.py:91-111): tolerance relative to the car's ABSOLUTE coordinate
template <typename R>
__device__ __forceinline__ bool view_crossed(const StepArgsT<R>& a, typename real2<R>::type p, typename real2<R>::type head,
                                             int cur, int eff) {
    if constexpr (is_f64<R>)   // absolute coordinates: origin is 0
        return (cur < eff) & isclose_abs<R>(head.x, p.x, p.x, R(1.0e-5), R(1.0e-8)) &
               isclose_abs<R>(head.y, p.y, p.y, R(1.0e-5), R(1.0e-8));
    else
        return (cur < eff) & isclose_abs<R>(head.x, p.x, p.x + a.origin_x, R(1.0e-5), R(1.0e-8)) &
               isclose_abs<R>(head.y, p.y, p.y + a.origin_y, R(1.0e-5), R(1.0e-8));
}
// where a car that crosses its path head steers next (FrontView.upcoming_node_position, navigation.py:73-89): the next
// route point, or the destination node when the head was the last one
template <typename R>
__device__ __forceinline__ const typename real2<R>::type* crossing_target(const StepArgsT<R>& a, int i, int cur, int end) {
    TB2_BOUNDS(cur >= 0 && cur < end, 6);
    return (end - cur < 2) ? &a.dest_xy[i] : &a.path_xy[cur + 1];
}
// t: the target - the cached head point (head_xy holds the destination once the route is exhausted, written by seg_move /
// prepare, so a parked car needs no extra load), or *crossing_target() on a crossing step
template <typename R>
__device__ __forceinline__ View<R> view_build(const StepArgsT<R>& a, typename real2<R>::type p, typename real2<R>::type t,
                                              int cur, int end, int eff, bool crossed) {
    View<R> v;
    v.x = p.x; v.y = p.y;
    v.has_path = cur < eff;
    v.path_len = v.has_path ? end - cur : 0;
    v.crossed = crossed;
    v.dx = t.x - v.x; v.dy = t.y - v.y;
    norm2_inv(v.dx, v.dy, v.d_node, v.inv_d);
    const Axis<R> ax = make_axis<R>(v.x, t.x), ay = make_axis<R>(v.y, t.y);
    v.nx = ax.n; v.ny = ay.n;
    v.spaces = axis_any<R>(ax, is_f64<R> ? R(0) : a.origin_x) && axis_any<R>(ay, is_f64<R> ? R(0) : a.origin_y);
    return v;
}
template <typename R>
__device__ __forceinline__ View<R> seg_view(const StepArgsT<R>& a, int i, typename real2<R>::type p,
                                            typename real2<R>::type head, int cur, int end, int eff,
                                            typename real2<R>::type& t) {
    const bool crossed = view_crossed<R>(a, p, head, cur, eff);
    t = head;
    if (crossed) t = *crossing_target<R>(a, i, cur, end);
    return view_build<R>(a, p, t, cur, end, eff, crossed);
}

template <typename R>
__device__ __forceinline__ Axis<R> axis_of(R start, R stop, R delta, int n) {
    Axis<R> ax;
    ax.start = start; ax.stop = stop; ax.delta = delta; ax.n = n;
    return ax;
}

// Filtered anti-parallel test of the bin's first light (static record: snorm16 face units; go bits from the go word):
// true iff a red face is anti-parallel to the car -> light vector (cvx, cvy).  Faces whose fp32 cosine falls inside the
// filter band are re-evaluated exactly.
template <typename R, typename Own>
__device__ __forceinline__ bool first_light_red_ahead(const StepArgsT<R>& a, int c, const Own& own, uint32_t gw,
                                                      int nf0, R cvx, R cvy) {
    const CellStatT<R> st = own.stat();   // (only the face half is used: the position half was fetched by the caller)
    const float fx = (float)cvx, fy = (float)cvy;
    const float inv = rsqrt_approx(fx * fx + fy * fy) * (1.0f / 32767.0f);
    uint32_t yes = 0, open = 0;   // bit k: face k certainly anti-parallel / not certainly not
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float cs = (fx * (float)st.fu[2 * k] + fy * (float)st.fu[2 * k + 1]) * inv;
        yes |= (cs < (float)K<double>::cosT - TB2_COS_BAND ? 1u : 0u) << k;
        open |= (cs > (float)K<double>::cosT + TB2_COS_BAND ? 0u : 1u) << k;   // NaN stays open
    }
    const uint32_t red = ~gw & ((1u << nf0) - 1u);
    if (yes & red) return true;
    const uint32_t band = open & ~yes & red;
    return band ? first_light_band_exact<R>(a, c, band, cvx, cvy) : false;
}

// navigation.car_obstacles (navigation.py:422-456) against candidate j at position o, then navigation.light_obstacles
// (navigation.py:459-498) for a bin whose go word is gw; load_stat() yields the bin's static light record.
// own: provider of the car's staged operands (OwnRegs / OwnSmem); target() - the car's target point - is only asked for
// by the exact / general fall-backs, stat() - the bin's static light record - only when the bin has a light.
template <typename R, typename Own>
__device__ __forceinline__ void seg_detect(const StepArgsT<R>& a, const uint8_t* __restrict__ go_cur, int fbase, int c,
                                           const View<R>& v, int j, typename real2<R>::type o, uint32_t gw,
                                           const Own& own, R& dd_car, int& lead, R& dd_light) {
    const Axis<R> ax = axis_of<R>(v.x, R(0), v.dx, v.nx), ay = axis_of<R>(v.y, R(0), v.dy, v.ny);
    const int nl = (int)(gw >> 16), nf0 = (int)((gw >> 8) & 0xffu);
    // outputs are SQUARED distances (0 = the reference's False, -0 = None): the one square root a step needs is taken in
    // seg_move, of the obstacle that counts (sqrt is monotone and correctly rounded, so min-then-sqrt == sqrt-then-min)
    dd_car = R(0);
    lead = -1;
    dd_light = R(0);
    if (!(v.spaces && (j != TB2_EMPTY || nl))) return;
    if constexpr (is_f64<R>) {
        // Double mode: the four linspace-membership tests of a car - leader x / y (kmin 0), first light x / y (kmin 1) -
        // run through the fp32 filter back to back, branch-free; the exact routine is called only for a pair that the
        // filter leaves open (on_axis is pure, so evaluating y although x fails changes nothing).
        const bool has_j = j != TB2_EMPTY;
        const bool lite = nl != 0 && nf0 <= 4;   // first light of the bin from the static record
        typename real2<R>::type lp = mk2(R(0), R(0));   // position of the bin's first light
        if (lite) lp = own.stat().xy0;
        // v.spaces: both axes have at least one sample, which is all the leader test (kmin 0) needs; the light test
        // skips the car's own sample (kmin 1): with a single-sample axis it fails (navigation.py:478-479 on an empty slice)
        const bool lite2 = lite & (min(ax.n, ay.n) >= 2);
        const AxisFilter ffx = axis_filter_nan(ax), ffy = axis_filter_nan(ay);
        const AxisHit cx = on_axis_quick(ax, ffx, 0, o.x), cy = on_axis_quick(ay, ffy, 0, o.y);
        const AxisHit lx = on_axis_quick(ax, ffx, 1, lp.x), ly = on_axis_quick(ay, ffy, 1, lp.y);
        bool car_hit = has_j & cx.yes & cy.yes, lit_hit = lite2 & lx.yes & ly.yes;
        const bool car_open = has_j & !(cx.no | cy.no) & !car_hit, lit_open = lite2 & !(lx.no | ly.no) & !lit_hit;
        if (car_open | lit_open) {   // rare: some coordinate is inside the filter band
            const typename real2<R>::type t = own.target();
            if (car_open)
                car_hit = (cx.yes || on_axis_exact(ax.start, t.x, ax.n, 0, o.x)) &&
                          (cy.yes || on_axis_exact(ay.start, t.y, ay.n, 0, o.y));
            if (lit_open)
                lit_hit = (lx.yes || on_axis_exact(ax.start, t.x, ax.n, 1, lp.x)) &&
                          (ly.yes || on_axis_exact(ay.start, t.y, ay.n, 1, lp.y));
        }
        if (car_hit) {
            dd_car = dot2<R>(o.x - v.x, o.y - v.y, o.x - v.x, o.y - v.y);
            if (dd_car != R(0)) lead = j;
        }
        if (nl) {
            if (!lite) {
                const typename real2<R>::type t = own.target();
                dd_light = light_scan_general<R>(a, go_cur, fbase, c, 0, v.x, t.x, v.y, t.y);
            } else if (lit_hit) {
                const R cvx = lp.x - v.x, cvy = lp.y - v.y;
                if (first_light_red_ahead<R>(a, c, own, gw, nf0, cvx, cvy)) dd_light = dot2<R>(cvx, cvy, cvx, cvy);
                else if (nl > 1) {
                    const typename real2<R>::type t = own.target();
                    dd_light = light_scan_general<R>(a, go_cur, fbase, c, 1, v.x, t.x, v.y, t.y);
                }
                else dd_light = -R(0);   // lights exhausted -> None
            }
        }
    } else {
        // float mode: plain fp32 membership tests
        if (j != TB2_EMPTY) {
            if (on_axis(ax, 0, o.x, a.origin_x) && on_axis(ay, 0, o.y, a.origin_y)) {
                dd_car = dot2<R>(o.x - v.x, o.y - v.y, o.x - v.x, o.y - v.y);
                if (dd_car != R(0)) lead = j;
            }
        }
        if (nl) {
            if (nf0 > 4) {
                const typename real2<R>::type t = own.target();
                dd_light = light_scan_general<R>(a, go_cur, fbase, c, 0, v.x, t.x, v.y, t.y);
            } else {
                const CellStatT<R> st = own.stat();
                if (on_axis(ax, 1, st.xy0.x, a.origin_x) && on_axis(ay, 1, st.xy0.y, a.origin_y)) {
                    const R cvx = st.xy0.x - v.x, cvy = st.xy0.y - v.y;
                    if (first_light_red_ahead<R>(a, c, own, gw, nf0, cvx, cvy)) dd_light = dot2<R>(cvx, cvy, cvx, cvy);
                    else if (nl > 1) {
                        const typename real2<R>::type t = own.target();
                        dd_light = light_scan_general<R>(a, go_cur, fbase, c, 1, v.x, t.x, v.y, t.y);
                    }
                    else dd_light = -R(0);   // lights exhausted -> None
                }
            }
        }
    }
}

// log(d) for d in (stop, free] - the only interval on which either speed factor needs a logarithm - from a 128-entry
// table of (c_k, 1/c_k, log c_k) over equal sub-intervals: log d = log c_k + log1p(u), u = (d - c_k)/c_k, |u| < 8e-3,
// degree-7 series.  <= 2 ulp from a correctly rounded log, ~15 instructions instead of libdevice's ~100.
#define TB2_LOG_N 128
struct alignas(32) LogEntry { double c, inv_c, log_c, pad; };
__device__ __forceinline__ double log_stop_free(const LogEntry* __restrict__ tab, double d, double stop, double scale) {
    int k = (int)((d - stop) * scale);
    k = k < 0 ? 0 : (k > TB2_LOG_N - 1 ? TB2_LOG_N - 1 : k);
    const double2 ci = *reinterpret_cast<const double2*>(&tab[k].c);
    const double lc = tab[k].log_c;
    const double u = (d - ci.x) * ci.y;
    double p = 1.0 / 7.0;
    p = fma(p, u, -1.0 / 6.0);
    p = fma(p, u, 1.0 / 5.0);
    p = fma(p, u, -1.0 / 4.0);
    p = fma(p, u, 1.0 / 3.0);
    p = fma(p, u, -0.5);
    p = fma(p, u, 1.0);
    return fma(u, p, lc);
}
__device__ __forceinline__ float log_stop_free(const LogEntry*, float d, float, float) { return logf(d); }

// next_curv(): curvature constants of the route point after the head (pt_curv[cur + 1]); insert(t, i, sec): table
// insertion policy (immediate, or deferred by the pipelined kernel).
// own: provider of the car's staged operands, each fetched where it is used (route() = (cursor, path_end, path_eff_end),
// target(), next_curv() only by the crossing / parking branches); insert(t, i, sec): table insertion policy (immediate,
// or deferred by the pipelined kernel).
template <typename R, bool AUX, bool ENS, typename Own, typename Insert>
__device__ __forceinline__ void seg_move(const StepArgsT<R>& a, const Rot<R>& rot, const LogEntry* __restrict__ logtab, int i,
                                         int rep, const View<R>& v, R dd_car, int lead, R dd_light, Own& own,
                                         Insert insert) {
    using R2 = typename real2<R>::type;
    // (ensembles: the replica index is asked of the provider where it is needed, not carried across the fp64 sections)
    auto rep_now = [&]() { return ENS ? own.rep(rep) : 0; };
    const R x = v.x, y = v.y, d_node = v.d_node;
    // rollout bookkeeping: Env.simulation_step tests FrontView.end_of_route on the agent BEFORE stepping
    // (environment.py:161-171, navigation.py:113-128; FrontView's default stop_distance is 5); it needs nothing but the
    // start-of-step state, so it runs first - nothing of it stays live across the arithmetic below
    if (ENS && a.agent_car >= 0 && i - rep_now() * a.cars_per_replica == a.agent_car && !a.done[rep_now()]) {
        const R2 d = own.dest(i);
        if (isclose_<R>(R(0), d.x - x, R(1.0e-5), R(5.0)) && isclose_<R>(R(0), d.y - y, R(1.0e-5), R(5.0))) {
            const int r = rep_now();
            a.done[r] = 1;
            a.trip_time[r] = (double)own.rt();   // start-of-step route time
            a.trip_step[r] = rot.step_index;
            own.on_done();
        }
    }
    // ---- velocity law + route advance (simulation.update_cars) ----
    R vx = R(0), vy = R(0);
    if (v.has_path) {
        own.store_route_time(i, own.rt() + a.dt);
        const R2 hc = own.hc();
        // Both speed factors have the form log(d / den) / L on (stop, free]:
        //   simulation.road_curvature_factor (simulation.py:135-164): den = stop*2*theta/pi, L = log(free/den)
        //   simulation.obstacle_factor       (simulation.py:167-186): den = stop,            L = log(free/stop)
        // so they share one code path, evaluated as (log d - log den) * (1/L) with log den and 1/L cached per route
        // point (curvature) or per simulation (obstacle): one log and no division per car, within 4 ulp of the
        // reference's expression.  Pass 0 evaluates the factor the car actually uses (obstacle factor when it has an
        // obstacle, curvature factor otherwise); pass 1 only runs for cars that blend both (models.weigh_factors).
        const bool c_t = dd_car != R(0), r_t = dd_light != R(0);  // Python truthiness (NaN truthy)
        const bool obs = c_t || r_t;
        // the obstacle that counts: min(d_car, d_light) when both, else whichever there is (simulation.py:113-131)
        const R dd_obs = (c_t && r_t) ? ((dd_car <= dd_light) ? dd_car : dd_light) : (c_t ? dd_car : dd_light);
        // (no obstacle: d_obs is never looked at - feed sqrt a value that stays on its fast path)
        const R d_obs = sqrt(obs ? dd_obs : R(1));
        const bool weigh = c_t && !r_t && d_obs > d_node;      // simulation.py:120-126 (here d_obs is d_car)
        auto factor = [&](bool as_obs, R d, R log_den, R inv_L) {
            if (a.stop_distance < d && d <= a.free_distance)
                return (log_stop_free(logtab, d, a.stop_distance, a.log_scale) - log_den) * inv_L;
            return (as_obs && d <= a.stop_distance) ? R(0) : R(1);
        };
        // the factor the car actually uses: obstacle factor when it has an obstacle, curvature factor otherwise
        R f = R(1);
        if (obs || hc.y != R(0))
            f = factor(obs, obs ? d_obs : d_node, obs ? a.log_stop : hc.x, obs ? a.inv_log_free_over_stop : hc.y);
        // models.weigh_factors (models.py:41-66) blends it with the curvature factor; d / free as d * (1/free): <= 1 ulp
        // in the argument
        if (weigh) {
            const R r1 = (hc.y != R(0)) ? factor(false, d_node, hc.x, hc.y) : R(1);
            f = f * cos_(d_obs * a.inv_free_distance) + r1 * sin_(d_node * a.inv_free_distance);
        }
        f = fabs(f);
        // unit(target - pos) * speed * factor (simulation.py:54-58) with one reciprocal instead of two divisions (<= 1 ulp)
        vx = v.dx * v.inv_d * a.speed_limit * f;
        vy = v.dy * v.inv_d * a.speed_limit * f;
        if (isclose0_<R>(vx, R(1.0e-5), R(0.1)) & isclose0_<R>(vy, R(1.0e-5), R(0.1))) {
            const bool acc = !r_t && (!c_t || d_obs > a.stop_distance);
            if (acc) { vx += a.default_acceleration; vy += a.default_acceleration; }
        }
        if (v.crossed) {
            const int3 q = own.route();   // (cur, end, eff)
            // new head: the point we just steered to, or the destination once no usable route point is left
            const R2 nh = (q.x + 1 >= q.z && v.path_len >= 2) ? own.dest(i) : own.target();
            own.store_cursor(i, q.x + 1);
            own.store_head(i, nh);
            if (q.x + 1 < q.y) own.store_head_curv(i, own.next_curv(q.x + 1));
            own.after_cross(q.x + 1, q.y);
        }
    } else {
        const int3 q = own.route();
        if (q.x != q.y) own.store_cursor(i, q.y);  // path becomes []
    }
    const R xn = x + vx * a.dt, yn = y + vy * a.dt;  // cars.py:89-90
    const int c = own.cell();
    own.store_pos(i, mk2(xn, yn));
    if (AUX) {
        // observation-only columns: nobody on the device reads them back - streaming stores (evict-first), so that they do
        // not push the position / bin-record lines the gathers of this and the next launch hit out of L2
        __stcs(&a.vel[i], mk2(vx, vy));
        __stcs(&a.d_node[i], d_node);
        __stcs(&a.d_car[i], dd_car != R(0) ? sqrt(dd_car) : R(0));
        __stcs(&a.d_light[i], dd_light != R(0) ? sqrt(dd_light) : dd_light);   // keeps +0 (False) / -0 (None)
        __stcs(&a.leader[i], lead >= 0 ? lead - rep_now() * a.cars_per_replica : -1);
        __stcs(&a.cell_prev[i], c);
    }
    // ---- bin of the new position + insertion into the next step's table ----
    const int cn = cell_of<R>(a.g, xn, yn);
    TB2_BOUNDS(cn >= 0 && cn < a.g.n_cells, 3);
    const int sec = own.sec();
    if (cn != c) own.store_cell(i, cn);
    // (a car that changed bin has no prefetched entry of its new bin: it simply takes the atomics, ~1 % of the cars)
    if (!a.skip_insert) insert(rep_now() * a.g.n_cells + cn, i, cn == c ? sec : TB2_EMPTY);
}

// Dependent-load waves of the plain (non-pipelined) form: (1) own state, coalesced; (2) the bin's dynamic record - table
// entry, next-table entry, go word - one sector; (3) the candidate leader's position and, if the bin has a light, its
// static record.  Everything else is arithmetic (the CSR / per-face arrays are only touched on the rare general paths).
template <typename R, bool AUX, bool ENS>
__device__ __forceinline__ void step_car(const StepArgsT<R>& a, const Rot<R>& rot, const LogEntry* __restrict__ logtab,
                                         const int i) {
    using R2 = typename real2<R>::type;
    // replica of this car (1 for a single simulation): offsets into the per-replica cell records and face states
    const int rep = ENS ? i / a.cars_per_replica : 0;
    const int cbase = rep * a.g.n_cells, fbase = rep * a.faces_per_replica;
    if (ENS && a.freeze_done && a.done[rep]) return;   // this rollout is over
    // ---- wave 1: own state ----
    const R2 p = rot.pos_in[i];
    const int cur = a.cursor[i];
    const int end = a.path_end[i];
    const int eff = a.path_eff_end[i];
    const int c = a.cell[i];
    const R2 head = a.head_xy[i];
    const R2 hc = a.head_curv[i];
    const R rt = a.route_time[i];
    TB2_BOUNDS(c >= 0 && c < a.g.n_cells, 0);
    TB2_BOUNDS(cur >= 0 && cur <= end && eff <= end, 1);
    // ---- wave 2: the bin's dynamic record (written by the previous step: bypass L1) ----
    const CellDynT* dyn = a.cell_dyn + (cbase + c);
    const int2 m = __ldcg(reinterpret_cast<const int2*>(&dyn->w[2 * rot.tab_cur_k]));
    const int sec_next = __ldcg(&dyn->w[2 * rot.tab_next_k + 1]);
    const uint32_t gw = (uint32_t)__ldcg(&dyn->w[TB2_DYN_GO + rot.go_k_cur]);
    // ---- wave 3: candidate leader = first other car of the bin (navigation.py:438-442) ----
    const int j = (m.x != i) ? m.x : m.y;
    TB2_BOUNDS(j == TB2_EMPTY || (j >= rep * a.cars_per_replica && j < (rep + 1) * a.cars_per_replica && j != i), 2);
    R2 o = mk2(R(0), R(0));
    if (j != TB2_EMPTY) o = __ldcg(&rot.pos_in[j]);

    OwnRegs<R> own;
    own.a = &a; own.pos_out = rot.pos_out; own.cur = cur; own.end = end; own.eff = eff; own.c = c; own.sec_ = sec_next; own.hc_ = hc; own.rt_ = rt;
    const View<R> v = seg_view<R>(a, i, p, head, cur, end, eff, own.t);
    R dd_car, dd_light;
    int lead;
    seg_detect<R>(a, rot.go_cur, fbase, c, v, j, o, gw, own, dd_car, lead, dd_light);
    seg_move<R, AUX, ENS>(a, rot, logtab, i, rep, v, dd_car, lead, dd_light, own,
                          [&](int rec, int ii, int sec) { table_insert(&a.cell_dyn[rec].w[2 * rot.tab_next_k], ii, sec); });
}

// every CTA keeps the (stop, free] log table in shared memory; the caller synchronises the CTA afterwards
__device__ __forceinline__ void load_log_table(LogEntry* dst, const LogEntry* __restrict__ src) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int k = (int)threadIdx.x; k < 2 * TB2_LOG_N; k += (int)blockDim.x) d4[k] = s4[k];
}

__device__ __forceinline__ void clear_table(CellDynT* dyn, int k, int n_records, int first, int stride) {
    for (int r = first; r < n_records; r += stride)
        *reinterpret_cast<int2*>(&dyn[r].w[2 * k]) = make_int2(TB2_EMPTY, TB2_EMPTY);
}

// One launch per step: CTAs [0, car_blocks) take one car per thread, the trailing CTAs tick the lights.
template <typename R, bool AUX, bool ENS>
__global__ void __launch_bounds__(TB2_BLOCK, TB2_MIN_BLOCKS) step_kernel(const __grid_constant__ StepArgsT<R> a) {
    if ((int)blockIdx.x >= a.car_blocks) {
        const int l = ((int)blockIdx.x - a.car_blocks) * TB2_BLOCK + (int)threadIdx.x;
        lights_tick_body(l, a.n_lights, a.lights_per_replica, a.faces_per_replica, a.g.n_cells, a.dt64, a.switch_time,
                         a.face_ptr, a.go_cur, a.go_next, a.t_cur, a.t_next, a.light_cellinfo, a.cell_dyn, a.go_k_cur ^ 1,
                         (ENS && a.freeze_done) ? a.done : nullptr);
        return;
    }
    const LogEntry* logtab = reinterpret_cast<const LogEntry*>(a.log_table);   // 4 KB, L1-resident
    const int i = (int)blockIdx.x * TB2_BLOCK + (int)threadIdx.x;
    // reset the table that the launch after next will fill
    clear_table(a.cell_dyn, a.tab_clear_k, a.g.n_cells * a.n_replicas, i, a.car_blocks * TB2_BLOCK);
    if (i >= a.n_cars) return;
    const Rot<R> rot{a.pos_in, a.pos_out, a.go_cur, a.tab_cur_k, a.tab_next_k, a.go_k_cur, a.step_index};
    step_car<R, AUX, ENS>(a, rot, logtab, i);
}

// ---------------------------------------------------------------------------------------------------------------
// Big worlds: the software-pipelined form of the step.  A persistent grid (TB2_PIPE_MIN_BLOCKS CTAs per SM) walks the
// cars in tiles of TB2_PIPE_BLOCK; every thread keeps the operands of its NEXT tiles in flight with cp.async while it
// computes the current one, so no segment ever waits for global memory:
//
//   group       contents (per thread)                                   issued                 consumed
//   G1(k+2)     pos, head point, cursor/end/eff_end/bin of tile k+2      end of tile k          seg_view of tile k+2
//   GL(k)       curvature constants, route_time of tile k                start of tile k        seg_move of tile k
//   G2(k+1)     dynamic bin record of tile k+1 (table entry, next-table  after seg_detect(k)    end of tile k (-> G3)
//               entry, go word) - address needs the bin id from G1(k+1)
//   G3(k+1)     candidate leader position + static light record of       end of tile k          seg_detect of tile k+1
//               tile k+1 - address needs the table entry from G2(k+1)
//
// The staged operands are thread-private slots in shared memory (structure of arrays, conflict-free), each thread waits
// only for its own cp.async groups: no block-wide barrier anywhere.  ~200 bytes of shared memory per thread replace the
// registers (and the occupancy) a load-everything-first kernel needs to cover three dependent memory waves.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem)), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Programmatic dependent launch (consecutive step launches on one stream): launch_dependents lets the NEXT launch's CTAs
// become resident as soon as this grid's CTAs retire, so that their launch latency and the part of their prologue that
// does not depend on this launch (the log table) overlap this grid's tail; griddepcontrol.wait blocks until the previous
// grid has completed and its memory is visible.  Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename R>
struct PipeSmemT {
    using R2 = typename real2<R>::type;
    R2 pos[2][TB2_PIPE_BLOCK];
    R2 head[2][TB2_PIPE_BLOCK];
    int32_t cur[2][TB2_PIPE_BLOCK], end[2][TB2_PIPE_BLOCK], eff[2][TB2_PIPE_BLOCK], cell[2][TB2_PIPE_BLOCK];
    R2 hc[1][TB2_PIPE_BLOCK];
    R rt[1][TB2_PIPE_BLOCK];
    int4 dyn[2][2][TB2_PIPE_BLOCK];   // [slot][half]: the bin's whole dynamic record (CellDynT) in two 16-byte halves
    R2 cand[TB2_PIPE_BLOCK];
    int4 st_lo[TB2_PIPE_BLOCK], st_hi[TB2_PIPE_BLOCK];   // CellStatT<R> in two 16-byte halves
    R2 tgt[2][TB2_PIPE_BLOCK];     // crossing step: the next target (next route point / destination)
    R2 ncurv[TB2_PIPE_BLOCK];      // crossing step: curvature constants of the next route point
    int32_t crossed[TB2_PIPE_BLOCK];
    LogEntry logtab[TB2_LOG_N];
    // ensembles: replica of the car in the slot, -1 if its rollout had already arrived when the tile was requested; replica
    // of the thread's NEXT tile and that replica's arrival flag (requested one tile ahead, rides in the own-state group)
    int32_t rep[2][TB2_PIPE_BLOCK], rep_next[TB2_PIPE_BLOCK], done_next[TB2_PIPE_BLOCK];
    // ensembles: global index of the car in the slot (-1: none) - with a live-replica list the loop walks the cars of the
    // replicas still running, not all of them - and the car's index within its replica for the thread's next tile
    int32_t idx[2][TB2_PIPE_BLOCK], c_next[TB2_PIPE_BLOCK];
    int32_t pend_car[TB2_PIPE_BLOCK];   // ensembles: the car whose table insertion is half done (walk positions are not cars)
    // word w of the staged record (w = 2k, 2k+1: table k; 6 + k: go word k)
    __device__ __forceinline__ const int32_t* word(int s, int tid, int w) const {
        return reinterpret_cast<const int32_t*>(&dyn[s][w >> 2][tid]) + (w & 3);
    }
    __device__ __forceinline__ int2 tab_entry(int s, int tid, int k) const { return *reinterpret_cast<const int2*>(word(s, tid, 2 * k)); }
    __device__ __forceinline__ int sec_next(int s, int tid, int k) const { return *word(s, tid, 2 * k + 1); }
    __device__ __forceinline__ int rep_of(int s, int tid, int) const { return rep[s][tid]; }
    __device__ __forceinline__ uint32_t go_word(int s, int tid, int k) const { return (uint32_t)*word(s, tid, TB2_DYN_GO + k); }
};

// ... or the thread's staging slots in shared memory (pipelined kernel)
template <typename R, typename SM, bool LATE2>
struct OwnSmemT {
    using R2 = typename real2<R>::type;
    const SM* sm;
    int s, tid;
    bool crossed;
    const StepArgsT<R>* a;
    R2* pos_out;
    int vv = 0;   // position of the car in the walk over the live replicas (ensembles with a live-replica list)
    __device__ __forceinline__ void store_route_time(int i, R v) { a->route_time[i] = v; }
    __device__ __forceinline__ void store_cursor(int i, int v) { a->cursor[i] = v; }
    __device__ __forceinline__ void store_head(int i, R2 v) { a->head_xy[i] = v; }
    __device__ __forceinline__ void store_head_curv(int i, R2 v) { a->head_curv[i] = v; }
    __device__ __forceinline__ void store_pos(int i, R2 v) { pos_out[i] = v; }
    __device__ __forceinline__ void store_cell(int i, int v) { a->cell[i] = v; }
    // the agent has arrived: also raise the flag the walk over the live-replica list looks at
    __device__ __forceinline__ void on_done() { if (a->live_list) a->done_v[vv / a->cars_per_replica] = 1; }
    __device__ __forceinline__ R2 target() const { return crossed ? sm->tgt[s][tid] : sm->head[s][tid]; }
    __device__ __forceinline__ int3 route() const { return make_int3(sm->cur[s][tid], sm->end[s][tid], sm->eff[s][tid]); }
    __device__ __forceinline__ R2 hc() const { if constexpr (LATE2) return sm->hc[s][tid]; else return sm->hc[0][tid]; }
    __device__ __forceinline__ R rt() const { if constexpr (LATE2) return sm->rt[s][tid]; else return sm->rt[0][tid]; }
    __device__ __forceinline__ int cell() const { return sm->cell[s][tid]; }
    __device__ __forceinline__ int sec() const { return sm->sec_next(s, tid, a->tab_next_k); }
    __device__ __forceinline__ int rep(int r) const { return sm->rep_of(s, tid, r); }   // re-read where it is used
    __device__ __forceinline__ R2 next_curv(int) const { return sm->ncurv[tid]; }
    __device__ __forceinline__ R2 dest(int i) const { return a->dest_xy[i]; }
    __device__ __forceinline__ void after_cross(int, int) {}
    __device__ __forceinline__ CellStatT<R> stat() const {
        union { CellStatT<R> st; int4 q[2]; } u;
        u.q[0] = sm->st_lo[tid]; u.q[1] = sm->st_hi[tid];
        return u.st;
    }
};

#ifdef TB2_PIPE_MAXNREG
#define TB2_PIPE_BOUNDS __maxnreg__(TB2_PIPE_MAXNREG)
#else
#define TB2_PIPE_BOUNDS __launch_bounds__(TB2_PIPE_BLOCK, TB2_PIPE_MIN_BLOCKS)
#endif
template <typename R, bool AUX, bool ENS>
__global__ void TB2_PIPE_BOUNDS pipe_kernel(const __grid_constant__ StepArgsT<R> a) {
    using R2 = typename real2<R>::type;
    constexpr int B = TB2_PIPE_BLOCK;
    __shared__ __align__(16) PipeSmemT<R> sm;
    const int tid = (int)threadIdx.x;
    griddep_launch_dependents();
    if ((int)blockIdx.x >= a.pipe_ctas) {   // trailing CTAs: TrafficLights.update
        const int l = ((int)blockIdx.x - a.pipe_ctas) * B + tid;
        griddep_wait();
        lights_tick_body(l, a.n_lights, a.lights_per_replica, a.faces_per_replica, a.g.n_cells, a.dt64, a.switch_time,
                         a.face_ptr, a.go_cur, a.go_next, a.t_cur, a.t_next, a.light_cellinfo, a.cell_dyn, a.go_k_cur ^ 1,
                         (ENS && a.freeze_done) ? a.done : nullptr);
        return;
    }
    const int stride = a.pipe_ctas * B;
    const Rot<R> rot{a.pos_in, a.pos_out, a.go_cur, a.tab_cur_k, a.tab_next_k, a.go_k_cur, a.step_index};

    // Ensembles: a rollout whose agent has arrived is frozen (tb2_rollout): its cars are neither loaded nor stepped.  The
    // flag is sampled once, when the tile's own-state columns are requested, and travels with the staging slot; `done` only
    // ever goes 0 -> 1 and the state of a finished rollout is never looked at again, so a stale 0 merely costs one more
    // (unobserved) step of that replica's cars.
    // The replica index costs an integer division and the flag a global load: both are produced ONE TILE AHEAD - the
    // division once per tile, the flag by a cp.async that travels with the previous tile's own-state group - so that
    // nothing here waits for memory (only the very first tile of a thread asks synchronously).
    auto slot_frozen = [&](int s) { if constexpr (ENS) return sm.rep[s][tid] < 0; else return false; };
    auto slot_rep = [&](int s) { if constexpr (ENS) return sm.rep[s][tid]; else return 0; };
    // Long sweeps: tb2_rollout hands the kernel the list of the replicas that were still running at its last read-back
    // (live_list, n_live); the loop then walks n_live * cars_per_replica "virtual" car indices - position k of the list
    // times cars_per_replica plus the car's index in its replica - instead of every car of every replica.  The list entry
    // and the arrival flag (done_v, indexed like the list) of the next tile are both fetched one tile ahead.  Everything
    // after issue_early takes the car's global index from the staging slot.
#define n_walk (a.n_walk)   /* = n_live * cars_per_replica, or n_cars without a list: read from the constant bank */
    auto car_of = [&](int vv, int s) {   // global car index of walk position vv staged in slot s, -1 if there is none
        if constexpr (ENS) return sm.idx[s][tid]; else return vv < a.n_cars ? vv : -1;
    };
    auto issue_early = [&](int vv, int s, bool first = false) {
        int ii = vv < a.n_cars ? vv : -1;
        if constexpr (ENS) {
            bool fz = false;
            int r = 0;
            ii = -1;
            if (vv < n_walk) {
                int c;
                if (first) {
                    const int rv = vv / a.cars_per_replica;
                    c = vv - rv * a.cars_per_replica;
                    r = a.live_list ? a.live_list[rv] : rv;
                    fz = a.freeze_done && (a.live_list ? a.done_v[rv] : a.done[r]);
                } else {
                    r = sm.rep_next[tid]; c = sm.c_next[tid]; fz = a.freeze_done && sm.done_next[tid] != 0;
                }
                ii = r * a.cars_per_replica + c;
                TB2_BOUNDS(r >= 0 && r < a.n_replicas && c >= 0 && c < a.cars_per_replica, 9);
            }
            sm.rep[s][tid] = fz ? -1 : r;
            sm.idx[s][tid] = ii;
            const int nx = vv + stride;   // the thread's next tile: its replica, and (asynchronously) its arrival flag
            if (nx < n_walk) {
                const int rn = nx / a.cars_per_replica;
                sm.c_next[tid] = nx - rn * a.cars_per_replica;
                if (a.live_list) {
                    cp_async<4>(&sm.rep_next[tid], &a.live_list[rn]);
                    if (a.freeze_done) cp_async<4>(&sm.done_next[tid], &a.done_v[rn]);
                } else {
                    sm.rep_next[tid] = rn;
                    if (a.freeze_done) cp_async<4>(&sm.done_next[tid], &a.done[rn]);
                }
            }
            if (fz) { cp_async_commit(); return; }
        }
        if (ii >= 0) {
            cp_async<sizeof(R2)>(&sm.pos[s][tid], &a.pos_in[ii]);
            cp_async<sizeof(R2)>(&sm.head[s][tid], &a.head_xy[ii]);
            cp_async<4>(&sm.cur[s][tid], &a.cursor[ii]);
            cp_async<4>(&sm.end[s][tid], &a.path_end[ii]);
            cp_async<4>(&sm.eff[s][tid], &a.path_eff_end[ii]);
            cp_async<4>(&sm.cell[s][tid], &a.cell[ii]);
        }
        cp_async_commit();
    };
    auto issue_late = [&](int vv, int s) {
        const int ii = car_of(vv, s);
        if (ii >= 0 && !slot_frozen(s)) {
            cp_async<sizeof(R2)>(&sm.hc[0][tid], &a.head_curv[ii]);
            cp_async<sizeof(R)>(&sm.rt[0][tid], &a.route_time[ii]);
        }
        cp_async_commit();
    };
    // (shared-memory reads come first in every issue_*: ptxas pads the first LDGSTS after an LDS with three dummy LDS)
    auto issue_dyn = [&](int vv, int s) {     // needs G1 of that tile
        const int ii = car_of(vv, s);
        if (ii >= 0 && !slot_frozen(s)) {
            const int rep = slot_rep(s);
            const int cell = sm.cell[s][tid];
            TB2_BOUNDS(cell >= 0 && cell < a.g.n_cells, 4);
            // a car that crosses its path head this step (~1.5 %) steers to the next route point: fetched with the record
            const int cur = sm.cur[s][tid], end = sm.end[s][tid];
            const bool crossed = view_crossed<R>(a, sm.pos[s][tid], sm.head[s][tid], cur, sm.eff[s][tid]);
            sm.crossed[tid] = crossed;
            // the bin's whole dynamic record (one sector): table entry, next-table entry, go words
            const char* dyn = reinterpret_cast<const char*>(a.cell_dyn) + ((size_t)(uint32_t)(rep * a.g.n_cells + cell) << 5);
            cp_async<16>(&sm.dyn[s][0][tid], dyn);
            cp_async<16>(&sm.dyn[s][1][tid], dyn + 16);
            if (crossed) cp_async<sizeof(R2)>(&sm.tgt[s][tid], crossing_target<R>(a, ii, cur, end));
        }
        cp_async_commit();
    };
    // G3 of tile `ii` (needs its G2), G1 of tile `i2` and GL of tile `ii`, committed in the order the loop consumes them
    auto issue_cand_early_late = [&](int vv, int s, int i2, int s2) {
        int j = TB2_EMPTY, cell = 0, cur = 0, end = 0;
        bool has_light = false, crossed = false;
        const int ii = car_of(vv, s);
        if (ii >= 0 && !slot_frozen(s)) {
            const int2 m = sm.tab_entry(s, tid, a.tab_cur_k);
            j = (m.x != ii) ? m.x : m.y;
            TB2_BOUNDS(j == TB2_EMPTY || (j >= 0 && j < a.n_cars && j != ii), 5);
            has_light = (sm.go_word(s, tid, a.go_k_cur) >> 16) != 0;
            crossed = sm.crossed[tid] != 0;
            cell = sm.cell[s][tid]; cur = sm.cur[s][tid]; end = sm.end[s][tid];
        }
        if (j != TB2_EMPTY) cp_async<sizeof(R2)>(&sm.cand[tid], &a.pos_in[j]);
        if (has_light) {
            const char* st = reinterpret_cast<const char*>(&a.cell_stat[cell]);
            cp_async<16>(&sm.st_lo[tid], st);
            cp_async<16>(&sm.st_hi[tid], st + 16);
        }
        if (crossed && cur + 1 < end) cp_async<sizeof(R2)>(&sm.ncurv[tid], &a.pt_curv[cur + 1]);
        cp_async_commit();
        issue_early(i2, s2);
        issue_late(vv, s);
    };

    // Warp 0 of each CTA pulls the own-state columns of the tile TB2_PIPE_L2_AHEAD iterations ahead from HBM into L2 (whole
    // tiles only), so that every cp.async below is an L2 hit: one prefetch instruction per column, lane l asks for the
    // l-th 128-byte line of the column's segment.  (The bulk form, cp.async.bulk.prefetch.L2, wants a warp-uniform address;
    // inside this loop that costs a ten-instruction waterfall per prefetch.)
    auto prefetch_tile_l2 = [&](int first) {
        if ((tid >> 5) != 0 || first + B > a.n_cars) return;
        if (ENS && a.live_list) return;   // (walk positions are not addresses; the tile is requested two tiles ahead anyway)
        const int off = (tid & 31) * 128;
        auto pf = [&](const void* base, int bytes) {
            if (off < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base) + off));
        };
        pf(&a.pos_in[first], B * (int)sizeof(R2));
        pf(&a.head_xy[first], B * (int)sizeof(R2));
        pf(&a.head_curv[first], B * (int)sizeof(R2));
        pf(&a.route_time[first], B * (int)sizeof(R));
        pf(&a.cursor[first], B * 4);
        pf(&a.path_end[first], B * 4);
        pf(&a.path_eff_end[first], B * 4);
        pf(&a.cell[first], B * 4);
    };

    int i = (int)blockIdx.x * B + tid;
    // ---- prologue: bring tile 0 to the loop invariant.  cp.async groups complete in commit order and wait_group counts
    // from the youngest, so groups are committed in the order the loop consumes them: pending = [G3(0), G1(1), GL(0)].
    // The first tile's operands are requested before anything else; the log table and the reset of the table that the
    // launch after next will fill ride in the shadow of that first HBM round trip. ----
    {   // the log table travels with the first group (no synchronous load at the head of the kernel)
        const char* src = reinterpret_cast<const char*>(a.log_table);
        char* dst = reinterpret_cast<char*>(sm.logtab);
        for (int k = tid * 16; k < (int)sizeof(sm.logtab); k += B * 16) cp_async<16>(dst + k, src + k);
    }
    griddep_wait();   // everything below reads what the previous step wrote
    issue_early(i, 0, true);
    for (int q = 1; q < TB2_PIPE_L2_AHEAD; ++q) prefetch_tile_l2((int)blockIdx.x * B + q * stride);
    if (ENS && a.freeze_done) {
        // rollouts: the tables of a replica whose agent has arrived are dead - whole replicas are skipped (late in a sweep
        // most of them are), a CTA resets the records of every pipe_ctas-th replica
        for (int rp = (int)blockIdx.x; rp < a.n_replicas; rp += a.pipe_ctas) {
            if (a.done[rp]) continue;
            CellDynT* dyn = a.cell_dyn + (size_t)rp * a.g.n_cells;
            for (int r = tid; r < a.g.n_cells; r += B)
                *reinterpret_cast<int2*>(&dyn[r].w[2 * a.tab_clear_k]) = make_int2(TB2_EMPTY, TB2_EMPTY);
        }
    } else {
        clear_table(a.cell_dyn, a.tab_clear_k, a.g.n_cells * a.n_replicas, (int)blockIdx.x * B + tid, stride);
    }
    cp_async_wait<0>();
    __syncthreads();   // every thread's slice of the log table has landed
    issue_dyn(i, 0);
    cp_async_wait<0>();
    issue_cand_early_late(i, 0, i + stride, 1);
    // table insertion of the previous tile, second half: the atomicMin that returned `pend_old` is consumed one tile later
    int pend_rec = -1, pend_old = 0;   // record index of the bin, value the first atomicMin returned
    int32_t* const tab_next = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(a.cell_dyn) + (uint32_t)a.tab_next_k * 8u);
    auto finish_insert = [&](int ii) {
        if (pend_rec < 0) return;
        if constexpr (ENS) ii = sm.pend_car[tid];
        const int loser = pend_old > ii ? pend_old : ii;   // whoever is not the bin minimum competes for second place
        if (loser != TB2_EMPTY) atomicMin(tab_next + (size_t)(uint32_t)pend_rec * 8u + 1, loser);
        pend_rec = -1;
    };
    int s = 0;
    // a single simulation: a thread past the last car has nothing left to do (there is no block-wide barrier in the
    // loop); an ensemble keeps walking - a later tile may belong to a rollout that is still running.  `i` is the position
    // in the walk, `ic` the car (the same thing unless a live-replica list is in use)
    for (; ENS ? (i - tid < n_walk) : (i < a.n_cars); i += stride, s ^= 1) {
        const int ic = ENS ? car_of(i, s) : i;
        const bool live = !ENS || (ic >= 0 && !slot_frozen(s));   // (an arrived rollout is frozen)
        prefetch_tile_l2(i - tid + TB2_PIPE_L2_AHEAD * stride);
        View<R> v{};                        // pending: G3(k), G1(k+1), GL(k)
        // the staged operands stay in shared memory; whoever needs one of them re-reads it (LDS) where it is used
        // instead of keeping it in a register across the fp64 sections
        OwnSmemT<R, PipeSmemT<R>, false> own{&sm, s, tid, false, &a, rot.pos_out, i};
        if (live) {
            own.crossed = sm.crossed[tid] != 0;
            v = view_build<R>(a, sm.pos[s][tid], own.target(), sm.cur[s][tid], sm.end[s][tid], sm.eff[s][tid], own.crossed);
        }
        finish_insert(i - stride);
        cp_async_wait<2>();                 // G3(k) landed
        R d_car = R(0), d_light = R(0);
        int lead = -1;
        if (live) {
            const int2 m = sm.tab_entry(s, tid, a.tab_cur_k);
            const int j = (m.x != ic) ? m.x : m.y;
            seg_detect<R>(a, rot.go_cur, slot_rep(s) * a.faces_per_replica, sm.cell[s][tid], v, j, sm.cand[tid],
                          sm.go_word(s, tid, a.go_k_cur), own, d_car, lead, d_light);
        }
        cp_async_wait<1>();                 // G1(k+1) landed
        issue_dyn(i + stride, s ^ 1);       // pending: GL(k), G2(k+1)
        cp_async_wait<1>();                 // GL(k) landed
        if (live)
            seg_move<R, AUX, ENS>(a, rot, sm.logtab, ic, 0, v, d_car, lead, d_light, own,
                                  [&](int rec, int ii, int sec) {
                                      if (ii > sec) return;
                                      pend_old = atomicMin(tab_next + (size_t)(uint32_t)rec * 8u, ii);
                                      pend_rec = rec;
                                      if constexpr (ENS) sm.pend_car[tid] = ii;
                                  });
        cp_async_wait<0>();                 // G2(k+1) landed
        issue_cand_early_late(i + stride, s ^ 1, i + 2 * stride, s);   // pending: G3(k+1), G1(k+2), GL(k+1)
    }
    finish_insert(i - stride);
    cp_async_wait<0>();
#undef n_walk
}

// ---------------------------------------------------------------------------------------------------------------
// TMA variant of the pipelined kernel (TB2_PIPE_TMA=1 at run time): the own-state columns of a tile travel as 1-D TMA bulk
// copies (cp.async.bulk.shared.global + mbarrier complete_tx, SASS UBLKCP / SYNCS) - every warp owns its 32-car slice,
// lane 0 issues eight bulk copies per tile into the warp's slot two tiles ahead, the other lanes only wait on the slot's
// mbarrier - instead of eight LDGSTS per thread; the data-dependent gathers (bin record, candidate, light record) stay
// per-thread cp.async.  Bit-identical (tests run it with TB2_PIPE_TMA=1).  Measured: 61 us/step vs 49 us for the LDGSTS
// form at 1M cars - the warp-level synchronisation and three more live values push the kernel into spilling at 96
// registers - so it is not the default.
// ---------------------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, 1-D): one lane moves a whole column segment of a tile, the barrier counts bytes
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TB2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TB2_DONE_%=;\n"
        "bra TB2_WAIT_%=;\n"
        "TB2_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename R>
struct PipeTmaSmemT {
    using R2 = typename real2<R>::type;
    R2 pos[2][TB2_PIPE_BLOCK];
    R2 head[2][TB2_PIPE_BLOCK];
    int32_t cur[2][TB2_PIPE_BLOCK], end[2][TB2_PIPE_BLOCK], eff[2][TB2_PIPE_BLOCK], cell[2][TB2_PIPE_BLOCK];
    R2 hc[2][TB2_PIPE_BLOCK];
    R rt[2][TB2_PIPE_BLOCK];
    unsigned long long full[2][TB2_PIPE_BLOCK / 32];   // per warp, per slot: mbarrier "own-state columns of the tile have landed"
    int2 m[2][TB2_PIPE_BLOCK];
    int32_t sec[2][TB2_PIPE_BLOCK], gw[2][TB2_PIPE_BLOCK];
    R2 cand[TB2_PIPE_BLOCK];
    int4 st_lo[TB2_PIPE_BLOCK], st_hi[TB2_PIPE_BLOCK];
    R2 tgt[2][TB2_PIPE_BLOCK];
    R2 ncurv[TB2_PIPE_BLOCK];
    int32_t crossed[TB2_PIPE_BLOCK];
    LogEntry logtab[TB2_LOG_N];
    __device__ __forceinline__ int sec_next(int s, int tid, int) const { return sec[s][tid]; }
    __device__ __forceinline__ int rep_of(int, int, int r) const { return r; }
};

template <typename R, bool AUX, bool ENS>
__global__ void TB2_PIPE_BOUNDS pipe_tma_kernel(const __grid_constant__ StepArgsT<R> a) {
    using R2 = typename real2<R>::type;
    constexpr int B = TB2_PIPE_BLOCK;
    __shared__ __align__(16) PipeTmaSmemT<R> sm;
    const int tid = (int)threadIdx.x;
    if ((int)blockIdx.x >= a.pipe_ctas) {   // trailing CTAs: TrafficLights.update
        const int l = ((int)blockIdx.x - a.pipe_ctas) * B + tid;
        lights_tick_body(l, a.n_lights, a.lights_per_replica, a.faces_per_replica, a.g.n_cells, a.dt64, a.switch_time,
                         a.face_ptr, a.go_cur, a.go_next, a.t_cur, a.t_next, a.light_cellinfo, a.cell_dyn, a.go_k_cur ^ 1);
        return;
    }
#define stride (a.pipe_ctas * B)   /* re-derived from the constant bank where it is used: one register less to carry */
    const Rot<R> rot{a.pos_in, a.pos_out, a.go_cur, a.tab_cur_k, a.tab_next_k, a.go_k_cur, a.step_index};

    // Own-state columns of a tile (pos, head point, curvature constants, route time, cursor / end / eff_end / bin): every
    // warp owns its 32-car slice of the CTA's tile and moves it with EIGHT TMA bulk copies issued by lane 0 into the
    // warp's slot; the slot's mbarrier counts the bytes.  (Segments are 16-byte multiples; the last, partial tile rounds up
    // inside the 256-byte padding of the arena's sub-buffers.)  Tiles past the end just arrive.
    auto produce = [&](int first_of_warp, int s) {
        if ((tid & 31) != 0) return;
        const int warp = tid >> 5;
        unsigned long long* bar = &sm.full[s][warp];
        int rows = a.n_cars - first_of_warp;
        rows = rows < 0 ? 0 : (rows > 32 ? 32 : rows);
        if (rows == 0) { mbar_arrive(bar); return; }
        const int b2 = (rows * (int)sizeof(R2) + 15) & ~15, b1 = (rows * (int)sizeof(R) + 15) & ~15, b4 = (rows * 4 + 15) & ~15;
        mbar_expect_tx(bar, 3 * b2 + b1 + 4 * b4);
        const int o = warp * 32;
        bulk_g2s(&sm.pos[s][o], &a.pos_in[first_of_warp], b2, bar);
        bulk_g2s(&sm.head[s][o], &a.head_xy[first_of_warp], b2, bar);
        bulk_g2s(&sm.hc[s][o], &a.head_curv[first_of_warp], b2, bar);
        bulk_g2s(&sm.rt[s][o], &a.route_time[first_of_warp], b1, bar);
        bulk_g2s(&sm.cur[s][o], &a.cursor[first_of_warp], b4, bar);
        bulk_g2s(&sm.end[s][o], &a.path_end[first_of_warp], b4, bar);
        bulk_g2s(&sm.eff[s][o], &a.path_eff_end[first_of_warp], b4, bar);
        bulk_g2s(&sm.cell[s][o], &a.cell[first_of_warp], b4, bar);
    };
    auto issue_dyn = [&](int ii, int s) {     // needs G1 of that tile
        if (ii < a.n_cars) {
            const int rep = ENS ? ii / a.cars_per_replica : 0;
            TB2_BOUNDS(sm.cell[s][tid] >= 0 && sm.cell[s][tid] < a.g.n_cells, 4);
            const CellDynT* dyn = a.cell_dyn + (rep * a.g.n_cells + sm.cell[s][tid]);
            cp_async<8>(&sm.m[s][tid], &dyn->w[2 * a.tab_cur_k]);
            cp_async<4>(&sm.sec[s][tid], &dyn->w[2 * a.tab_next_k + 1]);
            cp_async<4>(&sm.gw[s][tid], &dyn->w[TB2_DYN_GO + a.go_k_cur]);
            // a car that crosses its path head this step (~1.5 %) steers to the next route point: fetch it now
            const int cur = sm.cur[s][tid];
            const bool crossed = view_crossed<R>(a, sm.pos[s][tid], sm.head[s][tid], cur, sm.eff[s][tid]);
            sm.crossed[tid] = crossed;
            if (crossed) cp_async<sizeof(R2)>(&sm.tgt[s][tid], crossing_target<R>(a, ii, cur, sm.end[s][tid]));
        }
        cp_async_commit();
    };
    auto issue_cand = [&](int ii, int s) {    // needs G2 of that tile
        if (ii < a.n_cars) {
            const int2 m = sm.m[s][tid];
            const int j = (m.x != ii) ? m.x : m.y;
            TB2_BOUNDS(j == TB2_EMPTY || (j >= 0 && j < a.n_cars && j != ii), 5);
            if (j != TB2_EMPTY) cp_async<sizeof(R2)>(&sm.cand[tid], &a.pos_in[j]);
            if (((uint32_t)sm.gw[s][tid]) >> 16) {
                const char* st = reinterpret_cast<const char*>(&a.cell_stat[sm.cell[s][tid]]);
                cp_async<16>(&sm.st_lo[tid], st);
                cp_async<16>(&sm.st_hi[tid], st + 16);
            }
            if (sm.crossed[tid]) {
                const int cur = sm.cur[s][tid];
                if (cur + 1 < sm.end[s][tid]) cp_async<sizeof(R2)>(&sm.ncurv[tid], &a.pt_curv[cur + 1]);
            }
        }
        cp_async_commit();
    };

    int i = (int)blockIdx.x * B + tid;
    // ---- prologue ----
    const int lane = tid & 31;
    if (lane == 0) {
        mbar_init(&sm.full[0][tid >> 5], 1);
        mbar_init(&sm.full[1][tid >> 5], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    produce(i - lane, 0);
    produce(i - lane + stride, 1);
    {   // the log table travels as the first cp.async group (no synchronous load at the head of the kernel)
        const char* src = reinterpret_cast<const char*>(a.log_table);
        char* dst = reinterpret_cast<char*>(sm.logtab);
        for (int k = tid * 16; k < (int)sizeof(sm.logtab); k += B * 16) cp_async<16>(dst + k, src + k);
        cp_async_commit();
    }
    // the reset of the table that the launch after next will fill rides in the shadow of the first HBM round trip
    clear_table(a.cell_dyn, a.tab_clear_k, a.g.n_cells * a.n_replicas, (int)blockIdx.x * B + tid, stride);
    cp_async_wait<0>();
    __syncthreads();   // every thread's slice of the log table has landed
    mbar_wait(&sm.full[0][tid >> 5], 0);
    issue_dyn(i, 0);
    cp_async_wait<0>();
    issue_cand(i, 0);                       // per-thread cp.async groups pending: [G3(0)]
    // table insertion of the previous tile, second half: the atomicMin that returned `pend_old` is consumed one tile later
    int pend_rec = -1, pend_old = 0;   // record index of the bin, value the first atomicMin returned
    auto finish_insert = [&](int ii) {
        if (pend_rec < 0) return;
        const int loser = pend_old > ii ? pend_old : ii;   // whoever is not the bin minimum competes for second place
        if (loser != TB2_EMPTY) atomicMin(&a.cell_dyn[pend_rec].w[2 * a.tab_next_k + 1], loser);
        pend_rec = -1;
    };
    int s = 0;
    int par = 1;   // bit x: parity of the next phase of slot x's barrier to wait for (slot 0 has been waited for once)
    for (; i - tid < a.n_cars; i += stride, s ^= 1) {
        const int rep = ENS ? i / a.cars_per_replica : 0;
        const bool live = i < a.n_cars && !(ENS && a.freeze_done && a.done[rep]);   // (an arrived rollout is frozen)
        View<R> v{};
        // the staged operands stay in shared memory; whoever needs one of them re-reads it (LDS) where it is used
        // instead of keeping it in a register across the fp64 sections
        OwnSmemT<R, PipeTmaSmemT<R>, true> own{&sm, s, tid, false, &a, rot.pos_out};
        if (live) {
            own.crossed = sm.crossed[tid] != 0;
            v = view_build<R>(a, sm.pos[s][tid], own.target(), sm.cur[s][tid], sm.end[s][tid], sm.eff[s][tid], own.crossed);
        }
        finish_insert(i - stride);
        cp_async_wait<0>();                 // G3(k) landed
        R d_car = R(0), d_light = R(0);
        int lead = -1;
        if (live) {
            const int2 m = sm.m[s][tid];
            const int j = (m.x != i) ? m.x : m.y;
            seg_detect<R>(a, rot.go_cur, rep * a.faces_per_replica, sm.cell[s][tid], v, j, sm.cand[tid],
                          (uint32_t)sm.gw[s][tid], own, d_car, lead, d_light);
        }
        mbar_wait(&sm.full[s ^ 1][tid >> 5], (par >> (s ^ 1)) & 1);   // own state of tile k+1 landed
        par ^= 1 << (s ^ 1);
        issue_dyn(i + stride, s ^ 1);       // pending: [G2(k+1)]
        if (live)
            seg_move<R, AUX, ENS>(a, rot, sm.logtab, i, rep, v, d_car, lead, d_light, own,
                                  [&](int rec, int ii, int sec) {
                                      if (ii > sec) return;
                                      pend_old = atomicMin(&a.cell_dyn[rec].w[2 * a.tab_next_k], ii);
                                      pend_rec = rec;
                                  });
        __syncwarp();                       // every lane is done with slot s ...
        produce(i - lane + 2 * stride, s);  // ... which now receives tile k+2
        cp_async_wait<0>();                 // G2(k+1) landed
        issue_cand(i + stride, s ^ 1);      // pending: [G3(k+1)]
    }
    finish_insert(i - stride);
    cp_async_wait<0>();
#undef stride
}

// Small worlds (<= 4096 cars): ALL steps of a tb2_step batch in ONE launch.  The CTAs form a single thread-block cluster
// and meet at a hardware cluster barrier (release/acquire, so the positions, table entries and face states written in
// step s are visible to every CTA in step s+1) instead of returning to the host between steps - at 2k cars a step is a
// single wave of warps and the per-launch gap is a third of the step time.
// GRID = false: the CTAs are one thread-block cluster and meet at the hardware cluster barrier (<= 16 CTAs).
// GRID = true : cooperative launch, every CTA co-resident, grid-wide barrier (mid-size worlds, one wave of CTAs).
template <bool GRID>
__device__ __forceinline__ void step_barrier() {
    namespace cg = cooperative_groups;
    if constexpr (GRID) cg::this_grid().sync();
    else cg::this_cluster().sync();
}

template <typename R, bool ENS, bool GRID>
__global__ void __launch_bounds__(TB2_PERSIST_BLOCK, GRID ? 2 : 1)
persistent_kernel(const __grid_constant__ StepArgsT<R> a, const __grid_constant__ PersistArgsT<R> pa) {
    const int nthreads = (int)(gridDim.x * blockDim.x);
    const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int n_rec = a.g.n_cells * a.n_replicas;
    const int n_tick = a.n_lights > 0 ? a.n_lights : 1;   // thread 0 advances the clock even without lights
    int cp = pa.cur_pos, ct = pa.cur_tab, cgo = pa.cur_go, ck = pa.cur_t;
    __shared__ LogEntry logtab[TB2_LOG_N];
    load_log_table(logtab, reinterpret_cast<const LogEntry*>(a.log_table));
    __syncthreads();
    if (pa.order == TB2K_ORDER_LIGHTS_CARS) {               // L (C L) (C L) ... C
        for (int lg = gtid; lg < n_tick; lg += nthreads)
            lights_tick_body(lg, a.n_lights, a.lights_per_replica, a.faces_per_replica, a.g.n_cells, a.dt64, a.switch_time,
                             a.face_ptr, pa.go[cgo], pa.go[cgo ^ 1], pa.t[ck], pa.t[ck ^ 1], a.light_cellinfo, a.cell_dyn,
                             cgo ^ 1);
        cgo ^= 1; ck ^= 1;
        step_barrier<GRID>();
    }
    for (int s = 0; s < pa.n_steps; ++s) {
        const bool tick = pa.order != TB2K_ORDER_LIGHTS_CARS || s < pa.n_steps - 1;
        clear_table(a.cell_dyn, (ct + 2) % 3, n_rec, gtid, nthreads);
        if (gtid < a.n_cars) {
            const Rot<R> rot{pa.pos[cp], pa.pos[cp ^ 1], pa.go[cgo], ct, (ct + 1) % 3, cgo, a.step_index + s};
            if (pa.aux_last && s == pa.n_steps - 1) step_car<R, true, ENS>(a, rot, logtab, gtid);
            else step_car<R, false, ENS>(a, rot, logtab, gtid);
        }
        if (tick) {
            for (int lg = gtid; lg < n_tick; lg += nthreads)
                lights_tick_body(lg, a.n_lights, a.lights_per_replica, a.faces_per_replica, a.g.n_cells, a.dt64,
                                 a.switch_time, a.face_ptr, pa.go[cgo], pa.go[cgo ^ 1], pa.t[ck], pa.t[ck ^ 1],
                                 a.light_cellinfo, a.cell_dyn, cgo ^ 1);
            cgo ^= 1; ck ^= 1;
        }
        cp ^= 1; ct = (ct + 1) % 3;
        step_barrier<GRID>();
    }
}

// prepare, pass 1: curvature constants of every polyline point still ahead of its car
template <typename R>
__global__ void __launch_bounds__(128) prepare_points_kernel(int n_cars, const int32_t* __restrict__ cursor,
                                                             const int32_t* __restrict__ path_end,
                                                             const typename real2<R>::type* __restrict__ path_xy,
                                                             typename real2<R>::type* pt_curv, R stop_d, R free_d) {
    const int i = (int)blockIdx.x * 128 + (int)threadIdx.x;
    if (i >= n_cars) return;
    const int end = path_end[i];
    for (int q = cursor[i]; q < end; ++q) pt_curv[q] = curv_constants<R>(path_xy, q, end, stop_d, free_d);
}

// prepare, pass 2: per-car head cache, bin, cell-table insertion
template <typename R>
__global__ void __launch_bounds__(256) prepare_cars_kernel(int n_cars, GridArgsT<R> g,
                                                           const typename real2<R>::type* __restrict__ pos,
                                                           const int32_t* __restrict__ cursor,
                                                           const int32_t* __restrict__ path_eff_end,
                                                           const typename real2<R>::type* __restrict__ dest_xy,
                                                           const typename real2<R>::type* __restrict__ path_xy,
                                                           const typename real2<R>::type* __restrict__ pt_curv,
                                                           typename real2<R>::type* head_xy,
                                                           typename real2<R>::type* head_curv, int32_t* cell,
                                                           int32_t* cell_prev, CellDynT* cell_dyn, int table_k,
                                                           int cars_per_replica) {
    const int i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (i >= n_cars) return;
    const int cur = cursor[i];
    // head cache: the current route point while the car has a route (navigation.py:37-39 rule), else the destination
    typename real2<R>::type h = dest_xy[i], hc = mk2(R(0), R(0));
    if (cur < path_eff_end[i]) { h = path_xy[cur]; hc = pt_curv[cur]; }
    head_xy[i] = h;
    head_curv[i] = hc;
    const typename real2<R>::type p = pos[i];
    const int cn = cell_of<R>(g, p.x, p.y);
    cell[i] = cn;
    cell_prev[i] = cn;
    int32_t* t = &cell_dyn[(size_t)(i / cars_per_replica) * g.n_cells + cn].w[2 * table_k];
    table_insert(t, i, __ldcg(t + 1));
}

// prepare: unit vectors of the light faces (models.unit_vector, models.py:74-76)
template <typename R>
__global__ void __launch_bounds__(256) prepare_faces_kernel(int n_faces, const typename real2<R>::type* __restrict__ face_vec,
                                                            typename real2<R>::type* __restrict__ face_unit) {
    const int f = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (f >= n_faces) return;
    const typename real2<R>::type fv = face_vec[f];
    const R nf = norm2<R>(fv.x, fv.y);
    face_unit[f] = mk2(fv.x / nf, fv.y / nf);
}

// prepare: per-cell static light records (first light inline) + per-light cell info (static bits of the go word)
template <typename R>
__global__ void __launch_bounds__(256) prepare_cells_kernel(int n_cells, const int32_t* __restrict__ cell_light_ptr,
                                                            const int32_t* __restrict__ cell_light_idx,
                                                            const typename real2<R>::type* __restrict__ light_xy,
                                                            const int32_t* __restrict__ face_ptr,
                                                            const typename real2<R>::type* __restrict__ face_unit,
                                                            CellStatT<R>* out, int2* light_cellinfo) {
    const int c = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (c >= n_cells) return;
    CellStatT<R> r{};
    const int lb = cell_light_ptr[c], le = cell_light_ptr[c + 1];
    const int nl = le - lb;
    r.xy0 = mk2(R(0), R(0));
    for (int k = 0; k < 8; ++k) r.fu[k] = 0;
    for (int q = lb; q < le; ++q) {
        const int l = cell_light_idx[q];
        const int fb = face_ptr[l], nf = face_ptr[l + 1] - fb;
        if (q == lb) {
            r.xy0 = light_xy[l];
            for (int k = 0; k < nf && k < 4; ++k) {   // NaN units (zero-length face vector) quantise to 0: never anti-parallel
                r.fu[2 * k] = (int16_t)__float2int_rn((float)face_unit[fb + k].x * 32767.0f);
                r.fu[2 * k + 1] = (int16_t)__float2int_rn((float)face_unit[fb + k].y * 32767.0f);
            }
        }
        const uint32_t hi = ((uint32_t)(nl < 65535 ? nl : 65535) << 16) | ((uint32_t)(nf < 255 ? nf : 255) << 8);
        light_cellinfo[l] = make_int2(q == lb ? c : -1, (int)hi);
    }
    out[c] = r;
}

// tb2_prepare: the caller-supplied index arrays are dereferenced by every kernel - check them once (bit k of *bad = rule k)
static __global__ void __launch_bounds__(256) validate_kernel(int n_cars, long long n_points, int n_lights, int n_faces,
                                                              int n_cells, int n_cell_lights, const int32_t* cursor,
                                                              const int32_t* path_end, const int32_t* path_eff_end,
                                                              const int32_t* face_ptr, const int32_t* cell_light_ptr,
                                                              const int32_t* cell_light_idx, int* bad) {
    const int k = (int)(blockIdx.x * 256 + threadIdx.x);
    int b = 0;
    if (k < n_cars) {
        const long long c = cursor[k], e = path_end[k], f = path_eff_end[k];
        if (c < 0 || e < c || e > n_points) b |= 1;            // 0 <= cursor <= path_end <= P
        if (f < 0 || f > e) b |= 2;                            // path_eff_end <= path_end
    }
    if (k < n_lights) {
        const int f0 = face_ptr[k], f1 = face_ptr[k + 1];
        if (f0 < 0 || f1 < f0 || f1 > n_faces) b |= 4;         // face_ptr monotone within [0, F]
        if (f1 - f0 >= 32768) b |= 8;                          // faces per light < 2^15 (go word / cell record fields)
    }
    if (k == 0 && n_lights > 0 && (face_ptr[0] != 0 || face_ptr[n_lights] != n_faces)) b |= 4;
    if (k < n_cells) {
        const int l0 = cell_light_ptr[k], l1 = cell_light_ptr[k + 1];
        if (l0 < 0 || l1 < l0 || l1 > n_cell_lights) b |= 16;  // cell_light_ptr monotone within [0, LC]
        if (l1 - l0 >= 32768) b |= 8;
    }
    if (k < n_cell_lights && (cell_light_idx[k] < 0 || cell_light_idx[k] >= n_lights)) b |= 32;
    if (b) atomicOr(bad, b);
}

// (stop, free] log table (see log_stop_free)
static __global__ void build_log_table_kernel(LogEntry* tab, double stop, double free_d) {
    const int k = (int)threadIdx.x;
    if (k >= TB2_LOG_N) return;
    const double w = (free_d - stop) / (double)TB2_LOG_N;
    LogEntry e;
    e.c = stop + ((double)k + 0.5) * w;
    e.inv_c = 1.0 / e.c;
    e.log_c = log(e.c);
    e.pad = 0.0;
    tab[k] = e;
}
// all tables empty, no go words
static __global__ void init_dyn_kernel(CellDynT* dyn, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        int4* q = reinterpret_cast<int4*>(&dyn[k]);
        q[0] = make_int4(TB2_EMPTY, TB2_EMPTY, TB2_EMPTY, TB2_EMPTY);
        q[1] = make_int4(TB2_EMPTY, TB2_EMPTY, 0, 0);
    }
}
static __global__ void fill_i32_kernel(int32_t* p, size_t n, int32_t v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}
// go words of the current face-state buffer, for the lights that are the first of their bin
static __global__ void __launch_bounds__(256) rebuild_go_words_kernel(int n_lights_total, int lights_per_replica,
                                                                      int faces_per_replica, int n_cells,
                                                                      const int32_t* __restrict__ face_ptr,
                                                                      const uint8_t* __restrict__ go_cur,
                                                                      const int2* __restrict__ light_cellinfo,
                                                                      CellDynT* cell_dyn, int go_k) {
    const int lg = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (lg >= n_lights_total) return;
    const int rep = lg / lights_per_replica, l = lg - rep * lights_per_replica;
    const int2 info = light_cellinfo[l];
    if (info.x < 0) return;
    const uint32_t mask = go_mask_of(go_cur + (size_t)rep * faces_per_replica, face_ptr[l], face_ptr[l + 1]);
    cell_dyn[(size_t)rep * n_cells + info.x].w[TB2_DYN_GO + go_k] = (int32_t)((uint32_t)info.y | mask);
}

template <typename R>
void launch_step_t(const StepArgsT<R>& a, cudaStream_t stream) {
    const bool ens = a.n_replicas > 1 || a.agent_car >= 0;
    if (a.pipe_ctas > 0) {   // big world: persistent software-pipelined grid
        const int lb = (a.n_lights + TB2_PIPE_BLOCK - 1) / TB2_PIPE_BLOCK;
        const int light_blocks = a.tick ? (lb > 0 ? lb : 1) : 0;
        const int grid = a.pipe_ctas + light_blocks;
        if (a.pipe_tma) {   // TMA bulk-copy variant of the same pipeline (TB2_PIPE_TMA=1)
            if (ens) {
                if (a.write_aux) pipe_tma_kernel<R, true, true><<<grid, TB2_PIPE_BLOCK, 0, stream>>>(a);
                else pipe_tma_kernel<R, false, true><<<grid, TB2_PIPE_BLOCK, 0, stream>>>(a);
            } else {
                if (a.write_aux) pipe_tma_kernel<R, true, false><<<grid, TB2_PIPE_BLOCK, 0, stream>>>(a);
                else pipe_tma_kernel<R, false, false><<<grid, TB2_PIPE_BLOCK, 0, stream>>>(a);
            }
            return;
        }
        // launched with programmatic stream serialization (TB2_NO_PDL=1 turns it off): see griddep_wait()
        static const bool pdl = !(std::getenv("TB2_NO_PDL") && std::getenv("TB2_NO_PDL")[0] == '1');
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TB2_PIPE_BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        if (ens) {
            if (a.write_aux) cudaLaunchKernelEx(&cfg, pipe_kernel<R, true, true>, a);
            else cudaLaunchKernelEx(&cfg, pipe_kernel<R, false, true>, a);
        } else {
            if (a.write_aux) cudaLaunchKernelEx(&cfg, pipe_kernel<R, true, false>, a);
            else cudaLaunchKernelEx(&cfg, pipe_kernel<R, false, false>, a);
        }
        return;
    }
    const int lb = (a.n_lights + TB2_BLOCK - 1) / TB2_BLOCK;
    const int light_blocks = a.tick ? (lb > 0 ? lb : 1) : 0;
    const int grid = a.car_blocks + light_blocks;
    if (grid <= 0) return;
    if (ens) {
        if (a.write_aux) step_kernel<R, true, true><<<grid, TB2_BLOCK, 0, stream>>>(a);
        else step_kernel<R, false, true><<<grid, TB2_BLOCK, 0, stream>>>(a);
    } else {
        if (a.write_aux) step_kernel<R, true, false><<<grid, TB2_BLOCK, 0, stream>>>(a);
        else step_kernel<R, false, false><<<grid, TB2_BLOCK, 0, stream>>>(a);
    }
}

// How many cars the cooperative (grid-barrier) variant can take on this device: every CTA must be co-resident.
template <typename R>
int persistent_grid_capacity_t() {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_kernel<R, false, true>, TB2_PERSIST_BLOCK, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return sms * per_sm * TB2_PERSIST_BLOCK;
}

template <typename R>
cudaError_t launch_persistent_t(const StepArgsT<R>& a, const PersistArgsT<R>& pa, cudaStream_t stream) {
    int ctas = (a.n_cars + TB2_PERSIST_BLOCK - 1) / TB2_PERSIST_BLOCK;
    if (ctas < 1) ctas = 1;
    const bool ens = a.n_replicas > 1 || a.agent_car >= 0;
    if (ctas > TB2_PERSIST_MAX_CTAS) {   // mid-size world: cooperative launch + grid barrier
        void (*kern)(const StepArgsT<R>, const PersistArgsT<R>) =
            ens ? persistent_kernel<R, true, true> : persistent_kernel<R, false, true>;
        void* args[2] = {(void*)&a, (void*)&pa};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas, 1, 1), dim3(TB2_PERSIST_BLOCK, 1, 1), args, 0, stream);
    }
    void (*kern)(const StepArgsT<R>, const PersistArgsT<R>) =
        ens ? persistent_kernel<R, true, false> : persistent_kernel<R, false, false>;
    if (ctas > 8) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas, 1, 1);
    cfg.blockDim = dim3(TB2_PERSIST_BLOCK, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a, pa);
}

template <typename R>
void launch_prepare_t(const StepArgsT<R>& a, typename real2<R>::type* pt_curv, typename real2<R>::type* face_unit,
                      int n_faces, cudaStream_t stream) {
    const size_t n = (size_t)a.g.n_cells * a.n_replicas;
    const size_t blocks = (n + 255) / 256;
    init_dyn_kernel<<<(int)(blocks < 1184 ? (blocks ? blocks : 1) : 1184), 256, 0, stream>>>(a.cell_dyn, n);
    build_log_table_kernel<<<1, TB2_LOG_N, 0, stream>>>(reinterpret_cast<LogEntry*>(const_cast<void*>(a.log_table)),
                                                        (double)a.stop_distance, (double)a.free_distance);
    if (n_faces > 0) prepare_faces_kernel<R><<<(n_faces + 255) / 256, 256, 0, stream>>>(n_faces, a.face_vec, face_unit);
    const int L = a.lights_per_replica > 0 && a.n_lights > 0 ? a.n_lights / a.n_replicas : 0;
    if (L > 0) fill_i32_kernel<<<(2 * L + 255) / 256, 256, 0, stream>>>(reinterpret_cast<int32_t*>(a.light_cellinfo), (size_t)2 * L, -1);
    prepare_cells_kernel<R><<<(a.g.n_cells + 255) / 256, 256, 0, stream>>>(a.g.n_cells, a.cell_light_ptr, a.cell_light_idx,
                                                                          a.light_xy, a.face_ptr, face_unit, a.cell_stat,
                                                                          a.light_cellinfo);
    if (a.n_lights > 0)
        rebuild_go_words_kernel<<<(a.n_lights + 255) / 256, 256, 0, stream>>>(a.n_lights, a.lights_per_replica,
                                                                               a.faces_per_replica, a.g.n_cells, a.face_ptr,
                                                                               a.go_cur, a.light_cellinfo, a.cell_dyn, a.go_k_cur);
    if (a.n_cars > 0) {
        prepare_points_kernel<R><<<(a.n_cars + 127) / 128, 128, 0, stream>>>(a.n_cars, a.cursor, a.path_end, a.path_xy,
                                                                               pt_curv, a.stop_distance, a.free_distance);
        prepare_cars_kernel<R><<<(a.n_cars + 255) / 256, 256, 0, stream>>>(
            a.n_cars, a.g, a.pos_in, a.cursor, a.path_eff_end, a.dest_xy, a.path_xy, pt_curv, a.head_xy, a.head_curv, a.cell,
            a.cell_prev, a.cell_dyn, a.tab_cur_k, a.cars_per_replica);
    }
}

}  // namespace tb2k
