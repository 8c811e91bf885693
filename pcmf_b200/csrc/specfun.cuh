// Device special functions for the VEM core, FP64, evaluated in registers.
// Each mirrors a reference primitive (file:line cited) so the parity tests can compare value for value.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pcmf {

// boost::math::digamma as used by probability.h:107 / internal.h:64 (x > 0 on the VEM path).
// Upward recurrence to x >= 10, then the asymptotic Bernoulli series (the series Boost itself uses for x >= 10).
__device__ __forceinline__ double digamma_pos(double x) {
    double r = 0.0;
#pragma unroll 1
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    double s = 3617.0 / 8160.0;
    s = 1.0 / 12.0 - f * s;
    s = 691.0 / 32760.0 - f * s;
    s = 1.0 / 132.0 - f * s;
    s = 1.0 / 240.0 - f * s;
    s = 1.0 / 252.0 - f * s;
    s = 1.0 / 120.0 - f * s;
    s = 1.0 / 12.0 - f * s;
    return r + log(x) - 0.5 / x - f * s;
}

// boost::math::trigamma as used by internal.h:69
__device__ __forceinline__ double trigamma_pos(double x) {
    double r = 0.0;
#pragma unroll 1
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    const double f = 1.0 / (x * x);
    double s = 7.0 / 6.0;
    s = 691.0 / 2730.0 - f * s;
    s = 5.0 / 66.0 - f * s;
    s = 1.0 / 30.0 - f * s;
    s = 1.0 / 42.0 - f * s;
    s = 1.0 / 30.0 - f * s;
    s = 1.0 / 6.0 - f * s;
    return r + 1.0 / x + 0.5 * f + (1.0 / x) * f * s;
}

// internal::digammaInv (internal.h:56-74) with the reference's start rule and a fixed Newton count (macros.h:47 -> 6)
__device__ __forceinline__ double digamma_inv(double y, int nb_iter) {
    const double psi1 = -0.57721566490153286061;   // digamma(1)
    double x0 = (y >= -2.22) ? exp(y) + 0.5 : -1.0 / (y - psi1);
    double x = 0.0;
    for (int i = 0; i < nb_iter; ++i) {
        x = x0 - (digamma_pos(x0) - y) / trigamma_pos(x0);
        x0 = x;
    }
    return x;
}

// internal::custom_exp (internal.h:121-127)
__device__ __forceinline__ double custom_exp(double x) { return x > -100.0 ? exp(x) : 3e-44; }
// internal::custom_log (internal.h:104-110)
__device__ __forceinline__ double custom_log(double x) { return x > 0.0 ? log(x) : 0.0; }
// internal::logit (internal.h:137-141)
__device__ __forceinline__ double logit_clamped(double x) {
    if (x >= 1.0 - 1e-12) return 30.0;
    if (x <= 1e-12) return -30.0;
    return log(x / (1.0 - x));
}
// internal::expit (internal.h:151-155)
__device__ __forceinline__ double expit_clamped(double x) {
    if (x >= 30.0) return 1.0;
    if (x <= -30.0) return 0.0;
    return 1.0 / (1.0 + exp(-x));
}
// probability::gamma_entropy, scalar (probability.h:123-126); psi = digamma(p1) is passed in because the caller has it
__device__ __forceinline__ double gamma_entropy1(double p1, double p2, double psi) {
    return (1.0 - p1) * psi + p1 + lgamma(p1) - log(p2);
}

// ---- fast FP64 exp / reciprocal for the dense zero-inflation passes (the FP64 pipe is their roofline) --------
// exp(x) for |x| <= 700: n = rint(x log2 e), r = x - n ln2 (two-term Cody-Waite), degree-12 Taylor in |r| <= 0.347
// (truncation 1.7e-16) evaluated with Estrin's scheme (dependency depth 5 instead of 12: these kernels are bound by
// FP64 dependent-issue latency, not by instruction count), result scaled by 2^n through the exponent field.
__device__ __forceinline__ double fast_exp(double x) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to nearest integer in the low word
    double t = fma(x, 1.4426950408889634074, MAGIC);
    const int n = __double2loint(t);
    t -= MAGIC;
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
    const double r2 = r * r;
    const double p01 = 1.0 + r;
    const double p23 = fma(1.66666666666666666667e-01, r, 0.5);                           // 1/2! + r/3!
    const double p45 = fma(8.33333333333333333333e-03, r, 4.16666666666666666667e-02);    // 1/4! + r/5!
    const double p67 = fma(1.98412698412698412698e-04, r, 1.38888888888888888889e-03);    // 1/6! + r/7!
    const double p89 = fma(2.75573192239858906526e-06, r, 2.48015873015873015873e-05);    // 1/8! + r/9!
    const double pab = fma(2.50521083854417187751e-08, r, 2.75573192239858906526e-07);    // 1/10! + r/11!
    const double r4 = r2 * r2;
    const double q0 = fma(r2, p23, p01);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double r8 = r4 * r4;
    const double s0 = fma(r4, q1, q0);
    const double s1 = fma(r4, 2.08767569878680989792e-09, q2);                            // + r^4 / 12!
    const double p = fma(r8, s1, s0);
    return p * __hiloint2double((n + 1023) << 20, 0);
}
// 1/y for normal y: MUFU.RCP64H seed (relative error e <= 2^-23) + one third-order step r (1 + e + e^2) (error e^3)
__device__ __forceinline__ double fast_rcp(double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    const double e = fma(-y, r, 1.0);
    const double e2 = fma(e, e, e);
    return fma(r, e2, r);
}
// internal::expit (internal.h:151-155) with the fast primitives; also returns y = 1 + exp(-z) (1 when clamped) and
// whether the result is strictly inside (0,1).  The |z| >= 30 clamps are tested on the exponent word (integer pipe).
__device__ __forceinline__ double fast_expit(double z, double &y_out, bool &interior) {
    const int hi = __double2hiint(z);
    const bool big = (hi & 0x7fffffff) >= 0x403E0000;     // |z| >= 30 (30.0 = 0x403E000000000000), also NaN/inf
    const double zc = big ? 0.0 : z;
    const double y = 1.0 + fast_exp(-zc);
    const double pr = fast_rcp(y);
    interior = !big;
    y_out = big ? 1.0 : y;
    return big ? ((hi < 0) ? 0.0 : 1.0) : pr;
}

// ---- warp helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reduce-scatter of 32 per-lane values: on return v[0] of lane l holds sum over lanes of the input v[l].
// 31 shuffles instead of 32 x 5 (used for the column sums of prob_D in the dense ZI pass).
__device__ __forceinline__ double warp_reduce_scatter32(double (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const double send = up ? v[i] : v[i + s];
            const double keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// FP32 flavour (FP32-mode dense pass, zi_tc.cuh)
__device__ __forceinline__ float warp_reduce_scatter32f(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// double atomic min/max through the ordered-integer trick (values are finite)
__device__ __forceinline__ long long dbl_ordered(double d) {
    long long i = __double_as_longlong(d);
    return i >= 0 ? i : i ^ 0x7fffffffffffffffLL;
}
__device__ __forceinline__ double dbl_from_ordered(long long i) {
    return __longlong_as_double(i >= 0 ? i : i ^ 0x7fffffffffffffffLL);
}

}  // namespace pcmf
