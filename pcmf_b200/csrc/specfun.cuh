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

// double atomic min/max through the ordered-integer trick (values are finite)
__device__ __forceinline__ long long dbl_ordered(double d) {
    long long i = __double_as_longlong(d);
    return i >= 0 ? i : i ^ 0x7fffffffffffffffLL;
}
__device__ __forceinline__ double dbl_from_ordered(long long i) {
    return __longlong_as_double(i >= 0 ? i : i ^ 0x7fffffffffffffffLL);
}

}  // namespace pcmf
