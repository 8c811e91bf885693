// Stand-in for boost::math::digamma / trigamma (BH is absent from the build image): x > 0 on the VEM path.
// Upward recurrence to x >= 10, then the asymptotic (Bernoulli) series -- the series Boost uses for large arguments;
// agreement with scipy.special.digamma / polygamma(1, .) to <= 2e-15 relative is checked in tests/test_ref_build.py.
#ifndef PCMF_REF_SHIM_DIGAMMA_HPP
#define PCMF_REF_SHIM_DIGAMMA_HPP
#include <cmath>
namespace boost { namespace math {
inline double digamma(double x) {
    double r = 0.0;
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
    return r + std::log(x) - 0.5 / x - f * s;
}
inline double trigamma(double x) {
    double r = 0.0;
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
}}
#endif
