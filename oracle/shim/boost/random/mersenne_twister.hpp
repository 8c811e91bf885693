// Stand-in for the Boost.Random pieces utils/random.h uses.  The reference pins no random stream (test_random.cpp only
// checks moments), so the distributions here use the SAME inversion rules as the oracle and the library
// (u = (r + 0.5) / 2^32; shape-1 Gamma = -log(u) * scale): a seeded random init then gives the same a1, b1 in all three.
#ifndef PCMF_REF_SHIM_BOOST_RANDOM_HPP
#define PCMF_REF_SHIM_BOOST_RANDOM_HPP
#include <cmath>
#include <cstdint>
#include <random>
#include <stdexcept>
namespace boost { namespace random {
typedef std::mt19937 mt19937;
inline double shim_unif(std::mt19937 &g) { return ((double)g() + 0.5) / 4294967296.0; }
template <typename T = double> struct gamma_distribution {
    double shape, scale;
    gamma_distribution(double a, double b) : shape(a), scale(b) {}
    double operator()(std::mt19937 &g) const {
        if (shape != 1.0) throw std::runtime_error("shim: only shape-1 Gamma draws are needed on the VEM path");
        return -std::log(shim_unif(g)) * scale;
    }
};
template <typename T = double> struct uniform_real_distribution {
    double a, b;
    uniform_real_distribution(double a_, double b_) : a(a_), b(b_) {}
    double operator()(std::mt19937 &g) const { return a + (b - a) * shim_unif(g); }
};
template <typename T = int> struct uniform_int_distribution {
    double a, b;
    uniform_int_distribution(double a_, double b_) : a(a_), b(b_) {}
    uint32_t operator()(std::mt19937 &g) const { return (uint32_t)g(); }
};
template <typename T = int> struct poisson_distribution {
    double rate;
    explicit poisson_distribution(double r) : rate(r) {}
    int operator()(std::mt19937 &g) const { std::poisson_distribution<int> d(rate); return d(g); }
};
template <typename RNG, typename Dist> struct variate_generator {
    std::mt19937 &g; Dist d;
    variate_generator(std::mt19937 &g_, const Dist &d_) : g(g_), d(d_) {}
    auto operator()() -> decltype(d(g)) { return d(g); }
};
}}
#endif
