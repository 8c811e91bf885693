// C entry points around the reference's OWN classes (pCMF::variational_em_algo<model> from
// /root/reference/pkg/src/algorithm_variational_EM.h, the four models of *_gap_factor_model.{h,cpp}), compiled where
// they lie against the stand-in headers of oracle/shim/.  TEST INFRASTRUCTURE: oracle/_ref/libpcmf_ref.so is used by
// tests/test_ref_build.py to pin the restated oracle (pcmf_oracle.c) -- and through it the CUDA path -- against the
// reference's real control flow.  The calls below are exactly those of wrapper_gap_factor.h:286-392 /
// wrapper_sparse_gap_factor.h:318-434 (single-init branch), without the R marshalling.
#include <RcppEigen.h>
#include <cstring>
#include <string>

#include "algorithm_variational_EM.h"
#include "gap_factor_model.h"
#include "sparse_gap_factor_model.h"
#include "zi_gap_factor_model.h"
#include "zi_sparse_gap_factor_model.h"
#include "utils/internal.h"
#include "utils/likelihood.h"
#include "utils/probability.h"
#include "utils/random.h"

using Eigen::MatrixXd;
using Eigen::VectorXd;

namespace {

MatrixXd mat(const double *p, int r, int c) {
    MatrixXd m(r, c);
    std::memcpy(m.data(), p, sizeof(double) * (size_t)r * c);
    return m;
}

struct IRunner {
    virtual ~IRunner() {}
    virtual void init_sparse(double sel_bound, const MatrixXd &prob_S, const VectorXd &prior_S) = 0;
    virtual void init_sparse_random(double sel_bound, myRandom::RNGType &rng) = 0;
    virtual void init_zi() = 0;
    virtual void init_zi_given(const MatrixXd &prob_D, const VectorXd &prior_D) = 0;
    virtual void init_variational(const MatrixXd &a1, const MatrixXd &a2, const MatrixXd &b1, const MatrixXd &b2) = 0;
    virtual void init_random(myRandom::RNGType &rng) = 0;
    virtual void init_hyper(const MatrixXd &al1, const MatrixXd &al2, const MatrixXd &be1, const MatrixXd &be2) = 0;
    virtual void init_factor(const MatrixXd &U, const MatrixXd &V) = 0;
    virtual void perturb(myRandom::RNGType &rng, double noise_level) = 0;
    virtual void optimize() = 0;
    virtual void order_factor() = 0;
    virtual void output(Rcpp::List &l) = 0;
};
template <typename model>
struct Runner : IRunner {
    pCMF::variational_em_algo<model> algo;
    Runner(int n, int p, int K, const MatrixXd &X, bool monitor, int iter_max, int iter_min, double eps, int add_iter, int conv_mode)
        : algo(n, p, K, X, false, monitor, iter_max, iter_min, eps, add_iter, conv_mode) {}
    void init_sparse(double sb, const MatrixXd &ps, const VectorXd &pr) override { algo.init_sparse_param_gap(sb, ps, pr); }
    void init_sparse_random(double sb, myRandom::RNGType &rng) override { algo.init_sparse_param_gap(sb, rng); }
    void init_zi() override { algo.init_zi_param_gap(); }
    void init_zi_given(const MatrixXd &pd, const VectorXd &pr) override { algo.init_zi_param_gap(pd, pr); }
    void init_variational(const MatrixXd &a1, const MatrixXd &a2, const MatrixXd &b1, const MatrixXd &b2) override {
        algo.init_variational_param_gap(a1, a2, b1, b2);
    }
    void init_random(myRandom::RNGType &rng) override { algo.init_random(rng); }
    void init_hyper(const MatrixXd &al1, const MatrixXd &al2, const MatrixXd &be1, const MatrixXd &be2) override {
        algo.init_hyper_param_gap(al1, al2, be1, be2);                                        // wrapper_gap_factor.h:362-363
    }
    void init_factor(const MatrixXd &U, const MatrixXd &V) override { algo.init(U, V); }     // wrapper_gap_factor.h:353-354
    void perturb(myRandom::RNGType &rng, double noise_level) override { algo.perturb_param(rng, noise_level); }   // wrapper_gap_factor.h:314
    void optimize() override { algo.optimize(); }
    void order_factor() override { algo.order_factor(); }
    void output(Rcpp::List &l) override { algo.get_output(l); }
};

struct Handle {
    IRunner *r = nullptr;
    myRandom::RNGType rng;
    Rcpp::List out;
    std::string err;
    int sparse = 0, zi = 0;
};

const Rcpp::Any *lookup(const Rcpp::List &l, const char *group, const char *name) {
    const Rcpp::List *cur = &l;
    if (group && *group) {
        const Rcpp::Any *g = l.find(group);
        if (!g || g->kind != Rcpp::Any::LIST) return nullptr;
        cur = g->l.get();
    }
    return cur->find(name);
}

}  // namespace

extern "C" {

void *ref_create(int kind, int n, int p, int K, const double *X, int monitor, int iter_max, int iter_min, double eps,
                 int additional_iter, int conv_mode) {
    Handle *h = new Handle();
    const MatrixXd Xm = mat(X, n, p);
    h->zi = (kind == 1 || kind == 3); h->sparse = (kind == 2 || kind == 3);
    switch (kind) {
        case 0: h->r = new Runner<pCMF::gap_factor_model>(n, p, K, Xm, monitor, iter_max, iter_min, eps, additional_iter, conv_mode); break;
        case 1: h->r = new Runner<pCMF::zi_gap_factor_model>(n, p, K, Xm, monitor, iter_max, iter_min, eps, additional_iter, conv_mode); break;
        case 2: h->r = new Runner<pCMF::sparse_gap_factor_model>(n, p, K, Xm, monitor, iter_max, iter_min, eps, additional_iter, conv_mode); break;
        default: h->r = new Runner<pCMF::zi_sparse_gap_factor_model>(n, p, K, Xm, monitor, iter_max, iter_min, eps, additional_iter, conv_mode); break;
    }
    h->rng = myRandom::rngInit(5489u);
    return h;
}
void ref_destroy(void *hv) { Handle *h = (Handle *)hv; if (h) { delete h->r; delete h; } }
void ref_seed(void *hv, uint32_t seed) { ((Handle *)hv)->rng = myRandom::rngInit(seed); }     // wrapper_gap_factor.h:277-284
const char *ref_last_error(void *hv) { return ((Handle *)hv)->err.c_str(); }

#define GUARD(body) Handle *h = (Handle *)hv; try { body; return 0; } catch (const std::exception &e) { h->err = e.what(); return 1; }
int ref_init_sparse(void *hv, double sel_bound, const double *prob_S, const double *prior_S, int p, int K) {
    GUARD(h->r->init_sparse(sel_bound, mat(prob_S, p, K), mat(prior_S, p, 1)))
}
int ref_init_sparse_random(void *hv, double sel_bound) { GUARD(h->r->init_sparse_random(sel_bound, h->rng)) }
int ref_init_zi(void *hv) { GUARD(h->r->init_zi()) }
int ref_init_zi_given(void *hv, const double *prob_D, const double *prior_D, int n, int p) {
    GUARD(h->r->init_zi_given(mat(prob_D, n, p), mat(prior_D, p, 1)))
}
int ref_init_variational(void *hv, const double *a1, const double *a2, const double *b1, const double *b2, int n, int p, int K) {
    GUARD(h->r->init_variational(mat(a1, n, K), mat(a2, n, K), mat(b1, p, K), mat(b2, p, K)))
}
int ref_init_random(void *hv) { GUARD(h->r->init_random(h->rng)) }
int ref_init_hyper(void *hv, const double *al1, const double *al2, const double *be1, const double *be2, int n, int p, int K) {
    GUARD(h->r->init_hyper(mat(al1, n, K), mat(al2, n, K), mat(be1, p, K), mat(be2, p, K)))
}
int ref_init_factor(void *hv, const double *U, const double *V, int n, int p, int K) { GUARD(h->r->init_factor(mat(U, n, K), mat(V, p, K))) }
int ref_perturb_param(void *hv, double noise_level) { GUARD(h->r->perturb(h->rng, noise_level)) }
int ref_optimize(void *hv) { GUARD(h->r->optimize()) }
int ref_order_factor(void *hv) { GUARD(h->r->order_factor()) }
int ref_output(void *hv) { GUARD(h->out = Rcpp::List(); h->r->output(h->out)) }

// copies entry `name` of sub-list `group` ("" = top level) of the last ref_output() into `out` (capacity `cap` doubles);
// returns the number of values (matrices column-major; rows/cols reported), or -1 when absent / too large
long ref_get(void *hv, const char *group, const char *name, double *out, long cap, int *rows, int *cols) {
    Handle *h = (Handle *)hv;
    const Rcpp::Any *a = lookup(h->out, group, name);
    if (!a) return -1;
    long len = 0; int r = 1, c = 1;
    switch (a->kind) {
        case Rcpp::Any::MATD: r = a->md.rows(); c = a->md.cols(); len = (long)r * c; if (len > cap) return -1;
            for (long e = 0; e < len; ++e) out[e] = a->md[(int)e]; break;
        case Rcpp::Any::MATI: r = a->mi.rows(); c = a->mi.cols(); len = (long)r * c; if (len > cap) return -1;
            for (long e = 0; e < len; ++e) out[e] = (double)a->mi[(int)e]; break;
        case Rcpp::Any::DBL: len = 1; out[0] = a->d; break;
        case Rcpp::Any::INT: len = 1; out[0] = (double)a->i; break;
        case Rcpp::Any::BOOL: len = 1; out[0] = a->b ? 1.0 : 0.0; break;
        default: return -1;
    }
    if (rows) *rows = r;
    if (cols) *cols = c;
    return len;
}

// the reference's primitives, for the known-answer tests (test_internal.cpp, test_probability.cpp, test_likelihood.cpp)
double ref_digammaInv(double y, int nb) { return internal::digammaInv(y, nb); }
double ref_logit(double x) { return internal::logit(x); }
double ref_expit(double x) { return internal::expit(x); }
double ref_custom_log(double x) { return internal::custom_log(x); }
double ref_custom_exp(double x) { return internal::custom_exp(x); }
double ref_digamma(double x) { return boost::math::digamma(x); }
double ref_trigamma(double x) { return boost::math::trigamma(x); }

}  // extern "C"
