/*
 * pcmf_oracle.c -- CPU ORACLE for the pCMF variational-EM hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (pcmf_b200/) never links, imports or calls anything in oracle/.
 *
 * What it is: a dense FP64 restatement, loop for loop, of the reference's C++ path
 * (gdurif/pCMF v1.2.1, pkg/src).  Every function cites the reference file:line it follows.
 * It keeps the reference's data layout (column-major dense n x p X, UVt, prob_D and the
 * n x p scratch m_exp_ElogU_ElogV_k), its loop structure (`omp parallel for` over the
 * columns j with per-thread private accumulators reduced at the end, K `exp` per (i,j))
 * and its quirks.  Compiled `gcc -O2 -fopenmp` like an R package.
 *
 * PARITY STATUS: PINNED against the reference itself.  R, Rcpp, Eigen and Boost are absent
 * from this image, so the reference's package cannot be built as such; instead its OWN model
 * and algorithm sources (pkg/src/{model,gap_factor_model,zi_gap_factor_model,
 * sparse_gap_factor_model,zi_sparse_gap_factor_model}.cpp, algorithm.h, algorithm_variational_EM.h and the headers under utils) are
 * compiled where they lie against small stand-in headers (oracle/shim/: an eager dense
 * matrix class covering the Eigen expression forms those files use, a list-shaped Rcpp
 * stub, digamma/trigamma and mt19937 distributions for Boost) into
 * oracle/_ref/libpcmf_ref.so (`make -C oracle ref`).  tests/test_ref_build.py runs this
 * oracle and that build side by side -- whole driver, all four models, monitors, stopping
 * rule, factor ordering, seeded random init -- and they agree to 1e-10 relative (observed
 * ~1e-14); tests/golden/ref_*.npz are outputs of that build, committed so the check travels
 * (tests/test_golden_vectors.py, tests/test_gpu_golden.py).  Caveat, stated where it
 * applies: the linear algebra underneath the reference's formulas is the stand-in's, not
 * Eigen's (summation order differs; vectors have no static orientation).
 * Also pinned (tests/test_oracle_primitives.py): the reference's known-answer tests for the
 * primitives -- digammaInv 7.8834286312 / 22026.9657929151 (test_internal.cpp:35-41),
 * logit/expit clamps (:58-80), custom_log (:51-56), variance (:43-49), Bregman-Poisson
 * 23.57360306 (test_probability.cpp:119-135), Egamma/ElogGamma/entropy identities (:38-108),
 * likelihood identities (test_likelihood.cpp:39-172) -- plus scipy cross-checks of
 * digamma/trigamma/lgamma.
 *
 * Third-party arithmetic restated here (absent from /root/reference):
 *   boost::math::digamma / trigamma (package BH, version unpinned by DESCRIPTION:22):
 *     restated as upward recurrence to x >= 10 + the published asymptotic (Bernoulli)
 *     series -- the same series Boost uses for x >= 10; for x < 10 Boost recurs into [1,2]
 *     and uses a minimax rational, the two agree to a few ulp (checked against scipy).
 *   boost::random mt19937 + gamma_distribution (BH, unpinned, no test pins the stream):
 *     NOT reproduced; orc_init_random uses MT19937 + inversion for the shape-1 Gamma
 *     (an exponential), the same rule as oracle/shim/boost/random and the library, so seeded
 *     random inits coincide across the three.  Most parity tests inject a1,a2,b1,b2 explicitly.
 *
 * Deliberate deviation: likelihood::bernoulli_loglike_vec (likelihood.h:296-313) reads out
 * of bounds in the reference (prob(i,j) on a row vector) and, in the sparse call
 * (sparse_gap_factor_model.cpp:172), uses the prior of GENE k for FACTOR k; we implement the intended
 * term sum_ij P_ij log(pi_j) + (1-P_ij) log(1-pi_j) gated on 0<pi_j<1.  The tests
 * compare the sparse ELBO with the reference build up to exactly that term (tests/refcmp.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_GAP 0
#define ORC_ZI 1
#define ORC_SPARSE 2
#define ORC_ZI_SPARSE 3

/* ------------------------------------------------------------------------------------ */
/* special functions (restating boost::math::digamma/trigamma, see header)               */
/* ------------------------------------------------------------------------------------ */

double orc_digamma(double x) {
    /* call sites: probability.h:107, internal.h:64,69, macros.h:46 */
    if (x <= 0.0) {
        /* reflection psi(1-x) - pi/tan(pi x); not reached from the VEM path (shapes > 0) */
        if (x == floor(x)) return NAN;
        return orc_digamma(1.0 - x) - M_PI / tan(M_PI * x);
    }
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double f = 1.0 / (x * x);
    /* sum_k B_2k / (2k x^2k): 1/12, -1/120, 1/252, -1/240, 1/132, -691/32760, 1/12, -3617/8160 */
    double s = f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0
             - f * (1.0 / 132.0 - f * (691.0 / 32760.0 - f * (1.0 / 12.0 - f * (3617.0 / 8160.0))))))));
    return r + log(x) - 0.5 / x - s;
}

double orc_trigamma(double x) {
    /* call site: internal.h:69 */
    if (x <= 0.0) {
        if (x == floor(x)) return NAN;
        double s = M_PI / sin(M_PI * x);
        return -orc_trigamma(1.0 - x) + s * s;
    }
    double r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    double f = 1.0 / (x * x);
    /* 1/x + 1/(2x^2) + sum_k B_2k / x^(2k+1): 1/6, -1/30, 1/42, -1/30, 5/66, -691/2730, 7/6 */
    double s = (1.0 / x) * f * (1.0 / 6.0 - f * (1.0 / 30.0 - f * (1.0 / 42.0 - f * (1.0 / 30.0
             - f * (5.0 / 66.0 - f * (691.0 / 2730.0 - f * (7.0 / 6.0)))))));
    return r + 1.0 / x + 0.5 * f + s;
}

/* internal.h:56-74 */
double orc_digammaInv(double y, int nbIter) {
    double x0, x = 0;
    if (y >= -2.22) x0 = exp(y) + 0.5;
    else x0 = -1.0 / (y - orc_digamma(1.0));
    for (int i = 0; i < nbIter; i++) {
        x = x0 - (orc_digamma(x0) - y) / orc_trigamma(x0);
        x0 = x;
    }
    return x;
}

/* internal.h:83-90 */
double orc_variance(const double *s, int n) {
    double m = 0; for (int i = 0; i < n; i++) m += s[i]; m /= n;
    double v = 0; for (int i = 0; i < n; i++) v += (s[i] - m) * (s[i] - m);
    return v / (double)(n - 1);
}
/* internal.h:104-110 */
double orc_custom_log(double x) { return x > 0 ? log(x) : 0.0; }
/* internal.h:121-127 */
double orc_custom_exp(double x) { return x > -100 ? exp(x) : 3e-44; }
/* internal.h:137-141 */
double orc_logit(double x) {
    if (x >= 1 - 1e-12) return 30;
    if (x <= 1e-12) return -30;
    return log(x / (1 - x));
}
/* internal.h:151-155 */
double orc_expit(double x) {
    if (x >= 30) return 1;
    if (x <= -30) return 0;
    return 1.0 / (1 + exp(-x));
}
/* probability.h:76-78 / :106-108 */
double orc_Egamma(double p1, double p2) { return p1 / p2; }
double orc_ElogGamma(double p1, double p2) { return orc_digamma(p1) - log(p2); }
/* probability.h:123-126 (scalar gamma_entropy) */
double orc_gamma_entropy1(double p1, double p2) {
    return (1 - p1) * orc_digamma(p1) + p1 + lgamma(p1) - log(p2);
}
/* probability.h:174-176 */
double orc_bregman_poisson(double x, double y) { return x * log((x > 0 ? x : 1) / y) - x + y; }
/* matrix.h:65-82 specialised to bregman_poisson (test_probability.cpp:119-135) */
double orc_bregman_poisson_sum(const double *X, const double *Y, int len) {
    double r = 0; for (int i = 0; i < len; i++) r += orc_bregman_poisson(X[i], Y[i]); return r;
}

/* probability.h:157-160 */
static double gamma_entropy_sum(const double *p1, const double *p2, long len) {
    double r = 0;
    for (long i = 0; i < len; i++) r += orc_gamma_entropy1(p1[i], p2[i]);
    return r;
}
double orc_gamma_entropy(const double *p1, const double *p2, long len) { return gamma_entropy_sum(p1, p2, len); }

/* likelihood.h:108-113 (with logX) */
double orc_gamma_loglike_logX(const double *X, const double *logX, const double *p1, const double *p2, long len) {
    double r = 0;
    for (long i = 0; i < len; i++)
        r += p1[i] * log(p2[i]) + (p1[i] - 1) * logX[i] - p2[i] * X[i] - lgamma(p1[i]);
    return r;
}
/* likelihood.h:85-89 (log of X) */
double orc_gamma_loglike(const double *X, const double *p1, const double *p2, long len) {
    double r = 0;
    for (long i = 0; i < len; i++)
        r += p1[i] * log(p2[i]) + (p1[i] - 1) * log(X[i]) - p2[i] * X[i] - lgamma(p1[i]);
    return r;
}
/* likelihood.h:131-135 scalar */
double orc_poisson_loglike1(double x, double rate) { return x * log(rate > 0 ? rate : 1) - rate - lgamma(x + 1); }
/* likelihood.h:149-155 */
double orc_poisson_loglike(const double *X, const double *rate, long len) {
    double r = 0;
#pragma omp parallel for reduction(+:r)
    for (long i = 0; i < len; i++) r += X[i] * orc_custom_log(rate[i]) - rate[i] - lgamma(X[i] + 1);
    return r;
}
/* likelihood.h:172-180 (rate: one value per column) */
double orc_poisson_loglike_vec(const double *X, const double *rate, int n, int p) {
    double r = 0;
#pragma omp parallel for reduction(+:r)
    for (int j = 0; j < p; j++)
        for (int i = 0; i < n; i++)
            r += X[i + (long)j * n] * orc_custom_log(rate[j]) - rate[j] - lgamma(X[i + (long)j * n] + 1);
    return r;
}
/* likelihood.h:197-212 */
double orc_zi_poisson_loglike(const double *X, const double *rate, const double *prob, int n, int p) {
    double r = 0;
#pragma omp parallel for reduction(+:r)
    for (int j = 0; j < p; j++)
        for (int i = 0; i < n; i++) {
            long e = i + (long)j * n;
            r += X[e] > 0 ? log(prob[e]) + orc_poisson_loglike1(X[e], rate[e])
                          : log(1 - prob[e] + prob[e] * exp(-rate[e]));
        }
    return r;
}
/* likelihood.h:261-279 */
double orc_bernoulli_loglike(const double *X, const double *prob, long len) {
    double r = 0;
#pragma omp parallel for reduction(+:r)
    for (long e = 0; e < len; e++)
        if (prob[e] > 0 && prob[e] < 1) r += X[e] * log(prob[e]) + (1 - X[e]) * log(1 - prob[e]);
    return r;
}
/* likelihood.h:296-313, INTENDED semantics (see header): X is rows x cols col-major,
 * prob has one entry per column; gate on 0 < prob_j < 1. */
double orc_bernoulli_loglike_colvec(const double *X, const double *prob, int rows, int cols) {
    double r = 0;
#pragma omp parallel for reduction(+:r)
    for (int j = 0; j < cols; j++)
        if (prob[j] > 0 && prob[j] < 1)
            for (int i = 0; i < rows; i++)
                r += X[i + (long)j * rows] * log(prob[j]) + (1 - X[i + (long)j * rows]) * log(1 - prob[j]);
    return r;
}
/* same, prob has one entry per ROW (the sparse call sparse_gap_factor_model.cpp:172 passes
 * prob_S (p x K) with prior_S (p): intended index is the gene j = row). */
double orc_bernoulli_loglike_rowvec(const double *X, const double *prob, int rows, int cols) {
    double r = 0;
    for (int j = 0; j < cols; j++)
        for (int i = 0; i < rows; i++)
            if (prob[i] > 0 && prob[i] < 1)
                r += X[i + (long)j * rows] * log(prob[i]) + (1 - X[i + (long)j * rows]) * log(1 - prob[i]);
    return r;
}

/* ------------------------------------------------------------------------------------ */
/* MT19937 (public algorithm, Matsumoto & Nishimura) -- only for seed-driven random init */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } orc_rng;
void orc_rng_seed(orc_rng *r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}
uint32_t orc_rng_next(orc_rng *r) {
    if (r->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (r->mt[i] & 0x80000000u) | (r->mt[(i + 1) % 624] & 0x7fffffffu);
            r->mt[i] = r->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
static double rng_unif(orc_rng *r) { return ((double)orc_rng_next(r) + 0.5) / 4294967296.0; }

/* ------------------------------------------------------------------------------------ */
/* model state (gap_factor_model.h:72-98, model.h:144-151,284-288, zi...h:83-85,          */
/* sparse...h:81-85).  All matrices column-major like Eigen.                              */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int kind, n, p, K;
    int zi, sparse;
    double *X, *UVt, *T;                 /* n x p */
    double *EU, *ElogU, *a1, *a2, *a1o, *a2o, *alpha1, *alpha2, *alpha1o, *alpha2o, *EZ_j;   /* n x K */
    double *EV, *ElogV, *b1, *b2, *b1o, *b2o, *beta1, *beta2, *beta1o, *beta2o, *EZ_i;       /* p x K */
    double *prob_D, *prior_D, *freq_D;   /* n x p, p, p */
    double *prob_S, *prior_S; int *S; double sel_bound;   /* p x K, p, p x K */
    int *factor_order;
    orc_rng rng;
    int nz_only;   /* sparse-aware variant of the same loops: entries with X = 0, whose multinomial weights are multiplied by 0, skip the
                    * K exponentials (orc_set_sparse_aware; bench.py's second CPU figure).  0 = the reference's loops as they are. */
} orc_model;

static double *dalloc(long len, double v) {
    double *a = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
    for (long i = 0; i < len; i++) a[i] = v;
    return a;
}

/* constructors: model.cpp:76-91, gap_factor_model.cpp:52-87, zi...cpp:53-59, sparse...cpp:55-65 */
orc_model *orc_create(int kind, int n, int p, int K, const double *X) {
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->kind = kind; m->n = n; m->p = p; m->K = K;
    m->zi = (kind == ORC_ZI || kind == ORC_ZI_SPARSE);
    m->sparse = (kind == ORC_SPARSE || kind == ORC_ZI_SPARSE);
    long np = (long)n * p, nK = (long)n * K, pK = (long)p * K;
    m->X = dalloc(np, 0); memcpy(m->X, X, sizeof(double) * np);
    /* constructors' Ones: UVt (model.cpp:86-91), EU, EV, ElogU, ElogV (model.cpp:220-223) -- what init_hyper_param alone
     * (wrapper_gap_factor.h:362-363) leaves in place */
    m->UVt = dalloc(np, 1); m->T = dalloc(np, 0);
    m->EU = dalloc(nK, 1); m->ElogU = dalloc(nK, 1); m->EV = dalloc(pK, 1); m->ElogV = dalloc(pK, 1);
    m->a1 = dalloc(nK, 1); m->a2 = dalloc(nK, 1); m->a1o = dalloc(nK, 1); m->a2o = dalloc(nK, 1);
    m->b1 = dalloc(pK, 1); m->b2 = dalloc(pK, 1); m->b1o = dalloc(pK, 1); m->b2o = dalloc(pK, 1);
    m->alpha1 = dalloc(nK, 1); m->alpha2 = dalloc(nK, 1); m->alpha1o = dalloc(nK, 1); m->alpha2o = dalloc(nK, 1);
    m->beta1 = dalloc(pK, 1); m->beta2 = dalloc(pK, 1); m->beta1o = dalloc(pK, 1); m->beta2o = dalloc(pK, 1);
    m->EZ_i = dalloc(pK, 0); m->EZ_j = dalloc(nK, 0);
    if (m->zi) { m->prob_D = dalloc(np, 1); m->prior_D = dalloc(p, 1); m->freq_D = dalloc(p, 1); }
    if (m->sparse) {
        m->prob_S = dalloc(pK, 0); m->prior_S = dalloc(p, 0);
        m->S = (int *)malloc(sizeof(int) * pK); for (long i = 0; i < pK; i++) m->S[i] = 1;
        m->sel_bound = 0.5;
    }
    m->factor_order = (int *)malloc(sizeof(int) * K);
    for (int k = 0; k < K; k++) m->factor_order[k] = k;
    orc_rng_seed(&m->rng, 5489u);
    return m;
}
void orc_destroy(orc_model *m) {
    if (!m) return;
    double **all[] = { &m->X, &m->UVt, &m->T, &m->EU, &m->ElogU, &m->EV, &m->ElogV, &m->a1, &m->a2, &m->a1o, &m->a2o,
        &m->b1, &m->b2, &m->b1o, &m->b2o, &m->alpha1, &m->alpha2, &m->alpha1o, &m->alpha2o, &m->beta1, &m->beta2,
        &m->beta1o, &m->beta2o, &m->EZ_i, &m->EZ_j, &m->prob_D, &m->prior_D, &m->freq_D, &m->prob_S, &m->prior_S };
    for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); i++) free(*all[i]);
    free(m->S); free(m->factor_order); free(m);
}
void orc_seed(orc_model *m, uint32_t seed) { orc_rng_seed(&m->rng, seed); }
void orc_set_sparse_aware(orc_model *m, int flag) { m->nz_only = flag != 0; }
/* `my_model = tmp_model` of the multi-init loop (wrapper_gap_factor.h:339-345): deep copy of every member, including
 * the `old` iterates.  Both models must have been created with the same kind and sizes. */
void orc_copy_state(orc_model *dst, const orc_model *src) {
    long np = (long)src->n * src->p, nK = (long)src->n * src->K, pK = (long)src->p * src->K;
#define CPY(f, l) if (src->f) memcpy(dst->f, src->f, sizeof(*src->f) * (size_t)(l));
    CPY(UVt, np) CPY(T, np) CPY(EU, nK) CPY(ElogU, nK) CPY(EV, pK) CPY(ElogV, pK)
    CPY(a1, nK) CPY(a2, nK) CPY(a1o, nK) CPY(a2o, nK) CPY(b1, pK) CPY(b2, pK) CPY(b1o, pK) CPY(b2o, pK)
    CPY(alpha1, nK) CPY(alpha2, nK) CPY(alpha1o, nK) CPY(alpha2o, nK) CPY(beta1, pK) CPY(beta2, pK) CPY(beta1o, pK) CPY(beta2o, pK)
    CPY(EZ_i, pK) CPY(EZ_j, nK) CPY(prob_D, np) CPY(prior_D, src->p) CPY(freq_D, src->p) CPY(prob_S, pK) CPY(prior_S, src->p)
    CPY(S, pK) CPY(factor_order, src->K)
#undef CPY
    dst->sel_bound = src->sel_bound;
}

/* field access for the test harness */
double *orc_field(orc_model *m, const char *name, long *len) {
    long np = (long)m->n * m->p, nK = (long)m->n * m->K, pK = (long)m->p * m->K;
#define F(nm, ptr, l) if (!strcmp(name, nm)) { if (len) *len = (l); return (ptr); }
    F("X", m->X, np) F("UVt", m->UVt, np) F("T", m->T, np)
    F("EU", m->EU, nK) F("ElogU", m->ElogU, nK) F("EV", m->EV, pK) F("ElogV", m->ElogV, pK)
    F("a1", m->a1, nK) F("a2", m->a2, nK) F("b1", m->b1, pK) F("b2", m->b2, pK)
    F("a1old", m->a1o, nK) F("a2old", m->a2o, nK) F("b1old", m->b1o, pK) F("b2old", m->b2o, pK)
    F("alpha1", m->alpha1, nK) F("alpha2", m->alpha2, nK) F("beta1", m->beta1, pK) F("beta2", m->beta2, pK)
    F("EZ_i", m->EZ_i, pK) F("EZ_j", m->EZ_j, nK)
    F("prob_D", m->prob_D, np) F("prior_D", m->prior_D, m->p) F("freq_D", m->freq_D, m->p)
    F("prob_S", m->prob_S, pK) F("prior_S", m->prior_S, m->p)
#undef F
    if (len) *len = 0;
    return NULL;
}
int *orc_field_S(orc_model *m) { return m->S; }
int *orc_field_factor_order(orc_model *m) { return m->factor_order; }

/* ------------------------------------------------------------------------------------ */
/* sufficient statistics and Poisson rate                                                */
/* ------------------------------------------------------------------------------------ */
/* gap_factor_model.cpp:408-411 */
static void U_stats(orc_model *m) {
    long nK = (long)m->n * m->K;
#pragma omp parallel for
    for (long e = 0; e < nK; e++) { m->EU[e] = m->a1[e] / m->a2[e]; m->ElogU[e] = orc_digamma(m->a1[e]) - log(m->a2[e]); }
}
/* gap_factor_model.cpp:414-417 */
static void V_stats(orc_model *m) {
    long pK = (long)m->p * m->K;
#pragma omp parallel for
    for (long e = 0; e < pK; e++) { m->EV[e] = m->b1[e] / m->b2[e]; m->ElogV[e] = orc_digamma(m->b1[e]) - log(m->b2[e]); }
}
/* UVt(i,j) = sum_k EU(i,k) * W(j,k), W = EV (gap...cpp:453-455) or prob_S o EV (sparse...cpp:322-324);
 * only the first kk columns after permuting by `perm` are used (partial_deviance). */
static void uvt_into(const orc_model *m, double *out, const int *perm, int kk) {
    int n = m->n, p = m->p;
#pragma omp parallel for
    for (int j = 0; j < p; j++) {
        double *col = out + (long)j * n;
        for (int i = 0; i < n; i++) col[i] = 0;
        for (int l = 0; l < kk; l++) {
            int k = perm ? perm[l] : l;
            double w = m->EV[j + (long)k * p];
            if (m->sparse) w *= m->prob_S[j + (long)k * p];
            const double *eu = m->EU + (long)k * n;
            for (int i = 0; i < n; i++) col[i] += eu[i] * w;
        }
    }
}
static void update_poisson_param(orc_model *m) { uvt_into(m, m->UVt, NULL, m->K); }

/* ------------------------------------------------------------------------------------ */
/* hyper-parameter updates                                                               */
/* ------------------------------------------------------------------------------------ */
/* gap_factor_model.cpp:541-548 (digammaInv with 6 Newton steps, macros.h:47) */
static void update_prior_gamma_param(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
    for (int k = 0; k < K; k++) {
        double mElogU = 0, mEU = 0;
        for (int i = 0; i < n; i++) { mElogU += m->ElogU[i + (long)k * n]; mEU += m->EU[i + (long)k * n]; }
        mElogU /= n; mEU /= n;
#pragma omp parallel for
        for (int i = 0; i < n; i++) {
            long e = i + (long)k * n;
            m->alpha1[e] = orc_digammaInv(log(m->alpha2[e]) + mElogU, 6);
            m->alpha2[e] = m->alpha1[e] / mEU;
        }
        double mElogV = 0, mEV = 0;
        for (int j = 0; j < p; j++) { mElogV += m->ElogV[j + (long)k * p]; mEV += m->EV[j + (long)k * p]; }
        mElogV /= p; mEV /= p;
#pragma omp parallel for
        for (int j = 0; j < p; j++) {
            long e = j + (long)k * p;
            m->beta1[e] = orc_digammaInv(log(m->beta2[e]) + mElogV, 6);
            m->beta2[e] = m->beta1[e] / mEV;
        }
    }
}
/* zi_gap_factor_model.cpp:370-372 */
static void update_prior_zi_param(orc_model *m) {
    int n = m->n, p = m->p;
#pragma omp parallel for
    for (int j = 0; j < p; j++) {
        double s = 0; for (int i = 0; i < n; i++) s += m->prob_D[i + (long)j * n];
        m->prior_D[j] = s / n;
    }
}
/* sparse_gap_factor_model.cpp:474-476 */
static void update_prior_sparse_param(orc_model *m) {
    int p = m->p, K = m->K;
    for (int j = 0; j < p; j++) {
        double s = 0; for (int k = 0; k < K; k++) s += m->prob_S[j + (long)k * p];
        m->prior_S[j] = s / K;
    }
}
/* gap...cpp:233-236, zi...cpp:146-151, sparse...cpp:136-141, zi_sparse...cpp:154-161 */
static void update_hyper_param(orc_model *m) {
    update_prior_gamma_param(m);
    if (m->sparse) update_prior_sparse_param(m);
    if (m->zi) update_prior_zi_param(m);
}

/* ------------------------------------------------------------------------------------ */
/* initialisation (wrapper_sparse_gap_factor.h:389-417 gives the order)                  */
/* ------------------------------------------------------------------------------------ */
/* sparse_gap_factor_model.cpp:96-104 */
void orc_init_sparse_param(orc_model *m, double sel_bound, const double *prob_S, const double *prior_S) {
    long pK = (long)m->p * m->K;
    m->sel_bound = sel_bound;
    memcpy(m->prob_S, prob_S, sizeof(double) * pK);
    memcpy(m->prior_S, prior_S, sizeof(double) * m->p);
    for (long e = 0; e < pK; e++) m->S[e] = m->prob_S[e] >= sel_bound;
}
/* sparse_gap_factor_model.cpp:71-88 (uniform prob_S; our own stream, see header) */
void orc_init_sparse_param_random(orc_model *m, double sel_bound) {
    int p = m->p, K = m->K;
    m->sel_bound = sel_bound;
    for (int j = 0; j < p; j++) for (int k = 0; k < K; k++) m->prob_S[j + (long)k * p] = rng_unif(&m->rng);
    update_prior_sparse_param(m);
    for (long e = 0; e < (long)p * K; e++) m->S[e] = m->prob_S[e] >= sel_bound;
}
/* zi_gap_factor_model.cpp:98-105 */
void orc_init_zi_param(orc_model *m) {
    int n = m->n, p = m->p;
    for (int j = 0; j < p; j++) {
        double s = 0;
        for (int i = 0; i < n; i++) { double d = m->X[i + (long)j * n] > 0 ? 1.0 : 0.0; m->prob_D[i + (long)j * n] = d; s += d; }
        m->freq_D[j] = s / n; m->prior_D[j] = m->freq_D[j];
    }
}
/* zi_gap_factor_model.cpp:108-114 */
void orc_init_zi_param_given(orc_model *m, const double *prob_D, const double *prior_D) {
    memcpy(m->prob_D, prob_D, sizeof(double) * (long)m->n * m->p);
    memcpy(m->prior_D, prior_D, sizeof(double) * m->p);
}
static void post_init(orc_model *m, int hyper) {
    U_stats(m); V_stats(m); update_poisson_param(m);
    if (hyper) update_hyper_param(m);
}
/* gap_factor_model.cpp:113-123 */
void orc_init_variational_param(orc_model *m, const double *a1, const double *a2, const double *b1, const double *b2) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    memcpy(m->a1, a1, 8 * nK); memcpy(m->a2, a2, 8 * nK); memcpy(m->b1, b1, 8 * pK); memcpy(m->b2, b2, 8 * pK);
    post_init(m, 1);
}
/* gap_factor_model.cpp:126-132 */
void orc_init_hyper_param(orc_model *m, const double *al1, const double *al2, const double *be1, const double *be2) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    memcpy(m->alpha1, al1, 8 * nK); memcpy(m->alpha2, al2, 8 * nK); memcpy(m->beta1, be1, 8 * pK); memcpy(m->beta2, be2, 8 * pK);
}
/* gap_factor_model.cpp:95-110 */
void orc_init_all_param(orc_model *m, const double *al1, const double *al2, const double *be1, const double *be2,
                        const double *a1, const double *a2, const double *b1, const double *b2) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    orc_init_hyper_param(m, al1, al2, be1, be2);
    memcpy(m->a1, a1, 8 * nK); memcpy(m->a2, a2, 8 * nK); memcpy(m->b1, b1, 8 * pK); memcpy(m->b2, b2, 8 * pK);
    post_init(m, 0);
}
/* gap_factor_model.cpp:135-144 */
void orc_init_from_factor(orc_model *m, const double *U, const double *V) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    memcpy(m->a1, U, 8 * nK); memcpy(m->b1, V, 8 * pK);
    for (long e = 0; e < nK; e++) m->a2[e] = 1; for (long e = 0; e < pK; e++) m->b2[e] = 1;
    post_init(m, 1);
}
/* gap_factor_model.cpp:147-174: a1(i,.) ~ Gamma(1, rate sqrt(K/rowmean_i)), b1(j,.) ~ Gamma(1, rate sqrt(K/colmean_j)),
 * row by row, K draws each, U first then V.  Shape-1 Gamma == exponential -> inversion. */
void orc_init_random(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
    long nK = (long)n * K, pK = (long)p * K;
    for (long e = 0; e < nK; e++) { m->a1[e] = 1; m->a2[e] = 1; }
    for (long e = 0; e < pK; e++) { m->b1[e] = 1; m->b2[e] = 1; }
    for (int i = 0; i < n; i++) {
        double s = 0; for (int j = 0; j < p; j++) s += m->X[i + (long)j * n];
        double rate = sqrt(K / (s / p));
        for (int k = 0; k < K; k++) m->a1[i + (long)k * n] = -log(rng_unif(&m->rng)) / rate;
    }
    for (int j = 0; j < p; j++) {
        double s = 0; for (int i = 0; i < n; i++) s += m->X[i + (long)j * n];
        double rate = sqrt(K / (s / n));
        for (int k = 0; k < K; k++) m->b1[j + (long)k * p] = -log(rng_unif(&m->rng)) / rate;
    }
    post_init(m, 1);
}
/* gap_factor_model.cpp:177-192 */
void orc_perturb_param(orc_model *m, double noise_level) {
    int n = m->n, p = m->p, K = m->K;
    /* rUnif fills row by row (rowInd outer, colInd inner, random.h:224-237): noise_a1 first, then noise_b1 */
    for (int i = 0; i < n; i++) for (int k = 0; k < K; k++) m->a1[i + (long)k * n] += -noise_level + 2 * noise_level * rng_unif(&m->rng);
    for (int j = 0; j < p; j++) for (int k = 0; k < K; k++) m->b1[j + (long)k * p] += -noise_level + 2 * noise_level * rng_unif(&m->rng);
    for (long e = 0; e < (long)n * K; e++) if (m->a1[e] < 1e-6) m->a1[e] = 1e-6;
    for (long e = 0; e < (long)p * K; e++) if (m->b1[e] < 1e-6) m->b1[e] = 1e-6;
    post_init(m, 1);
}

/* ------------------------------------------------------------------------------------ */
/* E-step                                                                                */
/* ------------------------------------------------------------------------------------ */
/* the per-(i,j) kernel shared by gap...cpp:481-487, sparse...cpp:350-356 etc.:
 * vec_k = cexp(ElogU(i,k)+ElogV(j,k) - max) [* S(j,k)]; returns the sum */
static inline double multinomial_vec(const orc_model *m, int i, int j, double *vec) {
    int n = m->n, p = m->p, K = m->K;
    double mx = -INFINITY;
    for (int k = 0; k < K; k++) { vec[k] = m->ElogU[i + (long)k * n] + m->ElogV[j + (long)k * p]; if (vec[k] > mx) mx = vec[k]; }
    if (mx < 100) mx = 0;
    double s = 0;
    for (int k = 0; k < K; k++) {
        vec[k] = orc_custom_exp(vec[k] - mx);
        if (m->sparse) vec[k] *= (double)m->S[j + (long)k * p];
        s += vec[k];
    }
    return s;
}

/* gap_factor_model.cpp:458-499, zi...cpp:289-330, sparse...cpp:327-368, zi_sparse...cpp:316-357 */
static void update_variational_multinomial_param(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
    long nK = (long)n * K, pK = (long)p * K;
    memset(m->EZ_i, 0, 8 * pK); memset(m->EZ_j, 0, 8 * nK);
#pragma omp parallel
    {
        /* per-thread private accumulators + reduction, as the reference's
         * `omp declare reduction(+: Eigen::MatrixXd ...)` (gap_factor_model.cpp:474-478) */
        double *pEZ_j = dalloc(nK, 0);
        double *vec = (double *)malloc(8 * K);
#pragma omp for
        for (int j = 0; j < p; j++) {
            for (int i = 0; i < n; i++) {
                if (m->nz_only && m->X[i + (long)j * n] == 0) continue;   /* every term below carries the factor X_ij */
                double tv = multinomial_vec(m, i, j, vec);
                m->T[i + (long)j * n] = tv;
                tv = tv > 0 ? tv : 1;
                double w = m->X[i + (long)j * n];
                if (m->zi) w = m->prob_D[i + (long)j * n] * w;
                for (int k = 0; k < K; k++) {
                    double acc;
                    if (m->sparse) acc = (m->zi ? m->prob_D[i + (long)j * n] * m->prob_S[j + (long)k * p] * m->X[i + (long)j * n]
                                                : m->prob_S[j + (long)k * p] * m->X[i + (long)j * n]) * vec[k] / tv;
                    else acc = w * vec[k] / tv;
                    m->EZ_i[j + (long)k * p] += acc;   /* row j of EZ_i is owned by this thread */
                    pEZ_j[i + (long)k * n] += acc;
                }
            }
        }
#pragma omp critical
        for (long e = 0; e < nK; e++) m->EZ_j[e] += pEZ_j[e];
        free(pEZ_j); free(vec);
    }
}

/* gap_factor_model.cpp:502-525, sparse...cpp:371-394 */
static void intermediate_update_variational_multinomial_param(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
#pragma omp parallel
    {
        double *vec = (double *)malloc(8 * K);
#pragma omp for
        for (int j = 0; j < p; j++)
            for (int i = 0; i < n; i++) {
                if (m->nz_only && m->X[i + (long)j * n] == 0) continue;   /* T is only read where X != 0 (orc_elbo) */
                m->T[i + (long)j * n] = multinomial_vec(m, i, j, vec);
            }
        free(vec);
    }
}

/* ------------------------------------------------------------------------------------ */
/* M-step                                                                                */
/* ------------------------------------------------------------------------------------ */
/* gap...cpp:528-538, zi...cpp:333-343, sparse...cpp:397-421, zi_sparse...cpp:360-386 */
static void update_variational_gamma_param(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
    /* factor U (uses the OLD EV / prob_S / prob_D) */
    for (int k = 0; k < K; k++) {
        if (!m->zi) {
            double cs = 0;
            for (int j = 0; j < p; j++) cs += m->sparse ? m->EV[j + (long)k * p] * m->prob_S[j + (long)k * p] : m->EV[j + (long)k * p];
            for (int i = 0; i < n; i++) m->a2[i + (long)k * n] = m->alpha2[i + (long)k * n] + cs;
        } else {
#pragma omp parallel for
            for (int i = 0; i < n; i++) {
                double s = 0;   /* (prob_D * EV[o prob_S])(i,k) */
                for (int j = 0; j < p; j++) {
                    double w = m->EV[j + (long)k * p]; if (m->sparse) w *= m->prob_S[j + (long)k * p];
                    s += m->prob_D[i + (long)j * n] * w;
                }
                m->a2[i + (long)k * n] = m->alpha2[i + (long)k * n] + s;
            }
        }
        for (int i = 0; i < n; i++) m->a1[i + (long)k * n] = m->alpha1[i + (long)k * n] + m->EZ_j[i + (long)k * n];
    }
    U_stats(m);
    /* factor V (uses the NEW EU) */
    for (int k = 0; k < K; k++) {
        for (int j = 0; j < p; j++) m->b1[j + (long)k * p] = m->beta1[j + (long)k * p] + m->EZ_i[j + (long)k * p];
        if (!m->zi) {
            double cs = 0; for (int i = 0; i < n; i++) cs += m->EU[i + (long)k * n];
            for (int j = 0; j < p; j++)
                m->b2[j + (long)k * p] = m->beta2[j + (long)k * p] + (m->sparse ? m->prob_S[j + (long)k * p] * cs : cs);
        } else {
#pragma omp parallel for
            for (int j = 0; j < p; j++) {
                double s = 0;   /* (prob_D^T * EU)(j,k) */
                for (int i = 0; i < n; i++) s += m->prob_D[i + (long)j * n] * m->EU[i + (long)k * n];
                m->b2[j + (long)k * p] = m->beta2[j + (long)k * p] + (m->sparse ? m->prob_S[j + (long)k * p] * s : s);
            }
        }
    }
    V_stats(m);
}

/* sparse_gap_factor_model.cpp:424-471, zi_sparse_gap_factor_model.cpp:413-460 */
static void update_variational_sparse_param(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
#pragma omp parallel
    {
        double *vec = (double *)malloc(8 * K);
        double *tmp = (double *)malloc(8 * K);
#pragma omp for
        for (int j = 0; j < p; j++) {
            if (m->prior_S[j] == 1) { for (int k = 0; k < K; k++) m->prob_S[j + (long)k * p] = 1; }
            else if (m->prior_S[j] == 0) { for (int k = 0; k < K; k++) m->prob_S[j + (long)k * p] = 0; }
            else {
                for (int k = 0; k < K; k++) tmp[k] = 0;
                for (int i = 0; i < n; i++) {
                    double pd = m->zi ? m->prob_D[i + (long)j * n] : 1.0;
                    if (m->nz_only && m->X[i + (long)j * n] == 0) {           /* only the rate term is left: no exponentials */
                        for (int k = 0; k < K; k++) tmp[k] -= pd * m->EU[i + (long)k * n] * m->EV[j + (long)k * p];
                        continue;
                    }
                    double tv = multinomial_vec(m, i, j, vec);
                    tv = tv > 0 ? tv : 1;
                    for (int k = 0; k < K; k++) {
                        double r = vec[k] / tv;
                        if (m->zi) {
                            tmp[k] -= pd * m->EU[i + (long)k * n] * m->EV[j + (long)k * p];
                            tmp[k] += pd * m->X[i + (long)j * n] * r * (m->ElogU[i + (long)k * n] + m->ElogV[j + (long)k * p]);
                        } else {
                            tmp[k] -= m->EU[i + (long)k * n] * m->EV[j + (long)k * p];
                            tmp[k] += m->X[i + (long)j * n] * r * (m->ElogU[i + (long)k * n] + m->ElogV[j + (long)k * p]);
                        }
                    }
                }
                for (int k = 0; k < K; k++) m->prob_S[j + (long)k * p] = orc_expit(orc_logit(m->prior_S[j]) + tmp[k]);
            }
        }
        free(vec); free(tmp);
    }
    for (long e = 0; e < (long)p * K; e++) m->S[e] = m->prob_S[e] >= m->sel_bound;
}

/* zi_gap_factor_model.cpp:346-367, zi_sparse_gap_factor_model.cpp:389-410 */
static void update_variational_zi_param(orc_model *m) {
    int n = m->n, p = m->p;
#pragma omp parallel for
    for (int j = 0; j < p; j++) {
        double *col = m->prob_D + (long)j * n;
        if (m->prior_D[j] == 1) { for (int i = 0; i < n; i++) col[i] = 1; }
        else if (m->prior_D[j] == 0) { for (int i = 0; i < n; i++) col[i] = 0; }
        else {
            double lg = orc_logit(m->prior_D[j]);
            for (int i = 0; i < n; i++) {
                if (m->X[i + (long)j * n] != 0) col[i] = 1;
                else col[i] = orc_expit(lg - m->UVt[i + (long)j * n]);
            }
        }
    }
}

/* gap...cpp:195-210, zi...cpp:122-143, sparse...cpp:112-133, zi_sparse...cpp:125-151 */
static void update_variational_param(orc_model *m) {
    update_variational_multinomial_param(m);
    update_variational_gamma_param(m);
    if (m->sparse) update_variational_sparse_param(m);
    update_poisson_param(m);
    if (m->zi) update_variational_zi_param(m);
}
/* model.cpp:243-246 */
void orc_update_param(orc_model *m) { update_variational_param(m); update_hyper_param(m); }

/* individual phases, exposed so tests can compare intermediate quantities */
void orc_phase_estep(orc_model *m) { update_variational_multinomial_param(m); }
void orc_phase_mstep(orc_model *m) { update_variational_gamma_param(m); }
void orc_phase_sparse(orc_model *m) { if (m->sparse) update_variational_sparse_param(m); }
void orc_phase_poisson(orc_model *m) { update_poisson_param(m); }
void orc_phase_zi(orc_model *m) { if (m->zi) update_variational_zi_param(m); }
void orc_phase_hyper(orc_model *m) { update_hyper_param(m); }

/* gap_factor_model.cpp:239-252 (model.cpp:249-252) */
void orc_prepare_next_iterate(orc_model *m) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    memcpy(m->a1o, m->a1, 8 * nK); memcpy(m->a2o, m->a2, 8 * nK); memcpy(m->b1o, m->b1, 8 * pK); memcpy(m->b2o, m->b2, 8 * pK);
    memcpy(m->alpha1o, m->alpha1, 8 * nK); memcpy(m->alpha2o, m->alpha2, 8 * nK);
    memcpy(m->beta1o, m->beta1, 8 * pK); memcpy(m->beta2o, m->beta2, 8 * pK);
}
static double sqn(const double *a, long len) { double s = 0; for (long i = 0; i < len; i++) s += a[i] * a[i]; return s; }
static double sqd(const double *a, const double *b, long len) { double s = 0; for (long i = 0; i < len; i++) s += (a[i] - b[i]) * (a[i] - b[i]); return s; }
/* gap_factor_model.cpp:255-262 */
void orc_gap_between_iterates(orc_model *m, double *abs_gap, double *norm_gap) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    double paramNorm = sqrt(sqn(m->a1o, nK) + sqn(m->a2o, nK) + sqn(m->b1o, pK) + sqn(m->b2o, pK));
    double diffNorm = sqrt(sqd(m->a1, m->a1o, nK) + sqd(m->a2, m->a2o, nK) + sqd(m->b1, m->b1o, pK) + sqd(m->b2, m->b2o, pK));
    *abs_gap = diffNorm; *norm_gap = diffNorm / paramNorm;
}

/* ------------------------------------------------------------------------------------ */
/* ELBO and monitors                                                                     */
/* ------------------------------------------------------------------------------------ */
/* individual ELBO terms, written by orc_elbo for the tests: 0 gammaU, 1 entU, 2 gammaV, 3 entV, 4 bernS_prior,
 * 5 bernS_self, 6 bernD_prior, 7 bernD_self, 8 sumUVt, 9 data, 10 lgammaX */
static double g_elbo_terms[11];
double *orc_elbo_terms(void) { return g_elbo_terms; }

/* gap...cpp:287-326, zi...cpp:170-215, sparse...cpp:160-228, zi_sparse...cpp:181-250 */
double orc_elbo(orc_model *m) {
    int n = m->n, p = m->p, K = m->K;
    long np = (long)n * p, nK = (long)n * K, pK = (long)p * K;
    double *t = g_elbo_terms; for (int i = 0; i < 11; i++) t[i] = 0;
    double res = 0;
    res += (t[0] = orc_gamma_loglike_logX(m->EU, m->ElogU, m->alpha1, m->alpha2, nK));
    res += (t[1] = gamma_entropy_sum(m->a1, m->a2, nK));
    res += (t[2] = orc_gamma_loglike_logX(m->EV, m->ElogV, m->beta1, m->beta2, pK));
    res += (t[3] = gamma_entropy_sum(m->b1, m->b2, pK));
    if (m->sparse) {
        res += (t[4] = orc_bernoulli_loglike_rowvec(m->prob_S, m->prior_S, p, K));
        res -= (t[5] = orc_bernoulli_loglike(m->prob_S, m->prob_S, pK));
    }
    if (m->zi) {
        res += (t[6] = orc_bernoulli_loglike_colvec(m->prob_D, m->prior_D, n, p));
        res -= (t[7] = orc_bernoulli_loglike(m->prob_D, m->prob_D, np));
    }
    if (m->sparse) update_poisson_param(m);   /* sparse...cpp:176, zi_sparse...cpp:199 */
    double s = 0;
    if (m->zi) {
#pragma omp parallel for reduction(+:s)
        for (long e = 0; e < np; e++) s += m->prob_D[e] * m->UVt[e];
    } else {
#pragma omp parallel for reduction(+:s)
        for (long e = 0; e < np; e++) s += m->UVt[e];
    }
    res -= (t[8] = s);
    intermediate_update_variational_multinomial_param(m);
    double tmp = 0;
    if (!m->sparse) {
#pragma omp parallel for reduction(+:tmp)
        for (int j = 0; j < p; j++)
            for (int i = 0; i < n; i++) {
                long e = i + (long)j * n;
                if (m->nz_only && m->X[e] == 0) continue;
                tmp += (m->zi ? m->prob_D[e] : 1.0) * m->X[e] * log(m->T[e]);
            }
    } else {
#pragma omp parallel
        {
            double *vec = (double *)malloc(8 * K);
#pragma omp for reduction(+:tmp)
            for (int j = 0; j < p; j++)
                for (int i = 0; i < n; i++) {
                    long e = i + (long)j * n;
                    if (m->nz_only && m->X[e] == 0) continue;
                    double tv = multinomial_vec(m, i, j, vec);
                    tv = tv > 0 ? tv : 1;
                    for (int k = 0; k < K; k++) {
                        double r = vec[k] / tv;
                        tmp += (m->zi ? m->prob_D[e] : 1.0) * m->prob_S[j + (long)k * p] * m->X[e] * r
                               * (m->ElogU[i + (long)k * n] + m->ElogV[j + (long)k * p]);
                    }
                }
            free(vec);
        }
    }
    res += (t[9] = tmp);
    tmp = 0;
#pragma omp parallel for reduction(+:tmp)
    for (long e = 0; e < np; e++) tmp += (m->zi ? m->prob_D[e] : 1.0) * lgamma(m->X[e] + 1);
    res -= (t[10] = tmp);
    return res;
}

/* rate used by deviance-type monitors: UVt (gap, sparse) or prob_D o UVt (zi variants) */
static double *monitor_rate(orc_model *m, const double *uvt) {
    long np = (long)m->n * m->p;
    double *r = dalloc(np, 0);
    for (long e = 0; e < np; e++) r[e] = m->zi ? m->prob_D[e] * uvt[e] : uvt[e];
    return r;
}
/* gap...cpp:278-284, zi...cpp:160-167, sparse...cpp:151-157, zi_sparse...cpp:171-178 */
double orc_loglikelihood(orc_model *m) {
    long np = (long)m->n * m->p, nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    double r1 = m->zi ? orc_zi_poisson_loglike(m->X, m->UVt, m->prob_D, m->n, m->p) : orc_poisson_loglike(m->X, m->UVt, np);
    double r2 = orc_gamma_loglike(m->EU, m->a1, m->a2, nK);
    double r3 = orc_gamma_loglike(m->EV, m->b1, m->b2, pK);
    double r = r1 + r2 + r3;
    if (m->zi) r += orc_bernoulli_loglike(m->prob_D, m->prob_D, np);
    if (m->sparse) r += orc_bernoulli_loglike(m->prob_S, m->prob_S, pK);
    return r;
}
/* gap...cpp:329-333, zi...cpp:218-229, sparse...cpp:231-235, zi_sparse...cpp:253-257 */
double orc_deviance(orc_model *m) {
    long np = (long)m->n * m->p;
    double *rate = monitor_rate(m, m->UVt);
    double r1 = orc_poisson_loglike(m->X, rate, np), r2 = orc_poisson_loglike(m->X, m->X, np);
    free(rate);
    return -2 * (r1 - r2);
}
/* gap...cpp:336-341, zi...cpp:232-245, sparse...cpp:238-243, zi_sparse...cpp:260-265 */
double orc_exp_deviance(orc_model *m) {
    int n = m->n, p = m->p; long np = (long)n * p;
    double *rate = monitor_rate(m, m->UVt);
    double *cm = dalloc(p, 0);
    for (int j = 0; j < p; j++) { double s = 0; for (int i = 0; i < n; i++) s += m->X[i + (long)j * n]; cm[j] = s / n; }
    double r1 = orc_poisson_loglike(m->X, rate, np), r2 = orc_poisson_loglike(m->X, m->X, np);
    double r3 = orc_poisson_loglike_vec(m->X, cm, n, p);
    free(rate); free(cm);
    return (r1 - r3) / (r2 - r3);
}

/* gap...cpp:420-450, zi...cpp:264-286, sparse...cpp:294-319, zi_sparse...cpp:288-313 */
double orc_partial_deviance(orc_model *m, const int *factor, int k) {
    long np = (long)m->n * m->p;
    double *uvt = dalloc(np, 0);
    uvt_into(m, uvt, factor, k);
    double *rate = monitor_rate(m, uvt);
    double r1 = orc_poisson_loglike(m->X, rate, np), r2 = orc_poisson_loglike(m->X, m->X, np);
    free(uvt); free(rate);
    return -2 * (r1 - r2);
}

static void permute_cols_d(double *A, int rows, int K, const int *order) {
    /* Eigen `A *= perm` with perm.indices()[k] = order[k]: new column order[k] = old column k
     * (PermutationMatrix P has P(indices[k], k) = 1, so (A P)(:, k) = A(:, ... ) -- see note below) */
    double *B = (double *)malloc(8 * (size_t)rows * K);
    /* (A*P)_{i,c} = sum_r A_{i,r} P_{r,c}; P_{r,c} = 1 iff r == indices[c]  => (A*P)(:,c) = A(:, indices[c]) */
    for (int c = 0; c < K; c++) memcpy(B + (long)c * rows, A + (long)order[c] * rows, 8 * (size_t)rows);
    memcpy(A, B, 8 * (size_t)rows * K); free(B);
}
static void permute_cols_i(int *A, int rows, int K, const int *order) {
    int *B = (int *)malloc(sizeof(int) * (size_t)rows * K);
    for (int c = 0; c < K; c++) memcpy(B + (long)c * rows, A + (long)order[c] * rows, sizeof(int) * (size_t)rows);
    memcpy(A, B, sizeof(int) * (size_t)rows * K); free(B);
}
/* model.cpp:137-187 (greedy forward selection) + gap...cpp:344-366 / sparse...cpp:250-275 (reorder) */
void orc_order_factor(orc_model *m) {
    int K = m->K;
    int *left = (int *)malloc(sizeof(int) * K), *chosen = (int *)malloc(sizeof(int) * K), *tmp = (int *)malloc(sizeof(int) * K);
    int nleft = K, nchosen = 0;
    for (int k = 0; k < K; k++) left[k] = k;
    for (int k = 0; k < K; k++) {
        double val_min = -1; int ind_min = 0, ind_left = 0;
        for (int ind = 0; ind < nleft; ind++) {
            int c = 0;
            for (int l = 0; l < nchosen; l++) tmp[c++] = chosen[l];
            tmp[c++] = left[ind];
            for (int l = 0; l < nleft; l++) if (l != ind) tmp[c++] = left[l];
            double res = orc_partial_deviance(m, tmp, k + 1);
            if (ind == 0) { val_min = res; ind_min = left[ind]; ind_left = ind; }
            if (res < val_min) { val_min = res; ind_min = left[ind]; ind_left = ind; }
        }
        chosen[nchosen++] = ind_min;
        for (int l = ind_left; l < nleft - 1; l++) left[l] = left[l + 1];
        nleft--;
    }
    for (int k = 0; k < K; k++) m->factor_order[k] = chosen[k];
    free(left); free(chosen); free(tmp);
    /* reorder_factor */
    const int *o = m->factor_order; int n = m->n, p = m->p;
    permute_cols_d(m->EU, n, K, o); permute_cols_d(m->EV, p, K, o); permute_cols_d(m->ElogU, n, K, o); permute_cols_d(m->ElogV, p, K, o);
    permute_cols_d(m->a1, n, K, o); permute_cols_d(m->a2, n, K, o); permute_cols_d(m->b1, p, K, o); permute_cols_d(m->b2, p, K, o);
    permute_cols_d(m->alpha1, n, K, o); permute_cols_d(m->alpha2, n, K, o); permute_cols_d(m->beta1, p, K, o); permute_cols_d(m->beta2, p, K, o);
    if (m->sparse) { permute_cols_i(m->S, p, K, o); permute_cols_d(m->prob_S, p, K, o); }
}

/* matrix::rv_coeff (utils/matrix.h:100-104): trace(A A' B B') / sqrt(trace(A A' A A') trace(B B' B B')).  The reference
 * forms the n x n products; trace(A A' B B') = ||A' B||_F^2, so K x K Gram matrices give the same number. */
static double rv_coeff(const double *A, const double *B, int rows, int K) {
    double num = 0, da = 0, db = 0;
    for (int k = 0; k < K; k++)
        for (int l = 0; l < K; l++) {
            double ab = 0, aa = 0, bb = 0;
            for (int r = 0; r < rows; r++) {
                ab += A[r + (long)k * rows] * B[r + (long)l * rows];
                aa += A[r + (long)k * rows] * A[r + (long)l * rows];
                bb += B[r + (long)k * rows] * B[r + (long)l * rows];
            }
            num += ab * ab; da += aa * aa; db += bb * bb;
        }
    return num / sqrt(da * db);
}
/* gap_factor_model::custom_conv_criterion (gap_factor_model.cpp:265-269), conv_mode = 2 */
double orc_custom_conv_criterion(orc_model *m) {
    long nK = (long)m->n * m->K, pK = (long)m->p * m->K;
    double *uo = (double *)malloc(8 * (size_t)nK), *vo = (double *)malloc(8 * (size_t)pK);
    for (long e = 0; e < nK; e++) uo[e] = m->a1o[e] / m->a2o[e];
    for (long e = 0; e < pK; e++) vo[e] = m->b1o[e] / m->b2o[e];
    double r1 = rv_coeff(m->EU, uo, m->n, m->K), r2 = rv_coeff(m->EV, vo, m->p, m->K);
    free(uo); free(vo);
    return 1 - (r1 < r2 ? r1 : r2);
}

/* ------------------------------------------------------------------------------------ */
/* iteration driver: algorithm.h:158-210 (optimize), :391-397 (monitor), :412-450         */
/* (check_convergence), :402-404 (optim_criterion).  Per-iteration vectors must hold      */
/* iter_max doubles.  Returns nb_iter; *iter_io carries m_iter in and out (multi-init).   */
/* ------------------------------------------------------------------------------------ */
int orc_optimize(orc_model *m, int iter_max, int iter_min, double epsilon, int additional_iter, int conv_mode,
                 int monitor, int *iter_io, int *converged_out, double *loss_out,
                 double *conv_crit, double *abs_gap_v, double *norm_gap_v,
                 double *loglik_v, double *elbo_v, double *deviance_v) {
    int iter = iter_io ? *iter_io : 0;
    int converged = 0, nb_iter_stab = 0;
    while (iter < iter_max && !converged) {
        orc_update_param(m);
        if (monitor) {
            if (loglik_v) loglik_v[iter] = orc_loglikelihood(m);
            if (elbo_v) elbo_v[iter] = orc_elbo(m);
            if (deviance_v) deviance_v[iter] = orc_deviance(m);
        }
        double abs_gap, norm_gap;
        orc_gap_between_iterates(m, &abs_gap, &norm_gap);
        if (conv_mode == 0) conv_crit[iter] = abs_gap;
        else if (conv_mode == 1) conv_crit[iter] = norm_gap;
        else if (conv_mode == 2) conv_crit[iter] = orc_custom_conv_criterion(m);   /* algorithm.h:424-425 */
        else return -1;   /* algorithm.h:427 */
        if (monitor) { if (abs_gap_v) abs_gap_v[iter] = abs_gap; if (norm_gap_v) norm_gap_v[iter] = norm_gap; }
        if (fabs(conv_crit[iter]) < epsilon) nb_iter_stab++; else nb_iter_stab = 0;
        if (nb_iter_stab > additional_iter && iter > iter_min) converged = 1;
        orc_prepare_next_iterate(m);
        iter++;
    }
    if (loss_out) *loss_out = orc_elbo(m);
    if (iter_io) *iter_io = iter;
    if (converged_out) *converged_out = converged;
    return iter;
}

/* one E-step worth of work only (for the CPU baseline timing of the E-step kernel) */
void orc_estep_only(orc_model *m) { update_variational_multinomial_param(m); }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int nt) {
#ifdef _OPENMP
    omp_set_num_threads(nt);
#else
    (void)nt;
#endif
}
