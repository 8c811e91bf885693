/* A plain C99 consumer of include/pcmf_b200.h, linked against libpcmf_b200.so without Python in between: what an Rcpp
 * shim or any other host language sees.  Used by tests/test_cabi_client.py.
 *
 *   cabi_client sizes                       print sizeof / offsetof of the four boundary structs
 *   cabi_client fit <in.bin> <out.bin>      one pcmf_fit call (= one run_*_gap_factor call of the reference);
 *                                           environment: PCMF_CABI_NGPUS -> pcmf_options.ngpus (the library shards the cells
 *                                           over that many devices itself), PCMF_CABI_PRECISION -> pcmf_options.precision
 *
 * in.bin : int32 n, p, K, zi, sparse, iter_max, reorder | double X[n*p] (column-major) | a1[n*K] a2[n*K] b1[p*K] b2[p*K]
 *          | prob_S[p*K] prior_S[p]
 * out.bin: int32 rc, nb_iter, converged | double loss, exp_dev | EU[n*K] EV[p*K] a1 a2 b1 b2 | prior_prob_S[p] |
 *          prior_prob_D[p] | optim_criterion[iter_max] | int32 factor_order[K] */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcmf_b200.h"

static double *rd(FILE *f, size_t count) {
    double *v = (double *)malloc(sizeof(double) * (count ? count : 1));
    if (fread(v, sizeof(double), count, f) != count) { fprintf(stderr, "short read\n"); exit(3); }
    return v;
}
static double *zeros(size_t count) { return (double *)calloc(count ? count : 1, sizeof(double)); }

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "sizes") == 0) {
        printf("pcmf_problem %zu csr_on_device %zu csc_val_f64 %zu\n", sizeof(pcmf_problem), offsetof(pcmf_problem, csr_on_device),
               offsetof(pcmf_problem, csc_val_f64));
        printf("pcmf_options %zu sel_bound %zu seed %zu stream %zu precision %zu ngpus %zu dev_flags %zu\n", sizeof(pcmf_options), offsetof(pcmf_options, sel_bound),
               offsetof(pcmf_options, seed), offsetof(pcmf_options, stream), offsetof(pcmf_options, precision), offsetof(pcmf_options, ngpus),
               offsetof(pcmf_options, dev_flags));
        printf("pcmf_init %zu prior_D %zu hyper_rows %zu\n", sizeof(pcmf_init), offsetof(pcmf_init, prior_D), offsetof(pcmf_init, hyper_rows));
        printf("pcmf_result %zu S %zu nb_iter %zu loss %zu seconds_total %zu\n", sizeof(pcmf_result), offsetof(pcmf_result, S),
               offsetof(pcmf_result, nb_iter), offsetof(pcmf_result, loss), offsetof(pcmf_result, seconds_total));
        printf("version %s\n", pcmf_version());
        return 0;
    }
    if (argc < 4 || strcmp(argv[1], "fit") != 0) { fprintf(stderr, "usage: cabi_client sizes | fit in.bin out.bin\n"); return 2; }
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 2; }
    int32_t hdr[7];
    if (fread(hdr, sizeof(int32_t), 7, f) != 7) return 3;
    const int n = hdr[0], p = hdr[1], K = hdr[2], zi = hdr[3], sparse = hdr[4], iter_max = hdr[5], reorder = hdr[6];
    double *X = rd(f, (size_t)n * p);
    double *a1 = rd(f, (size_t)n * K), *a2 = rd(f, (size_t)n * K), *b1 = rd(f, (size_t)p * K), *b2 = rd(f, (size_t)p * K);
    double *prob_S = rd(f, (size_t)p * K), *prior_S = rd(f, (size_t)p);
    fclose(f);

    pcmf_problem prob;  memset(&prob, 0, sizeof prob);
    prob.n_local = prob.n_global = n;  prob.p = p;  prob.dense_f64 = X;
    pcmf_options opt;  memset(&opt, 0, sizeof opt);
    opt.K = K;  opt.zero_inflation = zi;  opt.sparsity = sparse;  opt.sel_bound = 0.5;
    opt.iter_max = iter_max;  opt.iter_min = iter_max / 2;  opt.epsilon = 1e-2;  opt.additional_iter = 10;  opt.conv_mode = 1;
    opt.monitor = 1;  opt.reorder_factor = reorder;  opt.ninit = 1;  opt.iter_init = 100;  opt.world = 1;  opt.progress_every = 10;
    if (getenv("PCMF_CABI_NGPUS")) opt.ngpus = atoi(getenv("PCMF_CABI_NGPUS"));
    if (getenv("PCMF_CABI_PRECISION")) opt.precision = atoi(getenv("PCMF_CABI_PRECISION"));
    pcmf_init init;  memset(&init, 0, sizeof init);
    init.a1 = a1;  init.a2 = a2;  init.b1 = b1;  init.b2 = b2;
    if (sparse) { init.prob_S = prob_S;  init.prior_S = prior_S; }
    pcmf_result res;  memset(&res, 0, sizeof res);
    res.EU = zeros((size_t)n * K);  res.EV = zeros((size_t)p * K);
    res.a1 = zeros((size_t)n * K);  res.a2 = zeros((size_t)n * K);  res.b1 = zeros((size_t)p * K);  res.b2 = zeros((size_t)p * K);
    res.prior_prob_S = zeros(p);  res.prior_prob_D = zeros(p);
    res.conv_crit = zeros(iter_max);  res.norm_gap = zeros(iter_max);  res.abs_gap = zeros(iter_max);
    res.optim_criterion = zeros(iter_max);  res.loglikelihood = zeros(iter_max);  res.deviance = zeros(iter_max);
    res.factor_order = (int32_t *)calloc(K, sizeof(int32_t));
    char err[512];  err[0] = 0;
    const int rc = pcmf_fit(&prob, &opt, &init, &res, err, sizeof err);
    if (rc != PCMF_OK) fprintf(stderr, "pcmf_fit: rc=%d %s\n", rc, err);

    FILE *o = fopen(argv[3], "wb");
    if (!o) { perror(argv[3]); return 2; }
    int32_t ih[3];  ih[0] = rc;  ih[1] = res.nb_iter;  ih[2] = res.converged;
    double dh[2];  dh[0] = res.loss;  dh[1] = res.exp_dev;
    fwrite(ih, sizeof(int32_t), 3, o);  fwrite(dh, sizeof(double), 2, o);
    fwrite(res.EU, sizeof(double), (size_t)n * K, o);  fwrite(res.EV, sizeof(double), (size_t)p * K, o);
    fwrite(res.a1, sizeof(double), (size_t)n * K, o);  fwrite(res.a2, sizeof(double), (size_t)n * K, o);
    fwrite(res.b1, sizeof(double), (size_t)p * K, o);  fwrite(res.b2, sizeof(double), (size_t)p * K, o);
    fwrite(res.prior_prob_S, sizeof(double), p, o);  fwrite(res.prior_prob_D, sizeof(double), p, o);
    fwrite(res.optim_criterion, sizeof(double), iter_max, o);
    fwrite(res.factor_order, sizeof(int32_t), K, o);
    fclose(o);
    return rc == PCMF_OK ? 0 : 10 + (rc < 0 ? -rc : rc);
}
