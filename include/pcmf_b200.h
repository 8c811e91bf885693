/*
 * pcmf_b200.h -- C ABI of libpcmf_b200.so: the B200-native variational-EM core of pCMF
 * (Gamma-Poisson factor model + zero-inflated + sparse-V variants).
 *
 * This is the drop-in boundary for the reference's native entry points.  The reference's R package calls
 *   .Call(_pCMF_run_gap_factor, ...25 args)            pkg/src/RcppExports.cpp:12-44,   pkg/R/RcppExports.R:212
 *   .Call(_pCMF_run_zi_gap_factor, ...27 args)         pkg/src/RcppExports.cpp:111-145, pkg/R/RcppExports.R:833
 *   .Call(_pCMF_run_sparse_gap_factor, ...28 args)     pkg/src/RcppExports.cpp:73-108,  pkg/R/RcppExports.R:595
 *   .Call(_pCMF_run_zi_sparse_gap_factor, ...30 args)  pkg/src/RcppExports.cpp:148-185, pkg/R/RcppExports.R:1085
 * which instantiate wrapper_gap_factor<model, variational_em_algo> (pkg/src/wrapper_gap_factor.h:164-393) or
 * wrapper_sparse_gap_factor<...> (pkg/src/wrapper_sparse_gap_factor.h:176-435).  A thin Rcpp shim keeps those
 * four symbols and forwards to pcmf_fit() below (INTEGRATION.md shows the shim); the four models are selected
 * by (zero_inflation, sparsity).  No torch / Rcpp / Eigen types appear here: plain pointers and sizes only.
 *
 * Conventions
 *   - every matrix crossing this boundary is COLUMN-MAJOR FP64 (R / Eigen layout), caller-owned, never freed or
 *     written by the library unless it is an output buffer;
 *   - every function returns 0 on success or a PCMF_E* code and writes a message to errbuf (no exceptions cross);
 *   - the library is synchronous; all kernels run on options.stream (or a private stream);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns PCMF_ECUDA.
 */
#ifndef PCMF_B200_H
#define PCMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCMF_OK 0
#define PCMF_EINVAL 1      /* bad argument; errbuf carries the reference's Rcpp::stop text where one exists */
#define PCMF_ECUDA 2       /* CUDA runtime failure or no device */
#define PCMF_ECOMM 3       /* all-reduce callback failed */
#define PCMF_EUNSUPPORTED 4
#define PCMF_EINTERRUPT 5  /* progress callback asked to stop (Rcpp::checkUserInterrupt, algorithm.h:170-174) */

typedef struct pcmf_solver pcmf_solver;   /* opaque */

/* Sum all-reduce of `count` doubles that live in DEVICE memory at `devptr`, in place, across all ranks.
 * Called on the calling host thread; the library has enqueued all producers of the buffer on `stream`
 * and enqueues the consumers on `stream` after the callback returns, so the callback may be asynchronous
 * with respect to the host as long as it is ordered on `stream` (torch.distributed/NCCL semantics).
 * Return 0 on success.  Replaces the reference's OpenMP `declare reduction(+: Eigen::MatrixXd ...)` over
 * thread-private accumulators (pkg/src/gap_factor_model.cpp:474-478). */
typedef int (*pcmf_allreduce_fn)(void *ctx, void *devptr, size_t count, void *stream);

/* Progress / interrupt hook, called on the calling thread every `progress_every` iterations
 * (reference: Rcout << "iter" + Rcpp::checkUserInterrupt() every 10 iterations when verbose, algorithm.h:170-174).
 * Return non-zero to stop (-> PCMF_EINTERRUPT). */
typedef int (*pcmf_progress_fn)(void *ctx, int iter, double conv_crit);

/* The count matrix held by THIS process: all p genes of n_local cells (rows).  Exactly one of the four
 * representations is non-NULL.  Reference: SEXP X, INTSXP or REALSXP n x p column-major
 * (wrapper_gap_factor.h:204-211), copied into the model (model.cpp:83). */
typedef struct {
    int64_t n_local;           /* rows held here */
    int64_t n_global;          /* total rows over all ranks (= n_local on one GPU) */
    int64_t row_offset;        /* global index of local row 0 (only used for the seeded random init) */
    int32_t p;
    const double *dense_f64;   /* n_local x p column-major, or NULL */
    const int32_t *dense_i32;  /* n_local x p column-major, or NULL */
    const int64_t *csr_rowptr; /* n_local+1, or NULL.  CSR may live in host OR device memory (csr_on_device) */
    const int32_t *csr_colidx; /* nnz, columns ascending within a row not required */
    const float *csr_val_f32;  /* nnz (counts are exact in FP32 up to 2^24), or NULL */
    const double *csr_val_f64; /* nnz, or NULL */
    int32_t csr_on_device;     /* 1: the three CSR arrays are device pointers on options.device */
    /* Matrix::dgCMatrix as R holds it (slots @p, @i, @x), host memory: compressed columns (genes), 0-based int32
     * cell indices, double values; explicit zeros are dropped.  Converted to the cell-major CSR on the device.
     * (The reference only accepts a base matrix, wrapper_gap_factor.h:204-211) */
    const int32_t *csc_colptr; /* p+1, or NULL */
    const int32_t *csc_rowidx; /* nnz */
    const double *csc_val_f64; /* nnz */
} pcmf_problem;

/* Mirrors the scalar arguments of run_*_gap_factor (pkg/R/RcppExports.R:212 etc.) + GPU plumbing. */
typedef struct {
    int32_t K;
    int32_t zero_inflation;    /* pCMF.R:198 */
    int32_t sparsity;          /* pCMF.R:198 */
    double sel_bound;          /* wrapper_sparse_gap_factor.h:177, must be in [0,1] (:236) */
    int32_t iter_max;          /* default 1000 */
    int32_t iter_min;          /* default iter_max/2 (pCMF.R:203-205) */
    double epsilon;            /* default 1e-2 */
    int32_t additional_iter;   /* default 10 */
    int32_t conv_mode;         /* 0 absolute gap, 1 normalised gap, 2 RV coefficient (algorithm.h:412-450, gap...cpp:265-269) */
    int32_t monitor;           /* 1: record ELBO, log-likelihood and deviance every iteration like monitor=TRUE does
                                * (algorithm.h:391-397); 2: ELBO only (extension: skips the two extra passes);
                                * 0: none.  norm/abs gap are always recorded */
    int32_t reorder_factor;    /* order factors by deviance after the loop (model.cpp:137-187) */
    int32_t ninit;             /* multi-init (wrapper_gap_factor.h:296-350) */
    int32_t iter_init;
    int32_t has_seed;
    uint32_t seed;
    int32_t verbose;
    /* GPU plumbing */
    int32_t device;            /* CUDA device ordinal */
    void *stream;              /* cudaStream_t to run on, or NULL for a private stream */
    int32_t rank, world;       /* cell-sharded run: this rank / number of ranks (1 = single GPU) */
    pcmf_allreduce_fn allreduce;
    void *allreduce_ctx;
    pcmf_progress_fn progress;
    void *progress_ctx;
    int32_t progress_every;    /* default 10 */
    /* precision of the hot kernels.  0 (default): FP64 mode, results within 1e-6 of the reference's C++ path (the parity
     * tests hold 1e-8).  1: FP32 mode -- the dense zero-inflation passes run on the tcgen05 tensor cores with TF32
     * operands split three ways and FP32 accumulation in tensor memory, the non-zero passes gather FP32 copies of the
     * factor statistics; everything that is summed over cells or genes, the M-step and the ELBO stay FP64.  Stated bound
     * (tests/test_gpu_fp32.py): variational parameters within 1e-4, ELBO within 1e-5 relative of the FP64 path.
     * The reference is FP64 only (MatrixXd everywhere); there is no reference file for this mode. */
    int32_t precision;
    /* > 1: pcmf_fit shards the cells over devices 0 .. ngpus-1 ITSELF (one host thread per device, NCCL communicator
     * created inside the library, no callback, no launcher): the GPU counterpart of `ncores` -> omp_set_num_threads
     * (wrapper_gap_factor.h:271-274).  Input and output buffers are the whole matrices, as on one GPU.  0 / 1: one
     * device (`device`), or an external launcher that sets rank / world / allreduce. */
    int32_t ngpus;
    int32_t dev_flags;         /* 0.  Development switches for A/B measurements and tests (include/pcmf_b200_dev.h) */
} pcmf_options;

/* Optional warm-start values (NULL = absent), column-major, sizes for the LOCAL rows.
 * Reference: the Nullable<MatrixXd> arguments U,V,a1,a2,b1,b2,alpha*,beta*,prob_S,prior_S of run_*
 * and the dispatch at wrapper_sparse_gap_factor.h:389-417. */
typedef struct {
    const double *U, *V;                         /* n_local x K, p x K  -> init_from_factor (gap_factor_model.cpp:135-144) */
    const double *a1, *a2, *b1, *b2;             /* -> init_variational_param (gap_factor_model.cpp:113-123) */
    const double *alpha1, *alpha2, *beta1, *beta2; /* hyper_rows = 0: K-vectors (identical rows, what every state reachable
                                                    * from Ones has); hyper_rows = 1: the reference's n_local x K / p x K
                                                    * matrices as given (gap_factor_model.cpp:126-132).  As in the reference's
                                                    * else-if chain (wrapper_gap_factor.h:352-375) they are only looked at when
                                                    * neither U, V nor a1..b2 are given: init_hyper_param alone, the variational
                                                    * state stays at the constructor's Ones, no random draw, no hyper update */
    const double *prob_S, *prior_S;              /* p x K, p -> init_sparse_param (sparse_gap_factor_model.cpp:96-104) */
    const double *prob_D, *prior_D;              /* n_local x p (dense), p -> init_zi_param(prob_D, prior_D)
                                                  * (zi_gap_factor_model.cpp:110-114; both or neither).  The dense matrix is
                                                  * only held until the first zero-inflation update replaces it */
    int32_t hyper_rows;                          /* layout of alpha1..beta2, see above */
} pcmf_init;

/* Caller-allocated outputs (NULL = not wanted); column-major; U-side sizes are for the LOCAL rows.
 * Field-for-field the reference's return list (algorithm.h:355-377, gap_factor_model.cpp:369-393,
 * sparse_gap_factor_model.cpp:278-286, zi_gap_factor_model.cpp:248-255). */
typedef struct {
    double *EU, *ElogU, *a1, *a2, *alpha1, *alpha2;   /* n_local x K each (hyper_params replicated over rows) */
    double *EV, *ElogV, *b1, *b2, *beta1, *beta2;     /* p x K each */
    double *prob_S, *prior_prob_S;                    /* p x K, p     (sparse) */
    int32_t *S;                                       /* p x K        (sparse) */
    double *prob_D;                                   /* n_local x p  (ZI; dense, only if the caller wants it) */
    double *freq_D, *prior_prob_D;                    /* p, p         (ZI) */
    int32_t *factor_order;                            /* K */
    /* convergence / monitor vectors: caller allocates iter_max doubles each (multi-init needs iter_init <= iter_max:
     * the kept run's first iter_init entries are written there, wrapper_gap_factor.h:296-350) */
    double *conv_crit, *norm_gap, *abs_gap, *optim_criterion, *loglikelihood, *deviance;
    /* scalars written by the library */
    int32_t nb_iter;
    int32_t converged;
    double loss;       /* ELBO after the loop (algorithm.h:209) */
    double exp_dev;    /* percentage of explained deviance (algorithm.h:367) */
    double seconds_setup, seconds_iter, seconds_total;   /* host wall-clock, informational */
} pcmf_result;

/* library / device information */
const char *pcmf_version(void);
int pcmf_device_count(void);

/* One-call fit: wrapper_*gap_factor<model, variational_em_algo> end to end
 * (init -> optimize -> order_factor -> get_output), host buffers in, host buffers out. */
int pcmf_fit(const pcmf_problem *prob, const pcmf_options *opt, const pcmf_init *init, pcmf_result *res,
             char *errbuf, size_t errlen);

/* Step-wise API (same engine; used by the parity tests, the multi-GPU host driver and the bench). */
int pcmf_create(const pcmf_problem *prob, const pcmf_options *opt, const pcmf_init *init, pcmf_solver **out,
                char *errbuf, size_t errlen);
/* algorithm<model>::optimize (algorithm.h:158-210): runs until iter_max / convergence, continuing from the current
 * iteration counter.  Writes the per-iteration vectors and scalars of `res` (other fields untouched). */
int pcmf_optimize(pcmf_solver *s, pcmf_result *res, char *errbuf, size_t errlen);
/* exactly `niter` x { update_param; [monitor]; check_convergence; prepare_next_iterate } without the stopping rule
 * (teacher-forced parity tests and fixed-length timing).  conv_crit_out: niter doubles or NULL.
 * elbo_out: niter doubles or NULL; when given, the ELBO is evaluated after every iteration like
 * algorithm::monitor() does when monitor=TRUE (algorithm.h:391-397). */
int pcmf_iterate(pcmf_solver *s, int niter, double *conv_crit_out, double *elbo_out, char *errbuf, size_t errlen);
/* ELBO of the current state (model::optim_criterion, gap_factor_model.cpp:287-326 and variants);
 * terms (nullable): 11 doubles in the order gammaU, entU, gammaV, entV, bernS_prior, bernS_self, bernD_prior,
 * bernD_self, sumUVt, data, lgammaX */
int pcmf_elbo(pcmf_solver *s, double *elbo_out, double *terms, char *errbuf, size_t errlen);
/* multi-init (wrapper_gap_factor.h:296-350, wrapper_sparse_gap_factor.h:328-387): options.ninit runs of
 * options.iter_init iterations from successive random initialisations, keeps the run with the smallest loss (sic) and
 * leaves the solver at iteration iter_init of that run; pcmf_optimize then continues.  The per-iteration vectors of
 * `res` hold the kept run's first iter_init entries. */
int pcmf_multi_init(pcmf_solver *s, pcmf_result *res, char *errbuf, size_t errlen);
/* variational_em_algo::order_factor (algorithm_variational_EM.h:189-192): greedy ordering by partial deviance
 * (model.cpp:137-187) + reorder_factor; permutes every K-indexed array of the state. */
int pcmf_order_factor(pcmf_solver *s, char *errbuf, size_t errlen);
/* monitors (gap_factor_model.cpp:278-284, :329-341 and variants): any pointer may be NULL */
int pcmf_monitors(pcmf_solver *s, double *loglikelihood, double *deviance, double *exp_deviance,
                  char *errbuf, size_t errlen);
/* copies the current state into caller buffers (algorithm::get_output) */
int pcmf_get_output(pcmf_solver *s, pcmf_result *res, char *errbuf, size_t errlen);
/* overwrite the variational state (teacher forcing): same post-steps as init_variational_param */
int pcmf_set_variational(pcmf_solver *s, const double *a1, const double *a2, const double *b1, const double *b2,
                         char *errbuf, size_t errlen);
void pcmf_destroy(pcmf_solver *s);

/* Per-gene statistics over the LOCAL cells of a CSR matrix (host or device resident): sum_i x, sum_i x^2, #non-zeros
 * and (colmax, may be NULL) the largest count.  Host outputs of length p.  This is what the R wrapper's prior_S
 * heuristic (pCMF.R:228-229: 1 - exp(-sd_j / mean(X[X != 0]))) and prefilter() (prefiltering.R:127-146: column max,
 * share of non-null entries, the same sd heuristic) need when X never exists as a dense matrix.  Counts are >= 0. */
int pcmf_csr_column_stats(int device, int64_t n_local, int32_t p, const int64_t *rowptr, const int32_t *colidx,
                          const float *val_f32, int32_t on_device, double *colsum, double *colsumsq, double *colnnz,
                          double *colmax, char *errbuf, size_t errlen);
/* The library keeps the device memory a finished fit releases for the next fit of the process;
 * this returns all of it to the driver. */
int pcmf_release_cached_memory(void);
#ifdef __cplusplus
}
#endif
#endif /* PCMF_B200_H */
