/*
 * pcmf_b200_dev.h -- development / measurement entry points of libpcmf_b200.so.  NOT part of the drop-in boundary: nothing
 * here replaces a reference interface, and a pCMF maintainer links against include/pcmf_b200.h only.  bench.py, the parity
 * tests and the profiling scripts use these: per-phase timings, the synthetic-input generator, microbenchmarks that give
 * the roofline denominators, and the bits of pcmf_options.dev_flags.
 */
#ifndef PCMF_B200_DEV_H
#define PCMF_B200_DEV_H

#include "pcmf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* pcmf_options.dev_flags: A/B switches, results are the same within rounding */
#define PCMF_DEV_THREAD_ZI 4        /* thread-per-row ZI kernels instead of the DMMA tiles */
#define PCMF_DEV_ZI_MT2 8           /* two row tiles per warp in the DMMA passes */
#define PCMF_DEV_NO_STORED_Q 16     /* the cell pass recomputes T instead of reading the gene passes' q */
#define PCMF_DEV_DENSE_ZI_INIT 32   /* dense first ZI pass (c = -inf) instead of the non-zero shortcut */
#define PCMF_DEV_FORCE_PSTORE 64    /* keep prob_D between the ZI passes whatever the size */
#define PCMF_DEV_NO_PSTORE 128      /* never keep it */
#define PCMF_DEV_ONE_CELL_BLOCK 512 /* gene passes in one cell block */
#define PCMF_DEV_TINY_CELL_BLOCKS 1024 /* 64-cell blocks */
#define PCMF_DEV_FORCE_TILES 2048   /* tiled non-zero passes whatever the size */
#define PCMF_DEV_NO_TILES 4096      /* CSR / CSC non-zero passes whatever the size */
#define PCMF_DEV_ZI_ALL_DMMA 8192   /* dense ZI passes with both products on the FP64 tensor path (no integer tcgen05 product) */
#define PCMF_DEV_ORDER_EXACT 16384   /* factor ordering with FP64 logarithms in every round (no FP32 pass with error bounds first) */

/* per-phase device timings of the last pcmf_iterate/pcmf_optimize call, milliseconds summed over iterations:
 * names[i] points to a static string; returns the number of phases written (<= cap) */
int pcmf_phase_times(pcmf_solver *s, const char **names, double *ms, int64_t *launches, int cap);
/* number of kernel launches issued by the solver since creation */
int64_t pcmf_launch_count(pcmf_solver *s);
/* Synthetic benchmark input generated ON DEVICE with the package's generative model
 * (pkg/R/data_generation.R:393-463: X = Poisson(U V^T) * Bernoulli(prob1_j)), straight into CSR.
 * U: n_local x K_true, V: p x K_true column-major host arrays; prob1: p host (NULL = no drop-out).
 * Two-pass (count, fill); outputs are device pointers owned by the caller (free with pcmf_device_free). */
int pcmf_generate_counts_device(int device, int64_t n_local, int32_t p, int32_t K_true, const double *U,
                                const double *V, const double *prob1, uint64_t seed, int64_t row_offset,
                                int64_t **rowptr_dev, int32_t **colidx_dev, float **val_dev, int64_t *nnz_out,
                                char *errbuf, size_t errlen);
/* Measured FP64 FMA throughput of the device in TFLOP/s (dependent-chain-free DFMA microbenchmark, ~50 ms): the
 * roofline denominator of the dense zero-inflation passes, which are FP64-pipe bound, not HBM bound. */
int pcmf_measure_fp64_tflops(int device, double *tflops_out);
/* Same for the FP64 tensor pipe (mma.sync m8n8k4 DMMA), which the dense zero-inflation passes run on. */
int pcmf_measure_fp64_mma_tflops(int device, double *tflops_out);
/* Measured gather bandwidth (GB/s) of 160-byte K-vector rows out of a `rows`-row FP64 table: the ceiling of the E-step
 * non-zero passes, which read one or two such rows per non-zero through L2. */
int pcmf_measure_gather_gbs(int device, int64_t rows, double *gbs_out);
int pcmf_device_free(int device, void *ptr);
/* copy device CSR arrays to host buffers (bench e2e leg) */
int pcmf_device_to_host(int device, void *dst_host, const void *src_dev, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PCMF_B200_DEV_H */
