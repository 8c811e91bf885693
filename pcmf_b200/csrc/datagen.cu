// On-device synthetic count generator for the large benchmark configurations.
// Generative model of the reference package (pkg/R/data_generation.R:393-463):
//   Lambda = U V^T, Xnzi ~ Poisson(Lambda), D_ij ~ Bernoulli(prob1_j) (prob1 = probability of NOT dropping out), X = Xnzi * D
// written straight into CSR (two passes: count, fill).  Philox counter-based streams keyed by the GLOBAL entry index,
// so the matrix does not depend on the launch configuration or on how cells are sharded over GPUs.
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>
#include <curand_kernel.h>

#include "../../include/pcmf_b200_dev.h"

namespace {

struct GenArgs {
    int64_t n, row_offset;
    int32_t p;
    int K;
    const double *U;   // n x K row-major
    const double *V;   // p x K row-major
    const double *prob1;   // p or null
    unsigned long long seed;
};

__device__ __forceinline__ int draw_entry(const GenArgs &a, int64_t i, int j) {
    curandStatePhilox4_32_10_t st;
    curand_init(a.seed, (unsigned long long)((a.row_offset + i) * (int64_t)a.p + j), 0ull, &st);
    if (a.prob1) {
        const float u = curand_uniform(&st);           // (0, 1]
        if (!(u <= (float)a.prob1[j])) return 0;       // dropped out
    }
    double lam = 0.0;
    const double *u = a.U + i * a.K, *v = a.V + (int64_t)j * a.K;
    for (int k = 0; k < a.K; ++k) lam = fma(u[k], v[k], lam);
    return (int)curand_poisson(&st, lam);
}

// one warp per cell; pass 0 counts, pass 1 writes (ordered by gene inside the cell)
template <bool FILL>
__global__ void k_generate(GenArgs a, int64_t *rowcount, const int64_t *rowptr, int32_t *colidx, float *val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < a.n; i += nw) {
        int64_t base = FILL ? rowptr[i] : 0;
        int64_t cnt = 0;
        for (int j0 = 0; j0 < a.p; j0 += 32) {
            const int j = j0 + lane;
            const int x = j < a.p ? draw_entry(a, i, j) : 0;
            const unsigned m = __ballot_sync(0xffffffffu, x != 0);
            if (FILL && x != 0) {
                const int64_t pos = base + __popc(m & ((1u << lane) - 1u));
                colidx[pos] = j;
                val[pos] = (float)x;
            }
            base += __popc(m);
            cnt += __popc(m);
        }
        if (!FILL && lane == 0) rowcount[i] = cnt;
    }
}

__global__ void k_transpose_in(int64_t rows, int K, const double *src_colmajor, double *dst_rowmajor) {
    const int64_t total = rows * K;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % K);
        const int64_t r = o / K;
        dst_rowmajor[o] = src_colmajor[r + (int64_t)k * rows];
    }
}

#define GK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            if (errbuf && errlen) snprintf(errbuf, errlen, "CUDA error %s at %s:%d", cudaGetErrorName(e_), __FILE__, __LINE__); \
            return PCMF_ECUDA;                                                                               \
        }                                                                                                    \
    } while (0)

}  // namespace

extern "C" int pcmf_generate_counts_device(int device, int64_t n_local, int32_t p, int32_t K_true, const double *U, const double *V,
                                           const double *prob1, uint64_t seed, int64_t row_offset, int64_t **rowptr_dev,
                                           int32_t **colidx_dev, float **val_dev, int64_t *nnz_out, char *errbuf, size_t errlen) {
    if (!U || !V || !rowptr_dev || !colidx_dev || !val_dev || !nnz_out || n_local <= 0 || p <= 0 || K_true <= 0) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "bad argument to pcmf_generate_counts_device");
        return PCMF_EINVAL;
    }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "no CUDA device available");
        return PCMF_ECUDA;
    }
    GK(cudaSetDevice(device));
    double *dU = nullptr, *dV = nullptr, *dUr = nullptr, *dVr = nullptr, *dP = nullptr;
    int64_t *rc = nullptr, *rp = nullptr;
    const size_t nK = (size_t)n_local * K_true, pK = (size_t)p * K_true;
    GK(cudaMalloc(&dU, nK * 8)); GK(cudaMalloc(&dUr, nK * 8)); GK(cudaMalloc(&dV, pK * 8)); GK(cudaMalloc(&dVr, pK * 8));
    GK(cudaMemcpy(dU, U, nK * 8, cudaMemcpyHostToDevice)); GK(cudaMemcpy(dV, V, pK * 8, cudaMemcpyHostToDevice));
    k_transpose_in<<<1024, 256>>>(n_local, K_true, dU, dUr);
    k_transpose_in<<<256, 256>>>(p, K_true, dV, dVr);
    if (prob1) { GK(cudaMalloc(&dP, (size_t)p * 8)); GK(cudaMemcpy(dP, prob1, (size_t)p * 8, cudaMemcpyHostToDevice)); }
    GK(cudaMalloc(&rc, (size_t)(n_local + 1) * 8)); GK(cudaMalloc(&rp, (size_t)(n_local + 1) * 8));
    GK(cudaMemset(rc, 0, (size_t)(n_local + 1) * 8));
    GenArgs a{n_local, row_offset, p, K_true, dUr, dVr, dP, (unsigned long long)seed};
    const int block = 256;
    const int grid = (int)std::min<int64_t>((n_local * 32 + block - 1) / block, 148 * 64);
    k_generate<false><<<grid, block>>>(a, rc, nullptr, nullptr, nullptr);
    size_t tmp_bytes = 0;
    GK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, rc, rp, n_local + 1));
    void *tmp = nullptr;
    GK(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    GK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rc, rp, n_local + 1));
    int64_t nnz = 0;
    GK(cudaMemcpy(&nnz, rp + n_local, 8, cudaMemcpyDeviceToHost));
    int32_t *ci = nullptr; float *vv = nullptr;
    GK(cudaMalloc(&ci, (size_t)std::max<int64_t>(nnz, 1) * 4)); GK(cudaMalloc(&vv, (size_t)std::max<int64_t>(nnz, 1) * 4));
    k_generate<true><<<grid, block>>>(a, nullptr, rp, ci, vv);
    GK(cudaDeviceSynchronize());
    GK(cudaGetLastError());
    cudaFree(tmp); cudaFree(rc); cudaFree(dU); cudaFree(dV); cudaFree(dUr); cudaFree(dVr); if (dP) cudaFree(dP);
    *rowptr_dev = rp; *colidx_dev = ci; *val_dev = vv; *nnz_out = nnz;
    return PCMF_OK;
}


// ---- per-gene statistics of a CSR matrix ---------------------------------------------------------------------
namespace {
__global__ void k_col_stats(int64_t n, const int64_t *rowptr, const int32_t *colidx, const float *val, double *s1, double *s2,
                            double *cn, double *mx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nw)
        for (int64_t e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
            const int j = colidx[e];
            const double x = (double)val[e];
            atomicAdd(s1 + j, x); atomicAdd(s2 + j, x * x); atomicAdd(cn + j, x != 0.0 ? 1.0 : 0.0);
            // counts are >= 0: the bit pattern of a non-negative double orders like the integer
            if (x > 0.0) atomicMax(reinterpret_cast<unsigned long long *>(mx + j), (unsigned long long)__double_as_longlong(x));
        }
}
__global__ void k_dfma_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}
}  // namespace

extern "C" int pcmf_csr_column_stats(int device, int64_t n_local, int32_t p, const int64_t *rowptr, const int32_t *colidx,
                                     const float *val_f32, int32_t on_device, double *colsum, double *colsumsq, double *colnnz,
                                     double *colmax, char *errbuf, size_t errlen) {
    if (!rowptr || !colidx || !val_f32 || !colsum || !colsumsq || !colnnz) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "bad argument to pcmf_csr_column_stats");
        return PCMF_EINVAL;
    }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "no CUDA device available");
        return PCMF_ECUDA;
    }
    GK(cudaSetDevice(device));
    const int64_t *rp = rowptr; const int32_t *ci = colidx; const float *vv = val_f32;
    int64_t *drp = nullptr; int32_t *dci = nullptr; float *dvv = nullptr;
    if (!on_device) {
        const int64_t nnz = rowptr[n_local];
        GK(cudaMalloc(&drp, (size_t)(n_local + 1) * 8)); GK(cudaMalloc(&dci, (size_t)std::max<int64_t>(nnz, 1) * 4));
        GK(cudaMalloc(&dvv, (size_t)std::max<int64_t>(nnz, 1) * 4));
        GK(cudaMemcpy(drp, rowptr, (size_t)(n_local + 1) * 8, cudaMemcpyHostToDevice));
        GK(cudaMemcpy(dci, colidx, (size_t)nnz * 4, cudaMemcpyHostToDevice));
        GK(cudaMemcpy(dvv, val_f32, (size_t)nnz * 4, cudaMemcpyHostToDevice));
        rp = drp; ci = dci; vv = dvv;
    }
    double *d = nullptr;
    GK(cudaMalloc(&d, (size_t)p * 4 * 8));
    GK(cudaMemset(d, 0, (size_t)p * 4 * 8));
    const int grid = (int)std::min<int64_t>((n_local * 32 + 255) / 256, 148 * 32);
    k_col_stats<<<grid, 256>>>(n_local, rp, ci, vv, d, d + p, d + 2 * (size_t)p, d + 3 * (size_t)p);
    GK(cudaMemcpy(colsum, d, (size_t)p * 8, cudaMemcpyDeviceToHost));
    GK(cudaMemcpy(colsumsq, d + p, (size_t)p * 8, cudaMemcpyDeviceToHost));
    GK(cudaMemcpy(colnnz, d + 2 * (size_t)p, (size_t)p * 8, cudaMemcpyDeviceToHost));
    if (colmax) GK(cudaMemcpy(colmax, d + 3 * (size_t)p, (size_t)p * 8, cudaMemcpyDeviceToHost));
    cudaFree(d); if (drp) cudaFree(drp); if (dci) cudaFree(dci); if (dvv) cudaFree(dvv);
    return PCMF_OK;
}

extern "C" int pcmf_measure_fp64_tflops(int device, double *tflops_out) {
    char *errbuf = nullptr; size_t errlen = 0;
    if (!tflops_out) return PCMF_EINVAL;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) return PCMF_ECUDA;
    GK(cudaSetDevice(device));
    cudaDeviceProp dp; GK(cudaGetDeviceProperties(&dp, device));
    double *out = nullptr; GK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1; GK(cudaEventCreate(&e0)); GK(cudaEventCreate(&e1));
    const int blocks = dp.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    k_dfma_peak<<<blocks, threads>>>(out, 1024);   // warm-up
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        GK(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(out, iters);
        GK(cudaEventRecord(e1));
        GK(cudaEventSynchronize(e1));
        float ms = 0; GK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops_out = best;
    return PCMF_OK;
}

// ---- roofline denominators measured on the device the bench runs on --------------------------------------------------
namespace {
__global__ void k_dmma_peak(double *out, int iters) {
    double d[8][2];
    for (int i = 0; i < 8; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; }
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
    if (s == 12345.678) out[0] = s;
}
// every thread gathers rows idx[e] of a (rows x 20) FP64 table with 256-bit loads, like the E-step non-zero passes
__global__ void k_gather_peak(const int32_t *__restrict__ idx, int64_t nidx, const double *__restrict__ tab, double *out) {
    double acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nidx; e += stride) {
        const double *row = tab + (int64_t)__ldg(idx + e) * 20;
#pragma unroll
        for (int k = 0; k < 5; ++k) {      // 256-bit loads, like load_vec in kernels.cuh
            double a, b, c, d;
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(row + 4 * k));
            acc = fma(a, b, acc); acc = fma(c, d, acc);
        }
    }
    if (acc == 12345.678) out[0] = acc;
}
__global__ void k_fill_idx(int32_t *idx, int64_t nidx, int32_t rows) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nidx; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t s = (uint64_t)e * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
        s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
        idx[e] = (int32_t)(s % (uint64_t)rows);
    }
}
}  // namespace

// FP64 tensor-pipe (DMMA m8n8k4) throughput in TFLOP/s: the roofline denominator of the DMMA zero-inflation passes.
extern "C" int pcmf_measure_fp64_mma_tflops(int device, double *tflops_out) {
    char *errbuf = nullptr; size_t errlen = 0;
    if (!tflops_out) return PCMF_EINVAL;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) return PCMF_ECUDA;
    GK(cudaSetDevice(device));
    cudaDeviceProp dp; GK(cudaGetDeviceProperties(&dp, device));
    double *out = nullptr; GK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1; GK(cudaEventCreate(&e0)); GK(cudaEventCreate(&e1));
    const int blocks = dp.multiProcessorCount * 4, threads = 256, iters = 1 << 13;
    k_dmma_peak<<<blocks, threads>>>(out, 256);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        GK(cudaEventRecord(e0));
        k_dmma_peak<<<blocks, threads>>>(out, iters);
        GK(cudaEventRecord(e1));
        GK(cudaEventSynchronize(e1));
        float ms = 0; GK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 512.0 * 8.0 * iters * ((double)blocks * threads / 32.0) / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops_out = best;
    return PCMF_OK;
}

// Gather bandwidth in GB/s of 160-byte K-vector rows out of a table of `rows` rows (L2-resident for the p x K gene
// table, larger than L2 for the n x K cell table): the ceiling of the E-step non-zero passes, which read one or two
// such rows per non-zero entry through L2.
extern "C" int pcmf_measure_gather_gbs(int device, int64_t rows, double *gbs_out) {
    char *errbuf = nullptr; size_t errlen = 0;
    if (!gbs_out || rows <= 0) return PCMF_EINVAL;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) return PCMF_ECUDA;
    GK(cudaSetDevice(device));
    cudaDeviceProp dp; GK(cudaGetDeviceProperties(&dp, device));
    const int64_t nidx = 100000000;
    int32_t *idx = nullptr; double *tab = nullptr, *out = nullptr;
    GK(cudaMalloc(&idx, nidx * 4)); GK(cudaMalloc(&tab, rows * 20 * 8)); GK(cudaMalloc(&out, 8));
    GK(cudaMemset(tab, 0, rows * 20 * 8));
    const int blocks = dp.multiProcessorCount * 16;
    k_fill_idx<<<blocks, 256>>>(idx, nidx, (int32_t)rows);
    cudaEvent_t e0, e1; GK(cudaEventCreate(&e0)); GK(cudaEventCreate(&e1));
    k_gather_peak<<<blocks, 256>>>(idx, nidx, tab, out);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        GK(cudaEventRecord(e0));
        k_gather_peak<<<blocks, 256>>>(idx, nidx, tab, out);
        GK(cudaEventRecord(e1));
        GK(cudaEventSynchronize(e1));
        float ms = 0; GK(cudaEventElapsedTime(&ms, e0, e1));
        const double gbs = (double)nidx * 160.0 / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(idx); cudaFree(tab); cudaFree(out);
    *gbs_out = best;
    return PCMF_OK;
}
