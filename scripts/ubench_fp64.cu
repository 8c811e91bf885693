// Micro-benchmarks that decide the kernel design (not part of the product):
//   1. FP64 DFMA peak, FP64 DMMA (mma.sync m8n8k4) peak, and whether the two pipes overlap;
//   2. gather bandwidth of K-vector rows (160 B) from an L2-resident table (p x K) and from a table larger than L2
//      (n x K), as the E-step non-zero passes do;
//   3. gather bandwidth of the same rows from shared memory with 160 B and 176 B row strides.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, double a, double b) {
    double f[NF > 0 ? NF : 1];
    double m0[NM > 0 ? NM : 1], m1[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { m0[i] = threadIdx.x * 1e-9; m1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(m0[i], m1[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += m0[i] + m1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NM>
void run_mix(const char *name, double *out) {
    const int iters = 4000, grid = 148 * 8, block = 256;
    k_mix<NF, NM><<<grid, block>>>(out, 10, 0.999, 1e-3);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_mix<NF, NM><<<grid, block>>>(out, iters, 0.999, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double thr = (double)grid * block;
    const double fl_f = thr * iters * 4.0 * NF * 2.0;                 // DFMA: 2 flop per thread-instr
    const double fl_m = thr / 32.0 * iters * 4.0 * NM * (8 * 8 * 4 * 2.0);   // DMMA: 512 flop per warp-instr
    printf("%-28s NF=%d NM=%d  %.3f ms   DFMA %.2f TF  DMMA %.2f TF  total %.2f TF\n", name, NF, NM, ms, fl_f / ms * 1e-9,
           fl_m / ms * 1e-9, (fl_f + fl_m) / ms * 1e-9);
}

// ---- L2 / HBM gather: one warp per "cell", lanes gather rows idx[e] of a (rows x KP) double table -------------
template <int KP, typename T>
__global__ void __launch_bounds__(256) k_gather(const int32_t *__restrict__ idx, int64_t nidx, const T *__restrict__ tab,
                                                double *out) {
    double acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nidx; e += stride) {
        const int j = __ldg(idx + e);
        if (sizeof(T) == 8) {
            const double2 *q = reinterpret_cast<const double2 *>(tab + (int64_t)j * KP);
#pragma unroll
            for (int k = 0; k < KP / 2; ++k) { const double2 t = __ldg(q + k); acc = fma(t.x, t.y, acc); }
        } else {
            const float4 *q = reinterpret_cast<const float4 *>(tab + (int64_t)j * KP);
#pragma unroll
            for (int k = 0; k < KP / 4; ++k) { const float4 t = __ldg(q + k); acc = fma((double)t.x, (double)t.y, acc); acc = fma((double)t.z, (double)t.w, acc); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int KP, typename T>
void run_gather(const char *name, int64_t rows, int64_t nidx, int grid, double *out, bool sorted_runs) {
    std::vector<int32_t> h(nidx);
    uint64_t s = 88172645463325252ull;
    // emulate the cell pass: per warp-iteration 32 ascending random genes
    for (int64_t e = 0; e < nidx; ++e) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[e] = (int32_t)(s % (uint64_t)rows); }
    (void)sorted_runs;
    int32_t *di; T *tab;
    CK(cudaMalloc(&di, nidx * 4)); CK(cudaMalloc(&tab, rows * KP * sizeof(T)));
    CK(cudaMemcpy(di, h.data(), nidx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(tab, 0, rows * KP * sizeof(T)));
    k_gather<KP, T><<<grid, 256>>>(di, nidx, tab, out);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_gather<KP, T><<<grid, 256>>>(di, nidx, tab, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-28s rows=%lld KP=%d elt=%zu grid=%d: %.3f ms  %.2f Gnnz/s  gathered %.2f TB/s\n", name, (long long)rows, KP, sizeof(T), grid, ms,
           nidx / ms * 1e-6, (double)nidx * KP * sizeof(T) / ms * 1e-9);
    cudaFree(di); cudaFree(tab);
}

// ---- shared-memory gather: rows of KP doubles at stride LD doubles -------------------------------------------
template <int KP, int LD, int ROWS>
__global__ void __launch_bounds__(256) k_sgather(const int32_t *__restrict__ idx, int per_thread, double *out) {
    extern __shared__ double tab[];
    for (int t = threadIdx.x; t < ROWS * LD; t += blockDim.x) tab[t] = t * 1e-9;
    __syncthreads();
    double acc = 0;
    const int32_t *my = idx + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int it = 0; it < per_thread; ++it) {
        const int j = __ldg(my + it * stride) % ROWS;
        if (LD % 2 == 0) {
            const double2 *q = reinterpret_cast<const double2 *>(tab + j * LD);
#pragma unroll
            for (int k = 0; k < KP / 2; ++k) { const double2 t = q[k]; acc = fma(t.x, t.y, acc); }
        } else {
#pragma unroll
            for (int k = 0; k < KP; k += 2) acc = fma(tab[j * LD + k], tab[j * LD + k + 1], acc);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int KP, int LD, int ROWS>
void run_sgather(const char *name, double *out) {
    const int grid = 148, per_thread = 2000;
    const int64_t nidx = (int64_t)grid * 256 * per_thread;
    std::vector<int32_t> h(nidx);
    uint64_t s = 1234567;
    for (int64_t e = 0; e < nidx; ++e) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[e] = (int32_t)(s % 1000000); }
    int32_t *di; CK(cudaMalloc(&di, nidx * 4)); CK(cudaMemcpy(di, h.data(), nidx * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)ROWS * LD * 8;
    CK(cudaFuncSetAttribute(k_sgather<KP, LD, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sgather<KP, LD, ROWS><<<grid, 256, smem>>>(di, 10, out);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_sgather<KP, LD, ROWS><<<grid, 256, smem>>>(di, per_thread, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-28s KP=%d LD=%d rows=%d: %.3f ms  %.2f Gnnz/s  gathered %.2f TB/s\n", name, KP, LD, ROWS, ms, nidx / ms * 1e-6,
           (double)nidx * KP * 8 / ms * 1e-9);
    cudaFree(di);
}

int main() {
    double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 64 * 256));
    run_mix<8, 0>("dfma only", out);
    run_mix<16, 0>("dfma only", out);
    run_mix<0, 4>("dmma only", out);
    run_mix<0, 8>("dmma only", out);
    run_mix<8, 1>("mix", out);
    run_mix<8, 2>("mix", out);
    run_mix<8, 4>("mix", out);
    run_mix<16, 2>("mix", out);
    run_mix<4, 4>("mix", out);
    const int64_t nidx = 200000000;
    run_gather<20, double>("L2 table (p x K) f64", 20000, nidx, 148 * 8, out, true);
    run_gather<20, double>("L2 table (p x K) f64", 20000, nidx, 148 * 16, out, true);
    run_gather<20, double>("L2 table (p x K) f64", 20000, nidx, 148 * 32, out, true);
    run_gather<20, float>("L2 table (p x K) f32", 20000, nidx, 148 * 16, out, true);
    run_gather<20, double>("small table 1000 rows (L1)", 1000, nidx, 148 * 16, out, true);
    run_gather<20, double>("HBM table (n x K) f64", 1000000, nidx, 148 * 16, out, true);
    run_gather<20, double>("HBM table (n x K) f64", 1000000, nidx, 148 * 32, out, true);
    run_gather<20, float>("HBM table (n x K) f32", 1000000, nidx, 148 * 16, out, true);
    run_sgather<20, 20, 1024>("smem stride 160B", out);
    run_sgather<20, 22, 1024>("smem stride 176B", out);
    run_sgather<20, 21, 1024>("smem stride 168B (LDS.64)", out);
    CK(cudaDeviceSynchronize());
    return 0;
}
