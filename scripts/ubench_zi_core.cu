// Compute-only ceiling of the ZI tile loop (no global traffic, fragments re-read from a fixed shared buffer):
// separates "FP64 pipe + issue" limits from staging / synchronisation effects.  Design micro-benchmark.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../pcmf_b200/csrc/zi_mma.cuh"
using namespace pcmf;
#define CKX(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int KP, int MT, int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) k_core(double *out, int ngroups) {
    constexpr int KS = KP / 4, NT = (KP + 7) / 8, NG = 8;
    __shared__ double s_f1[NG * KS * 32], s_f2[NG * 2 * NT * 32], s_tab[64];
    const int lane = threadIdx.x & 31, q = lane & 3;
    for (int t = threadIdx.x; t < NG * KS * 32; t += blockDim.x) s_f1[t] = 1e-3 * (t % 17);
    for (int t = threadIdx.x; t < NG * 2 * NT * 32; t += blockDim.x) s_f2[t] = 1e-3 * (t % 13);
    load_pow2_tab(s_tab);
    __syncthreads();
    double a1[MT][KS], c[MT], acc[MT][NT][2];
    for (int mt = 0; mt < MT; ++mt) { c[mt] = 0.5 + mt; for (int s = 0; s < KS; ++s) a1[mt][s] = 1e-2 * (lane + s + mt); }
    for (int mt = 0; mt < MT; ++mt) for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0; acc[mt][nt][1] = 0; }
    uint32_t bits = 0x10101010u;
#pragma unroll 1
    for (int it = 0; it < ngroups; ++it) {
        const int grp = it & 7;
        double d[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { d[mt][0] = 0.0; d[mt][1] = 0.0; }
        if (MODE & 1) {
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = s_f1[(grp * KS + s) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
            }
        } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { d[mt][0] = s_f1[grp * 32 + lane] + mt; d[mt][1] = d[mt][0] * 0.5; }
        }
        if (MODE & 2) {
            const int sh0 = ((grp & 3) << 3) + 2 * q;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double y; bool big;
                d[mt][0] = fast_expit_tab(c[mt] - d[mt][0], s_tab, (bits >> sh0) & 1u, 1.0, y, big);
                d[mt][1] = fast_expit_tab(c[mt] - d[mt][1], s_tab, (bits >> (sh0 + 1)) & 1u, 1.0, y, big);
            }
        }
        if (MODE & 4) {
#pragma unroll
            for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = s_f2[((grp * 2 + sp) * NT + nt) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], d[mt][sp], b);
                }
        } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { acc[mt][0][0] += d[mt][0]; acc[mt][0][1] += d[mt][1]; }
        }
        bits = (bits << 1) | (bits >> 31);
    }
    double s = 0;
    for (int mt = 0; mt < MT; ++mt) for (int nt = 0; nt < NT; ++nt) s += acc[mt][nt][0] + acc[mt][nt][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MT, int MODE, int MINB>
void run(const char *name, double *out) {
    const int ngroups = 4000, grid = 148 * MINB;
    k_core<20, MT, MODE, MINB><<<grid, 128>>>(out, 10);
    cudaEvent_t e0, e1; CKX(cudaEventCreate(&e0)); CKX(cudaEventCreate(&e1));
    CKX(cudaEventRecord(e0));
    k_core<20, MT, MODE, MINB><<<grid, 128>>>(out, ngroups);
    CKX(cudaEventRecord(e1)); CKX(cudaEventSynchronize(e1));
    float ms; CKX(cudaEventElapsedTime(&ms, e0, e1));
    const double entries = (double)grid * 4 * ngroups * (64.0 * MT);
    printf("%-36s MT=%d CTAs/SM=%d: %.3f ms, %.1f Gentry/s  (C4 pass = %.1f ms)\n", name, MT, MINB, ms, entries / ms * 1e-6,
           2e10 / (entries / ms * 1e-6) * 1e-6);
}
int main() {
    double *out; CKX(cudaMalloc(&out, 8 * 148 * 8 * 128));
    run<4, 7, 3>("mma1 + expit + mma2", out);
    run<4, 7, 2>("mma1 + expit + mma2", out);
    run<4, 7, 1>("mma1 + expit + mma2", out);
    run<2, 7, 4>("mma1 + expit + mma2", out);
    run<4, 5, 3>("mma1 + mma2 (no expit)", out);
    run<4, 2, 3>("expit only", out);
    run<4, 3, 3>("mma1 + expit", out);
    run<4, 6, 3>("expit + mma2", out);
    CKX(cudaDeviceSynchronize());
    return 0;
}
