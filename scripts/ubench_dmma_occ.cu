// DMMA / DFMA throughput versus resident warps per SM and independent chains per warp (design micro-benchmark).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// NM independent DMMA chains + NF independent DFMA chains per thread, DEP = chain length before accumulators reset
template <int NF, int NM>
__global__ void k(double *out, int iters, double a, double b) {
    double f[NF > 0 ? NF : 1], m0[NM > 0 ? NM : 1], m1[NM > 0 ? NM : 1];
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = threadIdx.x * 1e-9 + i;
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { m0[i] = threadIdx.x * 1e-9; m1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(m0[i], m1[i], a + r, b + i);
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += m0[i] + m1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF, int NM>
void run(int warps_per_sm, double *out) {
    const int iters = 20000, grid = 148, block = warps_per_sm * 32;
    k<NF, NM><<<grid, block>>>(out, 10, 0.999, 1e-3);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k<NF, NM><<<grid, block>>>(out, iters, 0.999, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double thr = (double)grid * block;
    const double fl = thr * iters * 4.0 * NF * 2.0 + thr / 32.0 * iters * 4.0 * NM * 512.0;
    printf("NF=%2d NM=%2d warps/SM=%2d : %.2f TF\n", NF, NM, warps_per_sm, fl / ms * 1e-9);
}
int main() {
    double *out; CK(cudaMalloc(&out, 8 * 148 * 1024));
    for (int w : {4, 8, 12, 16, 32}) { run<0, 1>(w, out); run<0, 2>(w, out); run<0, 4>(w, out); run<0, 8>(w, out); run<0, 12>(w, out); }
    for (int w : {4, 8, 12, 16, 32}) { run<8, 0>(w, out); run<16, 0>(w, out); run<8, 4>(w, out); run<16, 4>(w, out); run<8, 12>(w, out); }
    CK(cudaDeviceSynchronize());
    return 0;
}
