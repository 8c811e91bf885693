// Gather bandwidth of 160-byte rows with 128-bit (LDG.128) versus 256-bit (LDG.256, new on sm_100) loads.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int W>
__global__ void __launch_bounds__(256) k_gather(const int32_t *__restrict__ idx, int64_t nidx, const double *__restrict__ tab, double *out) {
    double acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nidx; e += stride) {
        const double *row = tab + (int64_t)__ldg(idx + e) * 20;
        if (W == 16) {
            const double2 *q = reinterpret_cast<const double2 *>(row);
#pragma unroll
            for (int k = 0; k < 10; ++k) { const double2 t = __ldg(q + k); acc = fma(t.x, t.y, acc); }
        } else {
#pragma unroll
            for (int k = 0; k < 5; ++k) { double a, b, c, d; ld256(row + 4 * k, a, b, c, d); acc = fma(a, b, acc); acc = fma(c, d, acc); }
        }
    }
    if (acc == 12345.678) out[0] = acc;
}
template <int W>
void run(int64_t rows, int grid) {
    const int64_t nidx = 200000000;
    std::vector<int32_t> h(nidx);
    uint64_t s = 88172645463325252ull;
    for (int64_t e = 0; e < nidx; ++e) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[e] = (int32_t)(s % (uint64_t)rows); }
    int32_t *di; double *tab, *out;
    CK(cudaMalloc(&di, nidx * 4)); CK(cudaMalloc(&tab, rows * 160)); CK(cudaMalloc(&out, 8));
    CK(cudaMemcpy(di, h.data(), nidx * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(tab, 0, rows * 160));
    k_gather<W><<<grid, 256>>>(di, nidx, tab, out);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_gather<W><<<grid, 256>>>(di, nidx, tab, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("load width %2d B, table rows %8lld, grid %5d: %.3f ms  gathered %.2f TB/s\n", W, (long long)rows, grid, ms, (double)nidx * 160 / ms * 1e-9);
    cudaFree(di); cudaFree(tab); cudaFree(out);
}
int main() {
    for (int grid : {148 * 8, 148 * 16}) { run<16>(20000, grid); run<32>(20000, grid); run<16>(1000000, grid); run<32>(1000000, grid); }
    return 0;
}
