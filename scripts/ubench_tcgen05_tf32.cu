// Round-2 preparation (FP32 mode of the cell-side zero-inflation pass): the smallest self-checking
// tcgen05 building block -- one CTA, D[128 x 64] (FP32, in TMEM) = A[128 x 32] . B[64 x 32]^T with kind::tf32, both operands
// K-major in shared memory WITHOUT swizzle (canonical 8-row x 16-byte core matrices), accumulator read back with tcgen05.ld.
// Inputs are small dyadic rationals, so the TF32 products and FP32 sums are exact and the check is bit-for-bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/ubench_tcgen05_tf32 scripts/ubench_tcgen05_tf32.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (r, k) of an R-row operand sits at
//   (k / 4) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 4) * 4   bytes,   SBO = 128, LBO = (R / 8) * 128
// (cute/atom/mma_traits_sm100.hpp: LayoutType::INTERLEAVE, ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint32_t canon_off(int r, int k, int R) { return (k >> 2) * (R / 8) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                  // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;        // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;        // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version 1 (Blackwell), bits [46,48)
    return d;                                                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) k_tf32(const float *A, const float *B, float *D) {
    __shared__ __align__(128) unsigned char sA[M * K * 4];
    __shared__ __align__(128) unsigned char sB[N * K * 4];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < M * K; e += 128) { const int r = e / K, k = e % K; *reinterpret_cast<float *>(sA + canon_off(r, k, M)) = A[e]; }
    for (int e = tid; e < N * K; e += 128) { const int r = e / K, k = e % K; *reinterpret_cast<float *>(sB + canon_off(r, k, N)) = B[e]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tacc = tmem_base;
    if (tid == 0) {
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format F32 = 1 at [4,6), a/b_format TF32 = 2 at
        // [7,10) / [10,13), a/b major K = 0 at 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lboA = (M / 8) * 128, lboB = (N / 8) * 128;
        for (int s = 0; s < K / 8; ++s) {                                  // one MMA per 8 TF32 of K (32 bytes = two 16-byte chunks)
            const uint64_t da = make_desc(smem_u32(sA) + 2 * s * lboA, lboA, 128);
            const uint64_t db = make_desc(smem_u32(sB) + 2 * s * lboB, lboB, 128);
            const uint32_t acc = s > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs (phase 0 of the barrier)
    asm volatile(
        "{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
        ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows of D), 8 columns at a time
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 8; ++c) D[row * N + c0 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tacc), "n"(64));
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N, -1.f), R(M * N);
    srand(1);
    for (auto &x : A) x = (float)(rand() % 17 - 8) / 8.0f;
    for (auto &x : B) x = (float)(rand() % 17 - 8) / 8.0f;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { float s = 0; for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k]; R[i * N + j] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    k_tf32<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double worst = 0;
    for (int e2 = 0; e2 < M * N; ++e2) { const double d = fabs((double)D[e2] - R[e2]); if (d > worst) worst = d; if (d != 0) ++bad; }
    printf("tcgen05 kind::tf32 128x64x32, K-major no-swizzle operands: %d of %d entries differ, max |diff| %.3g  (D[0]=%g ref %g, D[last]=%g ref %g)\n",
           bad, M * N, worst, D[0], R[0], D[M * N - 1], R[M * N - 1]);
    return bad ? 2 : 0;
}
