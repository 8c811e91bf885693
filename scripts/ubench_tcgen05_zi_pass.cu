// Round-2 preparation, third step: what the cell-side zero-inflation pass costs in an FP32 mode on tcgen05.
// One CTA = 128 cells (TMEM lanes); gene tiles of 64 stream through shared memory (bulk copies, double buffered):
//   S = EU . EVS^T (kind::tf32, K = 32 padded factors) -> tcgen05.ld -> P = 1 / (1 + 2^((S - c) log2 e)) with ex2.approx / rcp.approx
//   -> tcgen05.st -> PV += P . EVS (A = P from TMEM, K = 64 genes per tile), PV accumulated in TMEM over all tiles.
// Not pipelined beyond the operand double buffer (an upper bound on the time), no occupancy words, column sums or entropy yet.
// Self-check: PV of a few cells against a double-precision host evaluation (TF32 operands: ~1e-3 relative).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/ubench_tcgen05_zi_pass scripts/ubench_tcgen05_zi_pass.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int KF = 32;                 // padded factors
constexpr int TG = 64;                 // genes per tile
// -DSPLIT3: 3xTF32 -- every FP32 operand x is split into hi = tf32(x) and lo = x - hi, and a product a.b is taken as
// a_hi b_hi + a_lo b_hi + a_hi b_lo (the contraction is three times as long; the tensor time stays negligible)
#ifdef SPLIT3
constexpr int NS = 3;
#else
constexpr int NS = 1;
#endif
constexpr int K1 = NS * KF;            // contraction length of the first product
constexpr int NBUF = NS == 3 ? 1 : 2;  // operand tile buffers (3xTF32: one, so that two CTAs fit an SM)
#ifndef DRAIN
#define DRAIN 4                        // tiles between two read-outs of the TMEM accumulator
#endif
#ifdef SPLIT3
constexpr int COL_S = 0, COL_P = 0, COL_PLO = 64, COL_PV = 128, TMEM_COLS = 256;     // P_hi over S in place
#else
// plain TF32: P overwrites S in place, 96 of 128 allocated columns -> four CTAs share the 512 TMEM columns of an SM
constexpr int COL_S = 0, COL_P = 0, COL_PLO = 0, COL_PV = 64, TMEM_COLS = 128;
#endif
constexpr int TILE_B1 = TG * K1 * 4, TILE_B2 = (NS == 3 ? 2 : 1) * KF * TG * 4, TILE_C = TG * 4, TILE_BYTES = TILE_B1 + TILE_B2 + TILE_C;
__host__ __device__ inline float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r);
#else
    return x;
#endif
}

__host__ __device__ inline uint32_t canon_off(int r, int k, int R) { return (k >> 2) * (R / 8) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// gene tiles in the order the kernel wants them: [B1: 64 genes x 32 factors, K-major canonical | B2: 32 factors x 64 genes | c: 64]
__global__ void k_pack(int p, const float *EVS, const float *c, unsigned char *tiles) {
    const int ntiles = (p + TG - 1) / TG;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)ntiles * TG * KF; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / (TG * KF)), r = (int)(e % (TG * KF)) / KF, k = (int)(e % KF);
        const int j = t * TG + r;
        const float v = j < p ? EVS[(int64_t)j * KF + k] : 0.f;
        unsigned char *base = tiles + (int64_t)t * TILE_BYTES;
        if (NS == 1) {
            *reinterpret_cast<float *>(base + canon_off(r, k, TG)) = v;                       // B1[gene r][factor k]
            *reinterpret_cast<float *>(base + TILE_B1 + canon_off(k, r, KF)) = v;             // B2[factor k][gene r]
        } else {
            const float hi = tf32_hi(v), lo = v - hi;
            // first product: A' = [a_hi | a_lo | a_hi] against B1' = [b_hi | b_hi | b_lo] along K
            *reinterpret_cast<float *>(base + canon_off(r, k, TG)) = hi;
            *reinterpret_cast<float *>(base + canon_off(r, KF + k, TG)) = hi;
            *reinterpret_cast<float *>(base + canon_off(r, 2 * KF + k, TG)) = lo;
            // second product: B2_hi and B2_lo as two [factor][gene] operands
            *reinterpret_cast<float *>(base + TILE_B1 + canon_off(k, r, KF)) = hi;
            *reinterpret_cast<float *>(base + TILE_B1 + KF * TG * 4 + canon_off(k, r, KF)) = lo;
        }
        if (k == 0) *reinterpret_cast<float *>(base + TILE_B1 + TILE_B2 + r * 4) = j < p ? c[j] : -1e30f;
    }
}

__global__ void __launch_bounds__(128) k_zi_fp32(int64_t n, int p, const float *EU, const unsigned char *tiles, float *PV) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA = smem;                                  // 128 x 32 floats, canonical
    unsigned char *sT = smem + 128 * K1 * 4;                   // 2 tiles
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t full[2], done_s, done_pv;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int64_t cell0 = (int64_t)blockIdx.x * 128;
    const int ntiles = (p + TG - 1) / TG;
    for (int e = tid; e < 128 * KF; e += 128) {
        const int r = e / KF, k = e % KF;
        const float x = cell0 + r < n ? EU[(cell0 + r) * KF + k] : 0.f;
        if (NS == 1) *reinterpret_cast<float *>(sA + canon_off(r, k, 128)) = x;
        else {
            const float hi = tf32_hi(x), lo = x - hi;
            *reinterpret_cast<float *>(sA + canon_off(r, k, 128)) = hi;
            *reinterpret_cast<float *>(sA + canon_off(r, KF + k, 128)) = lo;
            *reinterpret_cast<float *>(sA + canon_off(r, 2 * KF + k, 128)) = hi;
        }
    }
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&done_s, 1); mbar_init(&done_pv, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tmem_base, lane_addr = t0 + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc1 = make_idesc(128, TG), idesc2 = make_idesc(128, KF);
    const uint32_t lboA = (128 / 8) * 128, lboB1 = (TG / 8) * 128, lboB2 = (KF / 8) * 128;
    if (tid == 0) { bulk_load(sT, tiles, TILE_BYTES, &full[0]); if (NBUF > 1 && ntiles > 1) bulk_load(sT + TILE_BYTES, tiles + TILE_BYTES, TILE_BYTES, &full[1]); }
    double pvacc[KF];
#pragma unroll
    for (int c = 0; c < KF; ++c) pvacc[c] = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        unsigned char *buf = sT + (t % NBUF) * TILE_BYTES;
        if (tid == 0) {
            mbar_wait(&full[t % NBUF], (t / NBUF) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int s = 0; s < K1 / 8; ++s) {
                const uint64_t da = make_desc(smem_u32(sA) + 2 * s * lboA, lboA, 128), db = make_desc(smem_u32(buf) + 2 * s * lboB1, lboB1, 128);
                const uint32_t acc = s > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(t0 + COL_S), "l"(da), "l"(db), "r"(idesc1), "r"(acc));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_s)) : "memory");
        }
        mbar_wait(&done_s, t & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float *cs = reinterpret_cast<const float *>(buf + TILE_B1 + TILE_B2);
#pragma unroll
        for (int c0 = 0; c0 < TG; c0 += 16) {
            uint32_t v[16], w[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(lane_addr + COL_S + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float z = cs[c0 + c] - __uint_as_float(v[c]);                 // logit(prior) - lambda
                float e2, r;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(-1.4426950408889634f * z));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e2));
                if (NS == 1) v[c] = __float_as_uint(r);
                else { const float hi = tf32_hi(r); v[c] = __float_as_uint(hi); w[c] = __float_as_uint(r - hi); }
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(lane_addr + COL_P + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                           "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
            if (NS == 3)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             ::"r"(lane_addr + COL_PLO + c0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
                               "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int part = 0; part < NS; ++part) {          // P_hi E_hi [+ P_lo E_hi + P_hi E_lo]
                const uint32_t pcol = part == 1 ? COL_PLO : COL_P;
                const unsigned char *b2 = buf + TILE_B1 + (part == 2 ? KF * TG * 4 : 0);
                for (int s = 0; s < TG / 8; ++s) {
                    const uint64_t db = make_desc(smem_u32(b2) + 2 * s * lboB2, lboB2, 128);
                    const uint32_t acc = (t % DRAIN != 0) || (s > 0) || (part > 0);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(t0 + COL_PV), "r"(t0 + pcol + 8 * s), "l"(db), "r"(idesc2), "r"(acc));
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_pv)) : "memory");
            // the second product has to be through with this operand buffer (and with P) before either is overwritten
            mbar_wait(&done_pv, t & 1);
            if (t + NBUF < ntiles) bulk_load(buf, tiles + (int64_t)(t + NBUF) * TILE_BYTES, TILE_BYTES, &full[t % NBUF]);
        }
        __syncthreads();
        // The TMEM accumulator adds with truncation: left alone over all 313 tiles it drifts by ~2e-4 relative.  Every DRAIN
        // tiles it is therefore read out and added to FP64 registers, and the next tile starts a fresh accumulator.
        if (t % DRAIN == DRAIN - 1 || t == ntiles - 1) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < KF; c0 += 8) {
                uint32_t v[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(lane_addr + COL_PV + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 8; ++c) pvacc[c0 + c] += (double)__uint_as_float(v[c]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    }
    const int64_t row = cell0 + warp * 32 + (tid & 31);
    if (row < n) for (int c = 0; c < KF; ++c) PV[row * KF + c] = (float)pvacc[c];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "n"(TMEM_COLS));
}

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 148 * 128 * 8;
    const int p = argc > 2 ? atoi(argv[2]) : 20000, K = 20;
    const int ntiles = (p + TG - 1) / TG;
    std::vector<float> EU((size_t)n * KF, 0.f), EVS((size_t)p * KF, 0.f), c(p);
    srand(3);
    for (int64_t i = 0; i < n; ++i) for (int k = 0; k < K; ++k) EU[i * KF + k] = (float)(rand() % 1000) / 1000.f * 0.6f;
    for (int j = 0; j < p; ++j) { for (int k = 0; k < K; ++k) EVS[(size_t)j * KF + k] = (float)(rand() % 1000) / 1000.f * 0.5f; c[j] = (float)(rand() % 2000) / 1000.f - 1.f; }
    float *dEU, *dEVS, *dc, *dPV; unsigned char *dT;
    cudaMalloc(&dEU, EU.size() * 4); cudaMalloc(&dEVS, EVS.size() * 4); cudaMalloc(&dc, c.size() * 4); cudaMalloc(&dPV, EU.size() * 4);
    cudaMalloc(&dT, (size_t)ntiles * TILE_BYTES);
    cudaMemcpy(dEU, EU.data(), EU.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dEVS, EVS.data(), EVS.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
    k_pack<<<592, 256>>>(p, dEVS, dc, dT);
    const size_t smem = 128 * K1 * 4 + NBUF * TILE_BYTES;
    cudaFuncSetAttribute(k_zi_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)((n + 127) / 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_zi_fp32<<<grid, 128, smem>>>(n, p, dEU, dT, dPV);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k_zi_fp32<<<grid, 128, smem>>>(n, p, dEU, dT, dPV);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    std::vector<float> PV(EU.size());
    cudaMemcpy(PV.data(), dPV, PV.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int64_t i : {(int64_t)0, (int64_t)77, n / 2 + 5, n - 1}) {
        std::vector<double> ref(K, 0.0);
        for (int j = 0; j < p; ++j) {
            double lam = 0; for (int k = 0; k < K; ++k) lam += (double)EU[i * KF + k] * EVS[(size_t)j * KF + k];
            const double P = 1.0 / (1.0 + exp(-((double)c[j] - lam)));
            for (int k = 0; k < K; ++k) ref[k] += P * EVS[(size_t)j * KF + k];
        }
        for (int k = 0; k < K; ++k) worst = fmax(worst, fabs(PV[i * KF + k] - ref[k]) / fabs(ref[k]));
    }
    const double entries = (double)n * p;
    printf("FP32 cell-side ZI pass on tcgen05 (%s operands, ex2/rcp.approx): n=%lld p=%d K=%d: %.3f ms = %.1f Gentries/s -> %.1f ms for 1M x 20k; "
           "max rel err of PV on 4 cells %.2e\n", NS == 3 ? "3xTF32" : "TF32", (long long)n, p, K, ms, entries / ms / 1e6, ms * (1e6 * 20000.0) / entries, worst);
    return worst < 5e-3 ? 0 : 2;
}
