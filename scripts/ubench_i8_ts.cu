// Round-2 building block for the FP64-mode cell-side zero-inflation pass: the SECOND product PV_i += P_ij EVS_j as EXACT
// integer arithmetic on the 5th-generation tensor cores (tcgen05.mma kind::i8, u8 x u8 -> s32 in tensor memory).
//   P is the 32-bit fixed-point word the pass stores anyway (rint(P (2^32 - 1))): its four BYTES are four base-256 digits, and a
//   word written to tensor memory is read by the MMA as four consecutive K elements -- the A operand needs no repacking:
//       A[cell][4 j + s] = byte s of Pq[cell][gene j]
//   EVS is cut into five base-256 digits e_0..e_4 (most significant first) per factor column; products of equal weight
//   2^(8 (s - t)) share an accumulator: B_d[n][4 j + s] = e_(s - d)[j][n], d = s - t in {3, 2, 1, 0, -1} -> 5 MMAs per 8 genes.
// This program checks, bit for bit against the host: (i) tcgen05.st.16x256b.x1 takes the mma.sync m8n8 accumulator layout of
// two row tiles as it is, (ii) byte order / K-major no-swizzle 8-bit operand layout, (iii) u8 x u8 -> s32; and times the MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/ubench_i8_ts scripts/ubench_i8_ts.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, NF = 32, ND = 5, TJ = 32, NG = TJ / 8;          // cells, factor columns (N), digit groups, genes per tile
constexpr int REC = ND * NF * 32;                                      // bytes of B per 8-gene group
constexpr int COL_ACC = 0, COL_A = 160, TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);
}
// D = S32 (2 at bit 4), A = B = U8 (0 at bits 7 / 10), K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_i8(int m, int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc));
}

// Pq: [M][genes] words; Bq: per 8-gene group REC bytes in the canonical layout; out: [ND][M][NF] int32
// reps > 1 repeats the MMA part for timing (the result is then reps x the product, still exact while it fits)
__global__ void __launch_bounds__(128) k_i8(const uint32_t *Pq, const unsigned char *Bq, int genes, int32_t *out, int reps, long long *cycles) {
    extern __shared__ __align__(128) unsigned char sB[];                // all groups' B (test sizes are small)
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int ngroups = genes / 8, ntiles = genes / TJ;
    for (int e = tid; e < ngroups * REC / 16; e += 128) reinterpret_cast<uint4 *>(sB)[e] = reinterpret_cast<const uint4 *>(Bq)[e];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tmem_base, idesc = make_idesc_i8(M, NF);
    uint32_t phase = 0;
    long long c_mma = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
        // each warp: its 32 cells as four row tiles of 8 in the mma.sync accumulator layout (row 8 mt + g, genes 2q, 2q + 1)
        for (int grp = 0; grp < NG; ++grp) {
            const int j = tile * TJ + grp * 8 + 2 * q;
            uint32_t v[4][2];
            for (int mt = 0; mt < 4; ++mt) {
                const int row = w * 32 + mt * 8 + g;
                v[mt][0] = Pq[(size_t)row * genes + j];
                v[mt][1] = Pq[(size_t)row * genes + j + 1];
            }
            for (int h = 0; h < 2; ++h) {
                const uint32_t addr = t0 + ((uint32_t)(w * 32 + h * 16) << 16) + COL_A + grp * 8;
                asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(addr), "r"(v[2 * h][0]), "r"(v[2 * h][1]), "r"(v[2 * h + 1][0]), "r"(v[2 * h + 1][1]) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long c0 = clock64();
            for (int r = 0; r < reps; ++r)
                for (int grp = 0; grp < NG; ++grp) {
                    const uint32_t bbase = smem_u32(sB) + (tile * NG + grp) * REC;
                    for (int d = 0; d < ND; ++d)
                        mma_i8_ts(t0 + COL_ACC + d * NF, t0 + COL_A + grp * 8, make_desc(bbase + d * NF * 32, (NF / 8) * 128, 128), idesc,
                                  (tile > 0) || (grp > 0) || (r > 0));
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            mbar_wait(&bar, phase);
            c_mma += clock64() - c0;
        }
        phase ^= 1;
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0 && cycles) *cycles = c_mma;
    // read the accumulators: lane = cell, column = factor
    for (int d = 0; d < ND; ++d)
        for (int c0 = 0; c0 < NF; c0 += 16) {
            uint32_t v[16];
            const uint32_t addr = t0 + ((uint32_t)(w * 32) << 16) + COL_ACC + d * NF + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 16; ++c) out[((size_t)d * M + tid) * NF + c0 + c] = (int32_t)v[c];
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "n"(TMEM_COLS));
}

static uint32_t canon8(int n, int kb, int N) { return (uint32_t)((kb >> 4) * (N / 8) * 128 + (n >> 3) * 128 + (n & 7) * 16 + (kb & 15)); }

int main() {
    const int genes = 96, K = 20;
    std::vector<uint32_t> Pq((size_t)M * genes);
    srand(7);
    for (auto &x : Pq) {
        const int r = rand() % 10;
        x = r == 0 ? 0u : (r == 1 ? 0xFFFFFFFFu : ((uint32_t)rand() << 16) ^ (uint32_t)rand());
    }
    // EVS digits: e[t][j][n], t = 0 most significant
    std::vector<unsigned char> dig((size_t)ND * genes * NF, 0);
    for (int t = 0; t < ND; ++t)
        for (int j = 0; j < genes; ++j)
            for (int n = 0; n < K; ++n) dig[((size_t)t * genes + j) * NF + n] = (unsigned char)(rand() & 255);
    const int ngroups = genes / 8;
    std::vector<unsigned char> Bq((size_t)ngroups * REC, 0);
    const int dlist[ND] = {3, 2, 1, 0, -1};
    for (int grp = 0; grp < ngroups; ++grp)
        for (int di = 0; di < ND; ++di)
            for (int n = 0; n < NF; ++n)
                for (int jl = 0; jl < 8; ++jl)
                    for (int s = 0; s < 4; ++s) {
                        const int t = s - dlist[di];
                        unsigned char v = 0;
                        if (t >= 0 && t < ND) v = dig[((size_t)t * genes + grp * 8 + jl) * NF + n];
                        Bq[(size_t)grp * REC + di * NF * 32 + canon8(n, 4 * jl + s, NF)] = v;
                    }
    std::vector<int64_t> ref((size_t)ND * M * NF, 0);
    for (int di = 0; di < ND; ++di)
        for (int i = 0; i < M; ++i)
            for (int n = 0; n < NF; ++n) {
                int64_t a = 0;
                for (int j = 0; j < genes; ++j)
                    for (int s = 0; s < 4; ++s) {
                        const int t = s - dlist[di];
                        if (t < 0 || t >= ND) continue;
                        a += (int64_t)((Pq[(size_t)i * genes + j] >> (8 * s)) & 255u) * dig[((size_t)t * genes + j) * NF + n];
                    }
                ref[((size_t)di * M + i) * NF + n] = a;
            }
    uint32_t *dP; unsigned char *dB; int32_t *dO; long long *dC;
    cudaMalloc(&dP, Pq.size() * 4); cudaMalloc(&dB, Bq.size()); cudaMalloc(&dO, ref.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dP, Pq.data(), Pq.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bq.data(), Bq.size(), cudaMemcpyHostToDevice);
    const size_t smem = Bq.size();
    cudaFuncSetAttribute(k_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_i8<<<1, 128, smem>>>(dP, dB, genes, dO, 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> out(ref.size());
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < ref.size(); ++i)
        if ((int64_t)out[i] != ref[i]) {
            if (bad < 8) printf("mismatch at d=%zu cell=%zu n=%zu: got %d want %lld\n", i / (M * NF), (i / NF) % M, i % NF, out[i], (long long)ref[i]);
            ++bad;
        }
    printf("i8 TS product: %zu mismatches of %zu (bit-for-bit)\n", bad, ref.size());
    // timing: MMAs only, many repetitions on every SM
    const int reps = 2000;
    k_i8<<<148, 128, smem>>>(dP, dB, genes, dO, reps, dC);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error (timing): %s\n", cudaGetErrorString(e)); return 1; }
    long long cyc = 0;
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    const double n_mma = (double)reps * (genes / TJ) * NG * ND;
    printf("M128 N%d K32 i8 TS MMA: %.1f cycles each (148 CTAs, %d back to back incl. commit wait)\n", NF, (double)cyc / n_mma, reps * NG * ND);
    return bad != 0;
}
