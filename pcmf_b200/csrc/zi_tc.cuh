// FP32 mode (pcmf_options.precision = 1): the dense zero-inflation passes on the 5th-generation tensor cores.
//
// The two dense passes are the one place on the path where a K-contraction is dense: per (cell, gene) entry
//      lambda = EU_i . EVS_j,    P = expit(c_j - lambda)  (1 / 0 on the entries X != 0 / under the gene shortcuts),
//      pass A: PV_i += P EVS_j        (zi_gap_factor_model.cpp:336, :346-367)
//      pass B: tmpMat_j += P EU_new_i (zi_gap_factor_model.cpp:341)
// FP64 has no tcgen05 kind, so FP64 mode runs them as mma.sync DMMA tiles (zi_mma.cuh).  Here both products are
// tcgen05.mma kind::tf32 with FP32 accumulation in tensor memory, every FP32 operand x split into hi = tf32(x) and
// lo = x - hi and a product a.b taken as a_hi b_hi + a_lo b_hi + a_hi b_lo ("3xTF32": relative error 2^-22 per term instead of
// the 2^-11 of plain TF32 operands).  One CTA = 128 rows = the 128 lanes of tensor memory; the other side streams through
// shared memory in tiles of 64 records prepared once per pass by k_tc_pack (one cp.async.bulk per tile, mbarrier):
//      S (128 x 64)   = A' (128 x 3 KF1) . B1'^T                    A' = [a_lo | a_hi | a_hi], B1' = [b_hi | b_lo | b_hi]
//      tcgen05.ld -> P = expit(z) with ex2.approx / rcp.approx -> tcgen05.st (P_hi over S in place, P_lo beside it)
//      D (128 x KF2)  = P_lo . E_hi + P_hi . E_lo + P_hi . E_hi      with P as the A operand STRAIGHT FROM TENSOR MEMORY
// The tensor-memory accumulator adds with TRUNCATION, once per MMA instruction (measured in round 1: 2e-4 relative drift of an
// accumulator left alone over 313 tiles; first cut of this kernel: 1e-4 on b2 at K = 64).  Hence (i) the small cross terms
// come FIRST in both products, so that the running sum is small while they are added and only the KF1 / 8 resp. 8 steps of
// the hi.hi part truncate at the size of the result, and (ii) the accumulator of the second product is read out into FP64
// registers after EVERY tile and restarted.  Everything that is summed over cells or genes leaves the kernel in FP64.
// Stated accuracy of the mode: tests/test_gpu_fp32.py.
//
// The expit is evaluated on x = -z log2(e) without a clamp and without a select on the entry kind:
//      t = 2^-|x|, r = 1 / (1 + t), P = z > 0 ? r : t r      -- exactly 1 / 0 once |x| is large (ex2 underflows to 0)
//      entries with X != 0 take lambda = -1e29, whole-gene shortcuts c = +-1e30, rows / records beyond the matrix an offset of
//      1e30: all of them land on P = 1 or P = 0 through the same arithmetic (zi...cpp:353-360, internal.h:151-155)
//      -[P log P + (1-P) log(1-P)] = log(1 + t) + |z| t r   -- no cancellation, 0 for the saturated entries, so no mask either;
//      log(1 + t) through the running product of (1 + t) <= 2, one lg2 per 16 entries.
//
// Shared-memory operand layout: K-major, no swizzle -- core matrices of 8 rows x 16 bytes, SBO = 128 B between 8-row
// groups, LBO = (rows / 8) x 128 B between 16-byte K chunks (descriptor version 1); one MMA consumes K = 8 (two chunks).
#pragma once
#include "kernels.cuh"

namespace pcmf {

constexpr int TC_TG = 64;          // streamed records per tile

template <int KF1, int KF2>
struct TcCfg {
    static_assert(KF1 % 8 == 0 && KF2 % 16 == 0 && KF2 >= KF1, "tf32 MMA: K in steps of 8, N in steps of 16 at M = 128");
    static constexpr int K1 = 3 * KF1;                                    // contraction length of the first product
    static constexpr int A_BYTES = 128 * K1 * 4;                          // the CTA's own 128 rows, split three ways
    static constexpr int TILE_B1 = TC_TG * K1 * 4;                        // [64 records][K1]
    static constexpr int TILE_B2 = 2 * KF2 * TC_TG * 4;                   // hi and lo: [KF2 factors][64 records]
    static constexpr int TILE_C = TC_TG * 4;                              // per record: c2 = -c log2(e)
    static constexpr int TILE_BYTES = TILE_B1 + TILE_B2 + TILE_C;
    static constexpr int COL_P = 0, COL_PLO = 64, COL_ACC = 128, TMEM_COLS = 256;
    static_assert(COL_ACC + KF2 <= TMEM_COLS, "accumulator does not fit the allocation");
    static constexpr size_t SMEM = A_BYTES + TILE_BYTES + 1024;
};

__device__ __forceinline__ float tc_tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__host__ __device__ inline uint32_t tc_canon_off(int r, int k, int R) {
    return (uint32_t)((k >> 2) * (R / 8) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t tc_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tc_mbar_init(uint64_t *b, int c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void tc_mma_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
#define TC_LD16(v, addr)                                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"      \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                    \
                 : "r"(addr))
#define TC_ST16(addr, v)                                                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"      \
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), \
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")

// Tile records of the streamed side, written once per pass: record t holds 64 consecutive rows of the row-major FP64
// matrices M1 (first product) and M2 (second product), converted to FP32 and split:
//   [ B1: 64 x 3 KF1, K-major canonical: b_hi | b_lo | b_hi ][ B2 hi: KF2 x 64 ][ B2 lo: KF2 x 64 ][ c2: 64 ]
// c2 = -c log2(e), the record's additive part of x = -z log2(e).  Pass A (records = genes): c = effective logit of the gene,
// +-1e30 for the whole-gene shortcuts (zi...cpp:353-356), -1e30 beyond `rows` (P = 0).  Pass B (records = cells, cz == null):
// c = 0 for a real cell, -1e30 beyond `rows` (P = 0 there: the row sums and the second product must not see them).
template <int KF1, int KF2>
__global__ void k_tc_pack(int64_t rows, int KP, const double *__restrict__ M1, const double *__restrict__ M2,
                          const double *__restrict__ cz, const int32_t *__restrict__ flag, unsigned char *__restrict__ out) {
    using C = TcCfg<KF1, KF2>;
    const int64_t ntiles = (rows + TC_TG - 1) / TC_TG;
    const int64_t total = ntiles * TC_TG * KF2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / (TC_TG * KF2);
        const int rem = (int)(e - t * (TC_TG * KF2)), r = rem / KF2, k = rem - r * KF2;
        const int64_t row = t * TC_TG + r;
        const bool in = row < rows;
        unsigned char *base = out + t * C::TILE_BYTES;
        if (k < KF1) {
            const float v = (in && k < KP) ? (float)M1[row * KP + k] : 0.f;
            const float hi = tc_tf32_hi(v), lo = v - hi;
            *reinterpret_cast<float *>(base + tc_canon_off(r, k, TC_TG)) = hi;                 // x a_lo
            *reinterpret_cast<float *>(base + tc_canon_off(r, KF1 + k, TC_TG)) = lo;           // x a_hi
            *reinterpret_cast<float *>(base + tc_canon_off(r, 2 * KF1 + k, TC_TG)) = hi;       // x a_hi
        }
        {
            const float v = (in && k < KP) ? (float)M2[row * KP + k] : 0.f;
            const float hi = tc_tf32_hi(v), lo = v - hi;
            *reinterpret_cast<float *>(base + C::TILE_B1 + tc_canon_off(k, r, KF2)) = hi;
            *reinterpret_cast<float *>(base + C::TILE_B1 + KF2 * TC_TG * 4 + tc_canon_off(k, r, KF2)) = lo;
        }
        if (k == 0) {
            float c;
            if (cz) {
                const int f = in ? flag[row] : 2;
                c = (f == 0) ? (float)cz[row] : ((f == 1) ? 1e30f : -1e30f);
            } else c = in ? 0.f : -1e30f;
            *reinterpret_cast<float *>(base + C::TILE_B1 + C::TILE_B2 + r * 4) = -1.4426950408889634f * c;
        }
    }
}

struct ZiTcArgs {
    int64_t n; int32_t p; int K, KP;
    const double *A1;            // row-major FP64, KP wide: the CTA's rows of the first product (EU_new | EVS at definition)
    const unsigned char *tiles;  // k_tc_pack output for the streamed side
    int64_t rows, streamed;      // rows of this side (n | p), records of the streamed side (p | n)
    const uint32_t *bits;        // pass A: bitT [ceil(p/32)][n];  pass B: bitB [ceil(n/32)][p]
    // pass A
    const double *EU;            // = A1 (sum prob_D o UVt = sum EU . PV)
    double *PV, *colsumP, *zsc;
    // pass B
    const double *cz; const int32_t *flag;     // per gene (rows)
    double *tmpMat;
    int64_t tiles_per_chunk;     // pass B: blockIdx.y owns this many tiles of cells
};

// CELLS = true: pass A (rows = cells, genes stream);  false: pass B (rows = genes, cells stream).
template <int KF1, int KF2, bool CELLS>
__global__ void __launch_bounds__(128, (TcCfg<KF1, KF2>::SMEM <= 110 * 1024) ? 2 : 1) k_zi_tc(ZiTcArgs a) {
    using C = TcCfg<KF1, KF2>;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *sA = tc_smem;
    unsigned char *sT = tc_smem + C::A_BYTES;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar_full, bar_s, bar_acc;
    __shared__ float s_cs[4][TC_TG];
    __shared__ double sh[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * 128, row = row0 + tid;
    const bool live = row < a.rows;
    // the CTA's own rows: FP64 state -> FP32, split, canonical layout
    for (int e = tid; e < 128 * KF1; e += 128) {
        const int r = e / KF1, k = e - r * KF1;
        const float x = (row0 + r < a.rows && k < a.KP) ? (float)a.A1[(row0 + r) * a.KP + k] : 0.f;
        const float hi = tc_tf32_hi(x), lo = x - hi;
        *reinterpret_cast<float *>(sA + tc_canon_off(r, k, 128)) = lo;
        *reinterpret_cast<float *>(sA + tc_canon_off(r, KF1 + k, 128)) = hi;
        *reinterpret_cast<float *>(sA + tc_canon_off(r, 2 * KF1 + k, 128)) = hi;
    }
    if (tid == 0) {
        tc_mbar_init(&bar_full, 1); tc_mbar_init(&bar_s, 1); tc_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base)), "n"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tmem_base, lane_addr = t0 + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc1 = tc_idesc(128, TC_TG), idesc2 = tc_idesc(128, KF2);
    const uint32_t lboA = (128 / 8) * 128, lboB1 = (TC_TG / 8) * 128, lboB2 = (KF2 / 8) * 128;

    const int64_t ntiles_all = (a.streamed + TC_TG - 1) / TC_TG;
    const int64_t tile_first = CELLS ? 0 : (int64_t)blockIdx.y * a.tiles_per_chunk;
    const int ntiles = (int)max((int64_t)0, min(ntiles_all, tile_first + (CELLS ? ntiles_all : a.tiles_per_chunk)) - tile_first);
    // x = -z log2(e) = lambda log2(e) + c2(record) + xrow: the row's own part.  Pass A: 0, or 1e30 log2(e) for rows beyond n
    // (P = 0 everywhere).  Pass B: -c_j log2(e) of the row's gene (+-1e30 under the whole-gene shortcuts, zi...cpp:353-356).
    constexpr float L2E = 1.4426950408889634f;
    float xrow = 0.f;
    if (CELLS) xrow = live ? 0.f : 1e30f * L2E;
    else {
        float crow = -1e30f;
        if (live) {
            const int f = a.flag[row];
            crow = (f == 0) ? (float)a.cz[row] : ((f == 1) ? 1e30f : -1e30f);
        }
        xrow = -L2E * crow;
    }
    if (tid == 0 && ntiles > 0) tc_bulk_load(sT, a.tiles + tile_first * C::TILE_BYTES, C::TILE_BYTES, &bar_full);
    double acc[KF2];
#pragma unroll
    for (int c = 0; c < KF2; ++c) acc[c] = 0.0;
    double ent = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        const int64_t s0 = (tile_first + t) * TC_TG;              // first streamed record of the tile
        // occupancy words of this thread's row for the 64 streamed records (requested before the tensor core is waited for)
        uint32_t bw0 = 0u, bw1 = 0u;
        if (live) {
            const int64_t w0 = s0 >> 5;
            bw0 = __ldg(a.bits + w0 * a.rows + row);
            if (s0 + 32 < a.streamed) bw1 = __ldg(a.bits + (w0 + 1) * a.rows + row);
        }
        if (tid == 0) {
            tc_mbar_wait(&bar_full, t & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int s = 0; s < C::K1 / 8; ++s)
                tc_mma(t0 + C::COL_P, tc_desc(tc_smem_u32(sA) + 2 * s * lboA, lboA, 128), tc_desc(tc_smem_u32(sT) + 2 * s * lboB1, lboB1, 128),
                       idesc1, s > 0);
            tc_commit(&bar_s);
        }
        tc_mbar_wait(&bar_s, t & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float4 *cs4 = reinterpret_cast<const float4 *>(sT + C::TILE_B1 + C::TILE_B2);
        float h_lin = 0.f, h_log = 0.f, rsum = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float pk[32];                                        // this row's P for 32 streamed records (column sums, pass A)
            const uint32_t bw = half ? bw1 : bw0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c0 = half * 32 + q * 16;
                uint32_t v[16], w[16];
                TC_LD16(v, lane_addr + C::COL_P + c0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float prod = 1.0f;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 cc = cs4[(c0 >> 2) + c4];
                    const float c2v[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = c4 * 4 + e;
                        const bool nz = (bw >> (q * 16 + c)) & 1u;
                        const float lam = nz ? -1e29f : __uint_as_float(v[c]);      // X != 0: z = +huge (or stays -huge under the gene shortcut)
                        const float x = fmaf(lam, L2E, c2v[e]) + xrow;
                        float t, r;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-fabsf(x)));
                        const float y = 1.0f + t;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
                        const float tr = t * r;
                        const float pr = x < 0.f ? r : tr;
                        if (CELLS) {
                            h_lin = fmaf(fabsf(x), tr, h_lin);
                            prod *= y;
                            pk[q * 16 + c] = pr;
                        } else rsum += pr;
                        const float hi = tc_tf32_hi(pr);
                        v[c] = __float_as_uint(hi);
                        w[c] = __float_as_uint(pr - hi);
                    }
                }
                if (CELLS) {
                    float l2;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(prod));    // (1 + t)^16 <= 65536
                    h_log += l2;
                }
                TC_ST16(lane_addr + C::COL_P + c0, v);
                TC_ST16(lane_addr + C::COL_PLO + c0, w);
            }
            if (CELLS) {
                const float tot = warp_reduce_scatter32f(pk, lane);   // lane l: sum over the warp's 32 cells of record half * 32 + l
                s_cs[warp][half * 32 + lane] = tot;
            }
        }
        // sum P log P + (1-P) log(1-P) = -ln 2 (sum log2(1 + t) + sum |x| t r)
        if (CELLS) ent -= 0.6931471805599453 * ((double)h_lin + (double)h_log);
        (void)rsum;
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int part = 0; part < 3; ++part) {               // P_lo E_hi + P_hi E_lo + P_hi E_hi: small terms first
                const uint32_t pcol = part == 0 ? C::COL_PLO : C::COL_P;
                const unsigned char *b2 = sT + C::TILE_B1 + (part == 1 ? KF2 * TC_TG * 4 : 0);
#pragma unroll
                for (int s = 0; s < TC_TG / 8; ++s)
                    tc_mma_ta(t0 + C::COL_ACC, t0 + pcol + 8 * s, tc_desc(tc_smem_u32(b2) + 2 * s * lboB2, lboB2, 128), idesc2,
                              (s > 0) || (part > 0));
            }
            tc_commit(&bar_acc);
        }
        // column sums of prob_D over this CTA's 128 cells, one atomic per gene and tile (overlaps the second product)
        if (CELLS && tid < TC_TG) {
            const int64_t j = s0 + tid;
            if (j < a.streamed) atomicAdd(a.colsumP + j, (double)((s_cs[0][tid] + s_cs[1][tid]) + (s_cs[2][tid] + s_cs[3][tid])));
        }
        if (tid == 0) {
            // the second product has to be through with the operand tile (and with P) before either is overwritten
            tc_mbar_wait(&bar_acc, t & 1);
            if (t + 1 < ntiles) tc_bulk_load(sT, a.tiles + (tile_first + t + 1) * C::TILE_BYTES, C::TILE_BYTES, &bar_full);
        }
        __syncthreads();
        {   // read the tile's product out into FP64 (the accumulator truncates: it is not left to grow)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < KF2; c0 += 16) {
                uint32_t v[16];
                TC_LD16(v, lane_addr + C::COL_ACC + c0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c0 + c] += (double)__uint_as_float(v[c]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    }
    // epilogue
    if (CELLS) {
        double sum_pl = 0.0;
        if (live) {
#pragma unroll
            for (int k = 0; k < KF2; ++k)
                if (k < a.KP) {
                    a.PV[row * a.KP + k] = acc[k];
                    sum_pl = fma(a.EU[row * a.KP + k], acc[k], sum_pl);     // sum prob_D o UVt = sum_ik EU_ik PV_ik (zi...cpp:182)
                }
        }
        sum_pl = block_sum(sum_pl, sh); ent = block_sum(ent, sh);
        if (tid == 0) { atomicAdd(a.zsc + 0, sum_pl); atomicAdd(a.zsc + 1, ent); }
    } else if (live) {
#pragma unroll
        for (int k = 0; k < KF2; ++k)
            if (k < a.K) atomicAdd(a.tmpMat + row * a.KP + k, acc[k]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "n"(C::TMEM_COLS));
}

}  // namespace pcmf
