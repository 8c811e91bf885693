// Dense zero-inflation passes on the FP64 tensor path (mma.sync m8n8k4 DMMA -- FP64 has no tcgen05 kind).
//
// Both passes are FP64-pipe bound: per (cell, gene) entry two K-contractions and one expit.
// The thread-per-row kernels of kernels.cuh (k_zi_pass_cells / k_zi_pass_genes) keep every accumulator in the
// registers of ONE thread (2 x KP doubles + 2 x KP operands -> ~250 registers, 8 warps per SM, ~50 % FP64-pipe
// utilisation, latency bound).  Here the same arithmetic is laid out as two chained 8x8 tile products per warp,
// flash-attention style:
//      D1 (8 rows x 8 cols)  = A1 (rows x K)  . B1 (K x cols)        lambda_ij = EU_i . EVS_j
//      P                     = expit(c - D1)  elementwise in the accumulator registers
//      D2 (8 rows x K)      += P (rows x 8 cols) . B2 (cols x K)     PV_i += P_ij EVS_j   /  tmpMat_j += P_ij EU_i
// with the accumulators spread over the warp (2 doubles per thread per tile), so a thread needs ~8x fewer registers
// per entry, 3x more warps are resident and the DMMA / DFMA issue slots are kept busy.
//
// Fragment layouts of mma.sync.aligned.m8n8k4.row.col.f64 (PTX ISA): with g = lane >> 2, q = lane & 3
//      A (8x4):  a0 = A[g][q]          B (4x8):  b0 = B[q][g]          C/D (8x8):  d0, d1 = D[g][2q], D[g][2q+1]
// The second product consumes P straight from the D1 registers: k-step s' of the second product takes d_{s'} as its
// A fragment, which pairs A-column q with tile column 2q + s'; the B2 fragments are staged with the same pairing.
//
// Reference arithmetic: update_variational_zi_param (zi_gap_factor_model.cpp:346-367, zi_sparse...:389-410), the two
// GEMMs of the M-step (zi...cpp:336, :341; zi_sparse...:363, :369), update_prior_zi_param (zi...cpp:370-372) and the
// ZI terms of the ELBO (zi...cpp:180-212).
#pragma once
#include "kernels.cuh"

namespace pcmf {

__device__ const double kPow2Tab32[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,
    1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,
    1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,
    1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};

// exp(x) for |x| <= 700 with a 32-entry table of 2^(i/32) (shared memory): m = rint(32 x / ln 2), r = x - m ln2/32
// (two-term Cody-Waite, |r| <= ln2/64 = 0.0108), degree-6 Taylor (truncation 3.6e-18), scaled through the exponent
// field.  11 FP64 instructions instead of the 20 of fast_exp.
__device__ __forceinline__ double fast_exp_tab(double x, const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    double t = fma(x, 46.16624130844683, MAGIC);
    const int m = __double2loint(t);
    t -= MAGIC;
    double r = fma(t, -0.02166084939249829, x);
    r = fma(t, -7.247021293269686e-19, r);
    double p = 1.38888888888888888889e-03;                  // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);              // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);              // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);              // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= tab[m & 31];
    return __hiloint2double(__double2hiint(p) + ((m >> 5) << 20), __double2loint(p));
}
// internal::expit (internal.h:151-155), returns y = 1 + exp(-z) too (1 when clamped) and the (0,1)-interior flag
__device__ __forceinline__ double fast_expit_tab(double z, const double *__restrict__ tab, double &y_out, bool &interior) {
    const int hi = __double2hiint(z);
    const bool big = (hi & 0x7fffffff) >= 0x403E0000;     // |z| >= 30, also NaN / inf
    const double zc = big ? 0.0 : z;
    const double y = 1.0 + fast_exp_tab(-zc, tab);
    const double pr = fast_rcp(y);
    interior = !big;
    y_out = big ? 1.0 : y;
    return big ? ((hi < 0) ? 0.0 : 1.0) : pr;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Stage TR rows (multiple of 8) of a row-major [rows][KP] matrix into shared memory in fragment order:
//   B1 fragments (first product, contraction over the K factors):   f1[(grp * KS + s) * 32 + lane] = M[8 grp + g][4 s + q]
//   B2 fragments (second product, contraction over the 8 tile rows): f2[((grp * 2 + s') * NT + nt) * 32 + lane]
//                                                                       = M[8 grp + 2 q + s'][8 nt + g]   (0 beyond KP)
// Rows >= rows_valid_end read as 0.  The copies are asynchronous (cp.async, 8 bytes each, zero-filled when out of
// range) so the next tile is staged while the current one is being consumed.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int KP, int TR>
__device__ __forceinline__ void stage_b1(const double *__restrict__ M, int64_t row0, int64_t rows_valid_end, double *f1) {
    constexpr int KS = KP / 4;
    for (int t = threadIdx.x; t < TR * KP; t += blockDim.x) {
        const int row = t / KP, k = t - row * KP;
        const int64_t gr = row0 + row;
        const bool ok = gr < rows_valid_end;
        const int grp = row >> 3, g = row & 7, s = k >> 2, q = k & 3;
        cp_async8(f1 + (grp * KS + s) * 32 + g * 4 + q, ok ? M + gr * KP + k : M, ok);
    }
}
template <int KP, int TR>
__device__ __forceinline__ void stage_b2(const double *__restrict__ M, int64_t row0, int64_t rows_valid_end, double *f2) {
    constexpr int NT = (KP + 7) / 8;
    for (int t = threadIdx.x; t < TR * NT * 8; t += blockDim.x) {
        const int row = t / (NT * 8), k = t - row * (NT * 8);
        const int64_t gr = row0 + row;
        const bool ok = gr < rows_valid_end && k < KP;
        const int grp = row >> 3, rr = row & 7, q = rr >> 1, sp = rr & 1, nt = k >> 3, g = k & 7;
        cp_async8(f2 + ((grp * 2 + sp) * NT + nt) * 32 + g * 4 + q, ok ? M + gr * KP + k : M, ok);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pass A: rows = cells (8 MT per warp, fixed for the CTA), columns = genes streamed through shared memory.
// Outputs PV = prob_D . EVS (n x KP), colsum(prob_D), sum prob_D o UVt and the Bernoulli self term.
// ------------------------------------------------------------------------------------------------------------
template <int KP, int MT, int TJ>
__global__ void __launch_bounds__(128, (KP <= 20) ? (MT <= 2 ? 4 : 3) : 2) k_zi_cells_mma(ZiAArgs a) {
    constexpr int KS = KP / 4, NT = (KP + 7) / 8, NG = TJ / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int F1 = NG * KS * 32, F2 = NG * 2 * NT * 32, BUF = F1 + F2 + 2 * TJ;
    double *s_buf = reinterpret_cast<double *>(smem_raw);       // 2 x [f1 | f2 | c | nzv], double buffered
    double *s_cs = s_buf + 2 * BUF;                             // TJ: column sums of this CTA
    double *s_tab = s_cs + TJ;                                  // 32
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    if (threadIdx.x < 32) s_tab[threadIdx.x] = kPow2Tab32[threadIdx.x];
    const int64_t cell0 = (int64_t)blockIdx.x * (32 * MT) + w * (8 * MT);   // CTA = 4 warps x 8 MT cells
    double a1[MT][KS];
    bool live[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int64_t i = cell0 + mt * 8 + g;
        live[mt] = i < a.n;
#pragma unroll
        for (int s = 0; s < KS; ++s) a1[mt][s] = live[mt] ? __ldg(a.EU + i * KP + 4 * s + q) : 0.0;
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    double sum_pl = 0, ent_lin = 0, ent_log = 0, prod = 1.0;
    uint32_t bits[MT], bits_nx[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { bits[mt] = 0u; bits_nx[mt] = live[mt] ? __ldg(a.bitT + (cell0 + mt * 8 + g)) : 0u; }

    // tile staging: fragments by cp.async, per-gene constants by plain loads (TJ <= blockDim.x values per tile)
    auto stage = [&](int j0, double *buf) {
        stage_b1<KP, TJ>(a.EVS, j0, a.p, buf);
        stage_b2<KP, TJ>(a.EVS, j0, a.p, buf + F1);
        cp_async_commit();
        for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
            const int jj = j0 + t;
            double c = -1e300, nzv = 0.0;
            if (jj < a.p) {
                const int f = a.flag[jj];
                c = (f == 0) ? a.cz[jj] : ((f == 1) ? 1e300 : -1e300);     // zi...cpp:353-356 whole-gene shortcuts
                nzv = (f == 2) ? 0.0 : 1.0;                                // zi...cpp:359-360
            }
            buf[F1 + F2 + t] = c; buf[F1 + F2 + TJ + t] = nzv;
        }
    };
    for (int t = threadIdx.x; t < TJ; t += blockDim.x) s_cs[t] = 0.0;
    stage(0, s_buf);
    int cur = 0;
    for (int j0 = 0; j0 < a.p; j0 += TJ, cur ^= 1) {
        const double *s_f1 = s_buf + cur * BUF, *s_f2 = s_f1 + F1, *s_c = s_f2 + F2, *s_nzv = s_c + TJ;
        if (j0 + TJ < a.p) { stage(j0 + TJ, s_buf + (cur ^ 1) * BUF); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
#pragma unroll 1
        for (int grp = 0; grp < NG; ++grp) {
            const int jg = j0 + grp * 8;
            if (jg >= a.p) break;
            if ((grp & 3) == 0) {
                // occupancy words of this 32-gene block were requested one block ahead; request the next block now
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    bits[mt] = bits_nx[mt];
                    bits_nx[mt] = (live[mt] && jg + 32 < a.p) ? __ldg(a.bitT + (int64_t)((jg + 32) >> 5) * a.n + (cell0 + mt * 8 + g)) : 0u;
                }
            }
            double d[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { d[mt][0] = 0.0; d[mt][1] = 0.0; }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = s_f1[(grp * KS + s) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
            }
            const double2 c2 = *reinterpret_cast<const double2 *>(s_c + grp * 8 + 2 * q);
            const double2 n2 = *reinterpret_cast<const double2 *>(s_nzv + grp * 8 + 2 * q);
            const int sh0 = ((grp & 3) << 3) + 2 * q;
            double cs0 = 0.0, cs1 = 0.0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const double lam0 = d[mt][0], lam1 = d[mt][1];
                const double z0 = c2.x - lam0, z1 = c2.y - lam1;
                double y0, y1; bool in0, in1;
                double p0 = fast_expit_tab(z0, s_tab, y0, in0), p1 = fast_expit_tab(z1, s_tab, y1, in1);
                const bool nz0 = (bits[mt] >> sh0) & 1u, nz1 = (bits[mt] >> (sh0 + 1)) & 1u;
                p0 = nz0 ? n2.x : p0; p1 = nz1 ? n2.y : p1;
                p0 = live[mt] ? p0 : 0.0; p1 = live[mt] ? p1 : 0.0;
                // -bernoulli_loglike(prob_D, prob_D) over 0 < P < 1: sum (1-P)(-z) - log(1 + e^-z)
                const bool g0 = in0 && !nz0 && live[mt], g1 = in1 && !nz1 && live[mt];
                ent_lin = fma(g0 ? (1.0 - p0) : 0.0, -z0, ent_lin);
                ent_lin = fma(g1 ? (1.0 - p1) : 0.0, -z1, ent_lin);
                prod *= (g0 ? y0 : 1.0) * (g1 ? y1 : 1.0);
                sum_pl = fma(p0, lam0, sum_pl);                  // sum prob_D o UVt (zi...cpp:182)
                sum_pl = fma(p1, lam1, sum_pl);
                cs0 += p0; cs1 += p1;
                d[mt][0] = p0; d[mt][1] = p1;
            }
            if (prod > 1e100) { ent_log += log(prod); prod = 1.0; }     // (1 + e^30)^(2 MT) < 1e105 per group
#pragma unroll
            for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = s_f2[((grp * 2 + sp) * NT + nt) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], d[mt][sp], b);
                }
            // column sums: reduce over the 8 row groups g (lane bits 2..4)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
            }
            if (g == 0) {
                atomicAdd(s_cs + grp * 8 + 2 * q, cs0);
                atomicAdd(s_cs + grp * 8 + 2 * q + 1, cs1);
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
            if (j0 + t < a.p) atomicAdd(a.colsumP + j0 + t, s_cs[t]);
            s_cs[t] = 0.0;
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        if (!live[mt]) continue;
        const int64_t i = cell0 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int k = nt * 8 + 2 * q;
            if (k < KP) *reinterpret_cast<double2 *>(a.PV + i * KP + k) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
    }
    double ent = ent_lin - (ent_log + log(prod));
    sum_pl = block_sum(sum_pl, sh); ent = block_sum(ent, sh);
    if (threadIdx.x == 0) { atomicAdd(a.zsc + 0, sum_pl); atomicAdd(a.zsc + 1, ent); }
}

// ------------------------------------------------------------------------------------------------------------
// Pass B: rows = genes (8 MT per warp), columns = cells of this CTA's chunk streamed through shared memory.
// Output tmpMat = prob_D^T . EU_new (p x KP), prob_D recomputed from the triple saved when it was defined.
// ------------------------------------------------------------------------------------------------------------
template <int KP, int MT, int TI>
__global__ void __launch_bounds__(128, (KP <= 20) ? (MT <= 2 ? 4 : 3) : 2) k_zi_genes_mma(ZiBArgs a) {
    constexpr int KS = KP / 4, NT = (KP + 7) / 8, NG = TI / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int F1 = NG * KS * 32, F2 = NG * 2 * NT * 32, BUF = F1 + F2;
    double *s_buf = reinterpret_cast<double *>(smem_raw);       // 2 x [EUz first-product | EUn second-product fragments]
    double *s_tab = s_buf + 2 * BUF;                            // 32
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    if (threadIdx.x < 32) s_tab[threadIdx.x] = kPow2Tab32[threadIdx.x];
    const int gene0 = blockIdx.x * (32 * MT) + w * (8 * MT);
    double a1[MT][KS], c[MT], nzv[MT];
    bool live[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int j = gene0 + mt * 8 + g;
        live[mt] = j < a.p;
        c[mt] = -1e300; nzv[mt] = 0.0;
        if (live[mt]) {
            const int f = a.flag[j];
            c[mt] = (f == 0) ? a.cz[j] : ((f == 1) ? 1e300 : -1e300);
            nzv[mt] = (f == 2) ? 0.0 : 1.0;
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) a1[mt][s] = live[mt] ? __ldg(a.EVSz + (int64_t)j * KP + 4 * s + q) : 0.0;
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_chunk;
    const int64_t r1 = min(a.n, r0 + a.rows_per_chunk);
    uint32_t bits[MT], bits_nx[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        bits[mt] = 0u;
        bits_nx[mt] = (live[mt] && r0 < r1) ? __ldg(a.bitB + (r0 >> 5) * a.p + (gene0 + mt * 8 + g)) : 0u;
    }
    auto stage = [&](int64_t i0, double *buf) {
        stage_b1<KP, TI>(a.EUz, i0, r1, buf);
        stage_b2<KP, TI>(a.EUn, i0, r1, buf + F1);
        cp_async_commit();
    };
    if (r0 < r1) stage(r0, s_buf);
    int cur = 0;
    for (int64_t i0 = r0; i0 < r1; i0 += TI, cur ^= 1) {
        const double *s_f1 = s_buf + cur * BUF, *s_f2 = s_f1 + F1;
        if (i0 + TI < r1) { stage(i0 + TI, s_buf + (cur ^ 1) * BUF); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
#pragma unroll 1
        for (int grp = 0; grp < NG; ++grp) {
            const int64_t ig = i0 + grp * 8;
            if (ig >= r1) break;
            if ((grp & 3) == 0) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    bits[mt] = bits_nx[mt];
                    bits_nx[mt] = (live[mt] && ig + 32 < r1) ? __ldg(a.bitB + ((ig + 32) >> 5) * a.p + (gene0 + mt * 8 + g)) : 0u;
                }
            }
            double d[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { d[mt][0] = 0.0; d[mt][1] = 0.0; }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = s_f1[(grp * KS + s) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
            }
            const int sh0 = ((grp & 3) << 3) + 2 * q;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double y; bool in_;
                double p0 = fast_expit_tab(c[mt] - d[mt][0], s_tab, y, in_);
                double p1 = fast_expit_tab(c[mt] - d[mt][1], s_tab, y, in_);
                p0 = ((bits[mt] >> sh0) & 1u) ? nzv[mt] : p0;
                p1 = ((bits[mt] >> (sh0 + 1)) & 1u) ? nzv[mt] : p1;
                d[mt][0] = p0; d[mt][1] = p1;      // cells beyond the chunk have EU_new = 0 in the staged tile
            }
#pragma unroll
            for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = s_f2[((grp * 2 + sp) * NT + nt) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], d[mt][sp], b);
                }
        }
        __syncthreads();   // every warp is done with this buffer before it is staged again
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        if (!live[mt]) continue;
        const int j = gene0 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = nt * 8 + 2 * q + e;
                if (k < a.K) atomicAdd(a.tmpMat + (int64_t)j * KP + k, acc[mt][nt][e]);
            }
    }
}

}  // namespace pcmf
