// The two non-zero passes of the E-step on the 2-D tiled copy of X (xtile.cuh): factor rows of both sides of a tile in
// shared memory, K split over groups of 4 lanes, one tile segment (a gene's cells / a cell's genes) per group.
// Same arithmetic and the same guards as k_gene_pass / k_estep_rows_q (kernels.cuh); reference lines cited there.
#pragma once
#include "xtile.cuh"
#include "zi_mma.cuh"     // mbarrier + bulk-copy helpers

namespace pcmf {

__device__ __forceinline__ double grp_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ int warp_max_int_groups(int v) {      // value is uniform inside a group of 4 lanes
    v = max(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = max(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = max(v, __shfl_xor_sync(0xffffffffu, v, 16));
    return v;
}
// Which factors a lane of a 4-lane group holds.  VEC: rows of doubles read 16 bytes at a time -- slot i < 2A of lane l is
// factor 8 (i / 2) + 2 l + (i & 1), the tail slot (KP = 8 A + 4) factor 8 A + l.  PAIR: rows of (eu, eu ElogU) pairs, one
// pair = 16 bytes per slot -- slot i is factor l + 4 i.  Either way a group reads 64 contiguous bytes per instruction
// (LDS.128): a quarter-warp phase then serves two groups instead of the four 32-byte pieces of 8-byte reads.
template <int KP, bool PAIR>
struct FMap {
    static constexpr int A = KP / 8, B = (KP % 8) / 4, F = KP / 4;
    __device__ __forceinline__ static int k(int i, int l) {
        if (PAIR) return l + 4 * i;
        return i < 2 * A ? 8 * (i >> 1) + 2 * l + (i & 1) : 8 * A + l;
    }
};
template <int KP>
__device__ __forceinline__ void load_row_vec(const double *__restrict__ row, int l, double (&u)[KP / 4]) {
    using M = FMap<KP, false>;
#pragma unroll
    for (int s = 0; s < M::A; ++s) {
        const double2 t = *reinterpret_cast<const double2 *>(row + 8 * s + 2 * l);
        u[2 * s] = t.x; u[2 * s + 1] = t.y;
    }
    if (M::B) u[2 * M::A] = row[8 * M::A + l];
}
// The reference's guarded multinomial terms of ONE entry (gap_factor_model.cpp:481-487) by a group of 4 lanes:
// e[i] = S_jk cexp(ElogU_ik + ElogV_jk - m), returns the (shifted) sum on the 4 lanes.
// Called by the whole group together (the condition that leads here is uniform in the group).
template <int KP, bool PAIR>
__device__ __noinline__ double exact_terms_group(const double *__restrict__ elu_row, const double *__restrict__ elv_row,
                                                 const double *__restrict__ sd_row, int K, int l, unsigned gmask,
                                                 double (&e)[KP / 4], double (&elu)[KP / 4]) {
    using M = FMap<KP, PAIR>;
    double m = -1e300;
#pragma unroll
    for (int i = 0; i < M::F; ++i) {
        const int k = M::k(i, l);
        elu[i] = elu_row[k];
        e[i] = elu[i] + elv_row[k];
        if (k < K) m = fmax(m, e[i]);
    }
    m = fmax(m, __shfl_xor_sync(gmask, m, 1));
    m = fmax(m, __shfl_xor_sync(gmask, m, 2));
    if (m < 100.0) m = 0.0;
    double T = 0.0;
#pragma unroll
    for (int i = 0; i < M::F; ++i) {
        const int k = M::k(i, l);
        e[i] = (k < K) ? custom_exp(e[i] - m) * sd_row[k] : 0.0;
        T += e[i];
    }
    T += __shfl_xor_sync(gmask, T, 1);
    T += __shfl_xor_sync(gmask, T, 2);
    return T;
}

// ------------------------------------------------------------------------------------------------------------
// Gene side: B1(j,k) = sum_i q_ij e_ijk [, B2(j,k) = sum_i q_ij e_ijk ElogU_ik | xlogT(j) = sum_i x_ij log T_ij], q left in
// cell order for the cell pass.  CTA = one block of 256 genes x a range of cell blocks; the cell rows (eu, or the pairs) of
// a tile arrive by bulk copy, double buffered; gene rows and the gene sums stay in shared memory for the whole CTA.
// The 32 rounds of a tile (8 segments each, longest first) are dealt to the 16 warps as (w, 31 - w): a long and a short one.
// ------------------------------------------------------------------------------------------------------------
template <int KP, bool WITH_B2, bool WITH_LOGT>
struct TileGeneCfg {
    static constexpr int NB = WITH_B2 ? 2 : 1;
    static constexpr int STAGE = NB * XT_TI * KP;                                   // doubles
    static constexpr size_t SMEM = sizeof(double) * ((size_t)(1 + NB) * XT_TJ * KP + (WITH_LOGT ? XT_TJ : 0) + 2 * STAGE);
};
template <int KP, bool WITH_B2, bool WITH_LOGT>
__global__ void __launch_bounds__(512, 1) k_tile_genes(TileGeneArgs A) {
    using C = TileGeneCfg<KP, WITH_B2, WITH_LOGT>;
    using M = FMap<KP, WITH_B2>;
    constexpr int TI = XT_TI, TJ = XT_TJ, F = KP / 4, ROUNDS = TJ / 8, STAGE = C::STAGE;
    static_assert(ROUNDS == 32, "two rounds per warp");
    extern __shared__ __align__(128) unsigned char xt_smem[];
    double *ev_s = reinterpret_cast<double *>(xt_smem);
    double *acc1 = ev_s + TJ * KP;
    double *acc2 = acc1 + TJ * KP;                                   // WITH_B2 only
    double *ltacc = acc1 + C::NB * TJ * KP;                          // WITH_LOGT only
    double *stage = ltacc + (WITH_LOGT ? TJ : 0);
    __shared__ uint64_t full[2];
    __shared__ uint8_t gflag[TJ];        // 0: product formula, 1: skip (sums copied / beyond p), 2: guarded formula, 3: every factor masked
    __shared__ uint16_t s_ofs[2][TJ + 2];   // segment tables of the current / next tile
    __shared__ uint8_t s_perm[2][TJ];
    __shared__ int64_t s_tb[2];
    const ColArgs &a = A.c;
    const XTile &x = A.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l = lane & 3, grp = lane >> 2;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int J = blockIdx.x, j0 = J * TJ;
    const int ib0 = (int)blockIdx.y * A.blocks_per_chunk, ib1 = min(x.nI, ib0 + A.blocks_per_chunk);
    const double gminU = dbl_from_ordered(a.guard[G_MINU]), gmaxU = dbl_from_ordered(a.guard[G_MAXU]);
    for (int e = tid; e < TJ * KP; e += blockDim.x) {
        const int c = e / KP;
        ev_s[e] = (j0 + c < a.p) ? a.ev[(int64_t)j0 * KP + e] : 0.0;
        acc1[e] = 0.0;
        if (WITH_B2) acc2[e] = 0.0;
    }
    int work = 0;
    for (int c = tid; c < TJ; c += blockDim.x) {
        const int j = j0 + c;
        uint8_t f = 1;
        if (j < a.p && !(a.reuse && a.reuse[j])) {
            if (a.allmasked[j]) f = 3;
            else {
                const double2 ext = a.colext[j];
                f = path_flags(ext.x, ext.y, gminU, gmaxU).force_exact ? 2 : 0;
            }
        }
        gflag[c] = f;
        work |= (f != 1);
        if (WITH_LOGT) ltacc[c] = 0.0;
    }
    // no gene of the block needs anything (all copied): nothing to stream
    const bool idle = !__syncthreads_or(work);
    // same statistics, same mask row as the spike-and-slab pass of the previous iteration (sparse...cpp:448-451 vs :350-356):
    // the sums are identical, copy them (first chunk only; nobody adds to these genes)
    if (a.reuse && blockIdx.y == 0)
        for (int e = tid; e < TJ * KP; e += blockDim.x) {
            const int j = j0 + e / KP;
            if (j < a.p && a.reuse[j]) {
                a.B1[(int64_t)j0 * KP + e] = a.B1_src[(int64_t)j0 * KP + e];
                if (WITH_B2) a.B2[(int64_t)j0 * KP + e] = a.B2_src[(int64_t)j0 * KP + e];
            }
        }
    if (idle) return;
    const double *cellrows = WITH_B2 ? reinterpret_cast<const double *>(A.pairs) : a.eu;
    auto issue = [&](int ib, int s) {
        const int rows = (int)min((int64_t)TI, A.n - (int64_t)ib * TI);
        const unsigned bytes = (unsigned)(rows * KP * sizeof(double) * C::NB);
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(stage + s * STAGE, cellrows + (int64_t)ib * TI * KP * C::NB, bytes, &full[s]);
    };
    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init();
        if (ib0 < ib1) issue(ib0, 0);
    }
    if (ib0 < ib1) {
        const int64_t t0 = (int64_t)ib0 * x.nJ + J;
        if (tid <= TJ) s_ofs[0][tid] = x.gofs[t0 * (TJ + 1) + tid];
        if (tid < TJ) s_perm[0][tid] = x.gperm[t0 * TJ + tid];
        if (tid == 0) s_tb[0] = x.tptr[t0];
    }
    __syncthreads();
    for (int ib = ib0; ib < ib1; ++ib) {
        const int it = ib - ib0, s = it & 1;
        if (tid == 0 && ib + 1 < ib1) issue(ib + 1, s ^ 1);
        // the next tile's segment tables travel through registers under this tile's work; its records are pulled into L2
        uint16_t nofs = 0; uint8_t nperm = 0; int64_t ntb = 0;
        if (ib + 1 < ib1) {
            const int64_t tn = (int64_t)(ib + 1) * x.nJ + J;
            if (tid <= TJ) nofs = x.gofs[tn * (TJ + 1) + tid];
            if (tid < TJ) nperm = x.gperm[tn * TJ + tid];
            if (tid == 32) {
                ntb = x.tptr[tn];
                const int64_t nte = x.tptr[tn + 1];
                if (nte > ntb) {
                    const uintptr_t p0 = reinterpret_cast<uintptr_t>(x.grec + ntb) & ~(uintptr_t)15;
                    const uintptr_t p1 = (reinterpret_cast<uintptr_t>(x.grec + nte) + 15) & ~(uintptr_t)15;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"((unsigned)(p1 - p0)) : "memory");
                }
            }
        }
        mbar_wait(&full[s], (it >> 1) & 1);
        const double *su = stage + s * STAGE;
        const int64_t tbase = s_tb[s];
        const uint16_t *ofs = s_ofs[s];
        const uint8_t *perm = s_perm[s];
        const uint2 *rec = x.grec + tbase;
        double *qout = x.q + tbase;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int r = half ? (ROUNDS - 1 - warp) : warp;
            const int c = perm[r * 8 + grp];
            const int beg = ofs[c];
            const int flag = gflag[c];
            const int len = (flag == 1) ? 0 : (int)ofs[c + 1] - beg;
            const int maxlen = warp_max_int_groups(len);
            if (maxlen == 0) continue;
            const bool normal = flag == 0;
            double v[F], b1[F], b2[WITH_B2 ? F : 1];
#pragma unroll
            for (int i = 0; i < F; ++i) { v[i] = ev_s[c * KP + M::k(i, l)]; b1[i] = 0.0; if (WITH_B2) b2[i] = 0.0; }
            double lt = 0.0;
            // a group reads its segment 4 records at a time (one per lane), two such loads ahead of the one being consumed
            uint2 nx = (l < len) ? __ldg(rec + beg + l) : make_uint2(0u, 0u);
            uint2 nx2 = (4 + l < len) ? __ldg(rec + beg + 4 + l) : make_uint2(0u, 0u);
            for (int base = 0; base < maxlen; base += 4) {
                const uint2 cur = nx;
                nx = nx2;
                if (base + 8 < maxlen) nx2 = (base + 8 + l < len) ? __ldg(rec + beg + base + 8 + l) : make_uint2(0u, 0u);
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    if (base + u4 >= maxlen) break;                          // warp-uniform
                    const unsigned w0 = __shfl_sync(0xffffffffu, cur.x, u4, 4), w1 = __shfl_sync(0xffffffffu, cur.y, u4, 4);
                    const int rr = (int)(w0 & 127u), pos = (int)(w0 >> 8);
                    const double xv = (double)__uint_as_float(w1);
                    double uu[F], ul[WITH_B2 ? F : 1];
                    if (WITH_B2) {
#pragma unroll
                        for (int i = 0; i < F; ++i) {
                            const double2 t2 = *reinterpret_cast<const double2 *>(su + 2 * (rr * KP + l + 4 * i));
                            uu[i] = t2.x; ul[i] = t2.y;
                        }
                    } else load_row_vec<KP>(su + rr * KP, l, reinterpret_cast<double (&)[KP / 4]>(uu));
                    double T = 0.0;
#pragma unroll
                    for (int i = 0; i < F; ++i) T = fma(uu[i], v[i], T);
                    T = grp_sum(T);
                    const bool act = base + u4 < len;
                    const bool ok = act && normal && T >= 1e-26;
                    const double q = ok ? xv * fast_rcp(T) : 0.0;
                    if (act && l == 0) qout[pos] = ok ? q : (flag == 3 ? 0.0 : -xv);   // 0: masked gene; -x: guarded formula
#pragma unroll
                    for (int i = 0; i < F; ++i) { b1[i] = fma(q, uu[i], b1[i]); if (WITH_B2) b2[i] = fma(q, ul[i], b2[i]); }
                    if (WITH_LOGT && ok && l == 0) lt += xv * log(T);
                    if (act && !ok && flag != 3) {
                        double e[F], elu[F];
                        const int64_t ig = (int64_t)ib * TI + rr, jg = j0 + c;
                        const double Tx = exact_terms_group<KP, WITH_B2>(a.ElogU + ig * KP, a.ElogV + jg * KP, a.Sd + jg * KP, a.K, l, gmask, e, elu);
                        const double qx = xv / (Tx > 0.0 ? Tx : 1.0);
#pragma unroll
                        for (int i = 0; i < F; ++i) {
                            acc1[c * KP + M::k(i, l)] += qx * e[i];
                            if (WITH_B2) acc2[c * KP + M::k(i, l)] += qx * e[i] * elu[i];
                        }
                        if (l == 0) {
                            if (WITH_LOGT) lt += xv * log(Tx);
                            atomicAdd(a.dbg + 1, 1ull);
                        }
                    }
                }
            }
            if (len > 0 && flag != 3) {
#pragma unroll
                for (int i = 0; i < F; ++i) {
                    acc1[c * KP + M::k(i, l)] += b1[i] * v[i];                 // e_ijk = eu_ik ev_jk
                    if (WITH_B2) acc2[c * KP + M::k(i, l)] += b2[i] * v[i];
                }
                if (WITH_LOGT && l == 0) ltacc[c] += lt;
            }
        }
        if (ib + 1 < ib1) {
            if (tid <= TJ) s_ofs[s ^ 1][tid] = nofs;
            if (tid < TJ) s_perm[s ^ 1][tid] = nperm;
            if (tid == 32) s_tb[s ^ 1] = ntb;
        }
        __syncthreads();      // the tile's buffer may be refilled, its genes taken by other groups
    }
    for (int e = tid; e < TJ * KP; e += blockDim.x) {
        const int c = e / KP;
        if (j0 + c < a.p && gflag[c] != 1) {
            if (acc1[e] != 0.0) atomicAdd(a.B1 + (int64_t)j0 * KP + e, acc1[e]);
            if (WITH_B2 && acc2[e] != 0.0) atomicAdd(a.B2 + (int64_t)j0 * KP + e, acc2[e]);
        }
    }
    if (WITH_LOGT)
        for (int c = tid; c < TJ; c += blockDim.x)
            if (j0 + c < a.p && gflag[c] != 1 && ltacc[c] != 0.0) atomicAdd(a.xlogT + j0 + c, ltacc[c]);
}

// ------------------------------------------------------------------------------------------------------------
// Cell side + fused U-side M-step: EZ_j(i,.) = eu_i o sum_j q_ij evw_j from the q the gene pass left, then exactly the
// epilogue of k_estep_rows_q.  CTA = 256 cells (two cell blocks); the evw rows of a gene block arrive by bulk copy,
// three buffers; the cells' sums stay in shared memory.  Warp w takes round w of the first cell block and round
// 15 - w of the second (rounds are sorted by length).
// ------------------------------------------------------------------------------------------------------------
template <int KP>
struct TileCellCfg {
    static constexpr int RB = 2, ROWS = RB * XT_TI, NS = 3, STAGE = XT_TJ * KP;
    static constexpr size_t SMEM = sizeof(double) * ((size_t)2 * ROWS * KP + NS * STAGE);
};
template <int KP>
__global__ void __launch_bounds__(512, 1) k_tile_cells(TileCellArgs A) {
    using C = TileCellCfg<KP>;
    using M = FMap<KP, false>;
    constexpr int TI = XT_TI, TJ = XT_TJ, RB = C::RB, ROWS = C::ROWS, F = KP / 4, RPT = TI / 8, NS = C::NS, STAGE = C::STAGE;
    static_assert(RPT == 16 && RB == 2, "one round per warp and cell block");
    extern __shared__ __align__(128) unsigned char xt_smem[];
    double *acc = reinterpret_cast<double *>(xt_smem);
    double *eu_s = acc + ROWS * KP;
    double *stage = eu_s + ROWS * KP;
    __shared__ uint64_t full[NS];
    __shared__ double sh[32];
    __shared__ double scs[2][64];
    __shared__ uint16_t s_ofs[2][RB][TI + 2];   // segment tables of the current / next gene block, per cell block
    __shared__ uint8_t s_perm[2][RB][TI];
    __shared__ int64_t s_tb[2][RB];
    const RowArgs &a = A.r;
    const XTile &x = A.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l = lane & 3, grp = lane >> 2;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int64_t i0 = (int64_t)blockIdx.x * ROWS;
    const int rows_here = (int)min((int64_t)ROWS, a.n - i0);
    const int nJ = x.nJ;
    for (int e = tid; e < ROWS * KP; e += blockDim.x) {
        acc[e] = 0.0;
        eu_s[e] = (e < rows_here * KP) ? a.eu[i0 * KP + e] : 0.0;
    }
    if (tid < 128) scs[tid >> 6][tid & 63] = 0.0;
    auto issue = [&](int Jt, int s) {
        const int rows = min(TJ, A.p - Jt * TJ);
        const unsigned bytes = (unsigned)(rows * KP * sizeof(double));
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(stage + s * STAGE, a.evw + (int64_t)Jt * TJ * KP, bytes, &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        issue(0, 0);
        if (nJ > 1) issue(1, 1);
    }
    // thread -> (cell block, table slot) for the segment tables
    const int trb = tid / 256, tsl = tid - trb * 256;
    const int64_t tib = (int64_t)blockIdx.x * RB + trb;
    const bool tlive = trb < RB && tib < x.nI;
    if (tlive) {
        const int64_t t0 = tib * nJ;
        if (tsl <= TI) s_ofs[0][trb][tsl] = x.cofs[t0 * (TI + 1) + tsl];
        if (tsl < TI) s_perm[0][trb][tsl] = x.cperm[t0 * TI + tsl];
        if (tsl == 0) s_tb[0][trb] = x.tptr[t0];
    }
    __syncthreads();
    for (int Jt = 0; Jt < nJ; ++Jt) {
        const int s = Jt % NS;
        if (tid == 0 && Jt + 2 < nJ) issue(Jt + 2, (Jt + 2) % NS);
        uint16_t nofs = 0; uint8_t nperm = 0; int64_t ntb = 0;
        if (tlive && Jt + 1 < nJ) {
            const int64_t tn = tib * nJ + Jt + 1;
            if (tsl <= TI) nofs = x.cofs[tn * (TI + 1) + tsl];
            if (tsl < TI) nperm = x.cperm[tn * TI + tsl];
            if (tsl == 160) {
                ntb = x.tptr[tn];
                const int64_t nte = x.tptr[tn + 1];
                if (nte > ntb) {       // the next tile's q and local genes into L2
                    const uintptr_t p0 = reinterpret_cast<uintptr_t>(x.q + ntb) & ~(uintptr_t)15;
                    const uintptr_t p1 = (reinterpret_cast<uintptr_t>(x.q + nte) + 15) & ~(uintptr_t)15;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"((unsigned)(p1 - p0)) : "memory");
                    const uintptr_t c0 = reinterpret_cast<uintptr_t>(x.ccol + ntb) & ~(uintptr_t)15;
                    const uintptr_t c1 = (reinterpret_cast<uintptr_t>(x.ccol + nte) + 15) & ~(uintptr_t)15;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c0), "r"((unsigned)(c1 - c0)) : "memory");
                }
            }
        }
        mbar_wait(&full[s], (Jt / NS) & 1);
        const double *st = stage + s * STAGE;
#pragma unroll 1
        for (int rb = 0; rb < RB; ++rb) {
            if ((int64_t)blockIdx.x * RB + rb >= x.nI) break;
            const int rr = rb ? (RPT - 1 - warp) : warp;
            const int64_t tbase = s_tb[Jt & 1][rb];
            const int lr = s_perm[Jt & 1][rb][rr * 8 + grp];
            const int beg = s_ofs[Jt & 1][rb][lr];
            const int len = (int)s_ofs[Jt & 1][rb][lr + 1] - beg;
            const int maxlen = warp_max_int_groups(len);
            if (maxlen == 0) continue;
            const uint8_t *cc = x.ccol + tbase + beg;
            const double *qq = x.q + tbase + beg;
            const int row = rb * TI + lr;
            double racc[F];
#pragma unroll
            for (int i = 0; i < F; ++i) racc[i] = 0.0;
            // a group reads its segment 4 entries at a time (one per lane), two such loads ahead of the one being consumed
            double qn = (l < len) ? __ldg(qq + l) : 0.0, qn2 = (4 + l < len) ? __ldg(qq + 4 + l) : 0.0;
            int cn = (l < len) ? (int)__ldg(cc + l) : 0, cn2 = (4 + l < len) ? (int)__ldg(cc + 4 + l) : 0;
            for (int base = 0; base < maxlen; base += 4) {
                const double qc = qn;
                const int ccur = cn;
                qn = qn2; cn = cn2;
                if (base + 8 < maxlen) {
                    const int idx = base + 8 + l;
                    qn2 = (idx < len) ? __ldg(qq + idx) : 0.0;
                    cn2 = (idx < len) ? (int)__ldg(cc + idx) : 0;
                }
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    if (base + u4 >= maxlen) break;                          // warp-uniform
                    const double q = __shfl_sync(0xffffffffu, qc, u4, 4);
                    const int col = __shfl_sync(0xffffffffu, ccur, u4, 4);
                    double w[F];
                    load_row_vec<KP>(st + col * KP, l, w);
                    const double qf = q > 0.0 ? q : 0.0;
#pragma unroll
                    for (int i = 0; i < F; ++i) racc[i] = fma(qf, w[i], racc[i]);
                    if (q < 0.0) {                                            // -x: the guarded formula (gap...cpp:481-487)
                        double e[F], elu[F];
                        const int64_t ig = i0 + row, jg = (int64_t)Jt * TJ + col;
                        const double Tx = exact_terms_group<KP, false>(a.ElogU + ig * KP, a.ElogV + jg * KP, a.Sd + jg * KP, a.K, l, gmask, e, elu);
                        const double qx = -q / (Tx > 0.0 ? Tx : 1.0);
#pragma unroll
                        for (int i = 0; i < F; ++i) acc[row * KP + M::k(i, l)] += qx * e[i] * a.wS[jg * KP + M::k(i, l)];
                        if (l == 0) atomicAdd(a.dbg, 1ull);
                    }
                }
            }
            if (len > 0) {
                double eur[F], cur[F];
                load_row_vec<KP>(eu_s + row * KP, l, eur);
                load_row_vec<KP>(acc + row * KP, l, cur);
#pragma unroll
                for (int i = 0; i < F; ++i) acc[row * KP + M::k(i, l)] = fma(eur[i], racc[i], cur[i]);
            }
        }
        if (tlive && Jt + 1 < nJ) {
            if (tsl <= TI) s_ofs[(Jt + 1) & 1][trb][tsl] = nofs;
            if (tsl < TI) s_perm[(Jt + 1) & 1][trb][tsl] = nperm;
            if (tsl == 160) s_tb[(Jt + 1) & 1][trb] = ntb;
        }
        __syncthreads();      // the gene block's buffer may be refilled, the cells taken by other groups
    }
    // U-side M-step + U_stats (gap...cpp:530-532, :408-411), as k_estep_rows_q
    constexpr int TPB = (512 / KP) * KP;
    UAcc ua;
    if (tid < TPB) {
        const int k = tid % KP;
        for (int e = tid; e < rows_here * KP; e += TPB) {
            if (k < a.K) {
                const int64_t o = i0 * KP + e;
                const double ez = acc[e];
                const double a1o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.a1_old0[o] : a.a1[o]);
                const double a2o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.a2_old0[o] : a.a2[o]);
                const double a1n = (a.alpha1r ? a.alpha1r[o] : a.alpha1[k]) + ez;
                const double a2n = (a.alpha2r ? a.alpha2r[o] : a.alpha2[k]) + (a.PV ? a.PV[o] : a.rate_vec[k]);
                a.a1[o] = a1n;
                a.a2[o] = a2n;
                double el;
                u_element(a1n, a2n, a1o, a2o, a.EU_new + o, &el, a.eu + o, ua);
                a.ElogU[o] = el;
                acc[e] = el;
            }
        }
        if (k < a.K) { atomicAdd(&scs[0][k], ua.cs_eu); atomicAdd(&scs[1][k], ua.cs_elog); }
    }
    __syncthreads();
    for (int r = tid; r < rows_here; r += blockDim.x) {
        double mn = 1e300, mx = -1e300;
        for (int k = 0; k < a.K; ++k) { const double el = acc[r * KP + k]; mn = fmin(mn, el); mx = fmax(mx, el); }
        a.rowext[i0 + r] = make_double2(mn, mx);
    }
    if (tid < a.K) { atomicAdd(a.cs_eu + tid, scs[0][tid]); atomicAdd(a.cs_elog + tid, scs[1][tid]); }
    const double gd = block_sum(ua.gap_d, sh), go = block_sum(ua.gap_o, sh), en = block_sum(ua.ent, sh);
    const double mn = warp_min(ua.mn), mx = warp_max(ua.mx);
    if (tid == 0) { atomicAdd(a.usc + 0, gd); atomicAdd(a.usc + 1, go); atomicAdd(a.usc + 2, en); }
    if (lane == 0 && mn <= mx) {
        atomicMin(a.guard_next + G_MINU, dbl_ordered(mn));
        atomicMax(a.guard_next + G_MAXU, dbl_ordered(mx));
    }
}

}  // namespace pcmf
