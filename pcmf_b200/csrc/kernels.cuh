// sm_100a kernels of the pCMF variational-EM core.  FP64 throughout (the reference is MatrixXd).
//
// Data layout in HBM:
//   * factor-side state is ROW-major with a padded leading dimension KP (K rounded up to a supported width,
//     multiple of 4 -> every K-vector is 32-byte aligned and read with 16-byte vector loads);
//     padded entries are kept at values that make them inert (eu = ev = 0, ...).
//   * X lives twice: CSR (cells ordered) for the U-side pass and CSC (genes ordered) for the V-side passes,
//     int32 indices + FP32 (or FP64) values, plus two 1-bit/entry occupancy bitmaps for the dense ZI passes.
//   * hyper-parameters alpha*/beta* have identical rows in every reachable state and are
//     stored as K-vectors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "specfun.cuh"

namespace pcmf {

// indices into the per-iteration scalar block (device doubles)
enum Scal : int {
    S_GAPU_D = 0, S_GAPU_O, S_ENTU, S_GAPV_D, S_GAPV_O, S_ENTV,
    S_BERNS_PRIOR, S_BERNS_SELF, S_BERND_PRIOR, S_BERND_SELF, S_SUM_PUVT, S_DATA, S_LGX,
    S_GAMMAU, S_GAMMAV, S_SUMUVT_CLOSED, S_ABS_GAP, S_NORM_GAP, S_ELBO_NODATA, S_LL_ZERO,
    S_COUNT = 32
};
// guard block: min/max of ElogU / ElogV stored as ordered long long
enum Guard : int { G_MINU = 0, G_MAXU, G_MINV, G_MAXV, G_COUNT };

// When may the multinomial weights be computed from the hoisted exponentials eu = exp(ElogU), ev = S o exp(ElogV)
// (T_ij = eu_i . ev_j) instead of the reference's per-entry formula (gap_factor_model.cpp:481-487:
// s_k = ElogU_ik + ElogV_jk, shift by max_k s_k only if it is >= 100, cexp floor 3e-44 below -100)?
//   (1) no shift and no overflow: rowmax_i + max ElogV < 99 (resp. colmax_j + max ElogU), both below 700;
//   (2) the floor only touches terms that are negligible: T >= 1e-26 (a floored term is <= 3.7e-44, so the two forms
//       differ by < K * 3.7e-18 relative, below FP64 rounding), or every factor of the gene is masked (T == 0, r = 0);
// anything else is recomputed with the exact formula for that entry only.
struct PathFlags { bool force_exact; };
__device__ __forceinline__ PathFlags path_flags(double own_min, double own_max, double other_min, double other_max) {
    PathFlags f;
    f.force_exact = !(own_max + other_max < 99.0) || !(own_max < 700.0) || !(other_max < 700.0);
    (void)own_min; (void)other_min;
    return f;
}
// exact multinomial terms of one entry: e_k = S_jk cexp(ElogU_ik + ElogV_jk - m), returns sum_k e_k (the shifted sum)
template <int KP>
__device__ __noinline__ double exact_terms(const double *__restrict__ elu, const double *__restrict__ elv,
                                           const double *__restrict__ sd, int K, double *e) {
    double m = -1e300;
    for (int k = 0; k < K; ++k) { e[k] = elu[k] + elv[k]; m = fmax(m, e[k]); }
    if (m < 100.0) m = 0.0;
    double T = 0.0;
    for (int k = 0; k < K; ++k) { e[k] = custom_exp(e[k] - m) * sd[k]; T += e[k]; }
    for (int k = K; k < KP; ++k) e[k] = 0.0;
    return T;
}

// The same guarded terms of ONE entry evaluated by a whole warp, lane k holding factor k (and k + 32 when KP = 64):
// e_c is the lane's term, the return value the (shifted) sum on every lane.  The non-zero passes hand the entries that
// cannot take the product formula to this routine one at a time (ballot + shuffle): right after a random initialisation
// ~1 % of the entries need it, and one lane evaluating K guarded exponentials alone stalled the other 31 for thousands
// of cycles (C4: first iterations 80 ms instead of 27 ms for the cell pass).
template <int C>
__device__ __forceinline__ double exact_terms_warp(const double (&elu)[C], const double (&elv)[C], const double (&sd)[C], int K, int lane,
                                                   double (&e)[C]) {
    double m = -1e300;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = elu[c] + elv[c]; if (c * 32 + lane < K) m = fmax(m, e[c]); }
    m = warp_max(m);
    if (m < 100.0) m = 0.0;
    double T = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = (c * 32 + lane < K) ? custom_exp(e[c] - m) * sd[c] : 0.0; T += e[c]; }
    return warp_sum(T);
}

// One padded K-vector (32-byte aligned, KP a multiple of 4) with 256-bit loads -- LDG.E.256, new on sm_100.  The gather of
// these rows is what bounds the non-zero passes; measured (scripts/ubench_gather256.cu) the same random 160-byte rows come
// in at 9.1 TB/s with 256-bit loads against 6.4 TB/s with 128-bit ones from an L2-resident table (6.8-7.6 vs 6.2 TB/s
// from a table larger than L2): the limit was the number of requests, not the bytes.
template <int KP>
__device__ __forceinline__ void load_vec(const double *__restrict__ p, double (&r)[KP]) {
    static_assert(KP % 4 == 0, "padded factor widths are multiples of 4");
#pragma unroll
    for (int i = 0; i < KP / 4; ++i)
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(r[4 * i]), "=d"(r[4 * i + 1]), "=d"(r[4 * i + 2]), "=d"(r[4 * i + 3])
                     : "l"(p + 4 * i));
}

__device__ __forceinline__ double block_sum(double v, double *sh /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;   // valid in warp 0
}

// ------------------------------------------------------------------------------------------------------------
// U-side epilogue shared by the E-step row kernel and the init kernel: given the new shape/rate of element (i,k)
// produce the sufficient statistics (gap_factor_model.cpp:408-411, probability.h:76-108) and the per-lane partial
// sums of everything that is reduced over cells.
// ------------------------------------------------------------------------------------------------------------
struct UAcc {
    double cs_eu = 0, cs_elog = 0, gap_d = 0, gap_o = 0, ent = 0;
    double mn = 1e300, mx = -1e300;
};
__device__ __forceinline__ void u_element(double a1n, double a2n, double a1o, double a2o, double *EU, double *ElogU,
                                          double *eu, UAcc &acc) {
    const double d1 = a1n - a1o, d2 = a2n - a2o;
    acc.gap_d += d1 * d1 + d2 * d2;            // gap_factor_model.cpp:258-259
    acc.gap_o += a1o * a1o + a2o * a2o;        // gap_factor_model.cpp:256-257
    const double psi = digamma_pos(a1n);
    const double el = psi - log(a2n);
    const double e = a1n / a2n;
    *EU = e;
    *ElogU = el;
    *eu = exp(el);
    acc.cs_eu += e;
    acc.cs_elog += el;
    acc.ent += gamma_entropy1(a1n, a2n, psi);  // probability.h:157-160
    acc.mn = fmin(acc.mn, el);
    acc.mx = fmax(acc.mx, el);
}

struct RowArgs {
    int64_t n;
    int K;
    const int64_t *rowptr;
    const int32_t *colidx;
    const void *val;
    // V side (old stats)
    const double *ev;      // p x KP: S o exp(ElogV)
    const double *evw;     // p x KP: ev o wS
    const double *ElogV;   // p x KP                                       (exact fallback)
    const double *Sd;      // p x KP mask as double                        (exact fallback)
    const double *wS;      // p x KP: prob_S o colw (or colw, or 1)        (exact fallback)
    const uint8_t *allmasked;   // p: 1 when S(j,.) is all zero (sparse models)
    unsigned long long *dbg;    // [0] entries recomputed with the exact formula
    // U side, updated in place
    double *a1, *a2, *EU_new, *ElogU, *eu;
    double2 *rowext;                 // n: (min_k, max_k) of ElogU per cell, read then refreshed
    const double *alpha1, *alpha2;   // K-vectors
    const double *alpha1r, *alpha2r; // or n x KP (hyper-parameter rows that differ, gap_factor_model.cpp:126-132): override the vectors
    const double *rate_vec;          // KP: colsum(EV o prob_S) (non-ZI)   gap...cpp:531 / sparse...cpp:400
    const double *PV;                // n x KP: prob_D * (EV o prob_S) (ZI) zi...cpp:336 / zi_sparse...cpp:363
    int first_iter;                  // 1: old iterate is Ones (gap_factor_model.cpp:72-75; never copied before iter 0)
                                     // 2: old iterate is a1_old0 / a2_old0 (first iteration of a later multi-init run:
                                     //    the old copy still holds the previous run's last iterate, wrapper_gap_factor.h:296-350)
    const double *a1_old0, *a2_old0;
    const long long *guard_cur;      // global extremes of the stats being read
    long long *guard_next;           // min/max of the new ElogU
    double *cs_eu, *cs_elog;         // KP each (atomic)
    double *usc;                     // 3 doubles (atomic): gap_d, gap_o, entropy of the local cells (all-reduced later)
    // stored multinomial weights (k_estep_rows_q): q_ij = x_ij / T_ij written by the gene passes in CSC order
    const uint32_t *csr2csc;         // nnz: CSC position of every CSR entry
    const double *qcsc;              // nnz: q in CSC order; NaN marks an entry that needs the exact formula
};

// E-step, cell side + fused U-side M-step.  One warp per cell; lanes stride over the cell's non-zeros.
// Reference: update_variational_multinomial_param (gap...cpp:458-499, zi...:289-330, sparse...:327-368,
// zi_sparse...:316-357) restricted to the entries with X != 0 (the others contribute 0), then
// update_variational_gamma_param, factor U part (gap...cpp:530-532 and variants) + U_stats.
template <int KP, typename VT>
__global__ void __launch_bounds__(256) k_estep_rows(RowArgs a) {
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * warps_per_block + w;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    const VT *__restrict__ val = static_cast<const VT *>(a.val);
    const double gminV = dbl_from_ordered(a.guard_cur[G_MINV]), gmaxV = dbl_from_ordered(a.guard_cur[G_MAXV]);
    const bool same_w = (a.evw == a.ev);
    UAcc acc[(KP + 31) / 32];

    constexpr int C = (KP + 31) / 32;
    for (int64_t i = warp0; i < a.n; i += nwarps) {
        double u[KP], r[KP];
        load_vec<KP>(a.eu + i * KP, u);
#pragma unroll
        for (int k = 0; k < KP; ++k) r[k] = 0.0;
        double sx[C], elu[C];                      // exact-path contributions / ElogU_i of the factor(s) this lane owns
#pragma unroll
        for (int c = 0; c < C; ++c) { sx[c] = 0.0; elu[c] = (c * 32 + lane < KP) ? a.ElogU[i * KP + c * 32 + lane] : 0.0; }
        const double2 ext = a.rowext[i];
        const PathFlags pf = path_flags(ext.x, ext.y, gminV, gmaxV);
        const int64_t e0 = a.rowptr[i], e1 = a.rowptr[i + 1];
        for (int64_t eb = e0; eb < e1; eb += 32) {              // warp-uniform trip count: the exact path is cooperative
            const int64_t e = eb + lane;
            const bool active = e < e1;
            int j = 0; double x = 0.0; bool slow = false;
            if (active) {
                j = __ldg(a.colidx + e);
                x = (double)__ldg(val + e);
                double v[KP];
                load_vec<KP>(a.ev + (int64_t)j * KP, v);
                double T0 = 0.0, T1 = 0.0;
#pragma unroll
                for (int k = 0; k < KP; k += 2) { T0 = fma(u[k], v[k], T0); T1 = fma(u[k + 1], v[k + 1], T1); }
                const double T = T0 + T1;
                if (!pf.force_exact && T >= 1e-26) {
                    const double q = x / T;
                    if (!same_w) load_vec<KP>(a.evw + (int64_t)j * KP, v);
#pragma unroll
                    for (int k = 0; k < KP; ++k) r[k] = fma(q, v[k], r[k]);
                } else slow = !a.allmasked[j];   // all-masked gene: T = 0, r_ijk = 0 (gap...cpp:487), nothing to add
            }
            unsigned need = __ballot_sync(0xffffffffu, slow);
            if (need && lane == 0) atomicAdd(a.dbg, (unsigned long long)__popc(need));
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const int js = __shfl_sync(0xffffffffu, j, src);
                const double xs = __shfl_sync(0xffffffffu, x, src);
                double elv[C], sd[C], ws[C], ex[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int k = c * 32 + lane;
                    const bool in = k < KP;
                    elv[c] = in ? a.ElogV[(int64_t)js * KP + k] : 0.0; sd[c] = in ? a.Sd[(int64_t)js * KP + k] : 0.0;
                    ws[c] = in ? a.wS[(int64_t)js * KP + k] : 0.0;
                }
                const double Tx = exact_terms_warp<C>(elu, elv, sd, a.K, lane, ex);
                const double q = xs / (Tx > 0.0 ? Tx : 1.0);
#pragma unroll
                for (int c = 0; c < C; ++c) sx[c] = fma(q * ex[c], ws[c], sx[c]);
            }
        }
        // reduce over lanes; lane (k & 31) keeps component k
        double mine[(KP + 31) / 32], um[(KP + 31) / 32];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const double t = warp_sum(r[k]);
            if ((k & 31) == lane) { mine[k >> 5] = t; um[k >> 5] = u[k]; }
        }
        double emn = 1e300, emx = -1e300;
#pragma unroll
        for (int c = 0; c < (KP + 31) / 32; ++c) {
            const int k = c * 32 + lane;
            if (k < a.K) {
                const int64_t o = i * KP + k;
                const double ez = um[c] * mine[c] + sx[c];
                const double a1o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.a1_old0[o] : a.a1[o]);
                const double a2o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.a2_old0[o] : a.a2[o]);
                const double a1n = (a.alpha1r ? a.alpha1r[o] : a.alpha1[k]) + ez;
                const double a2n = (a.alpha2r ? a.alpha2r[o] : a.alpha2[k]) + (a.PV ? a.PV[o] : a.rate_vec[k]);
                a.a1[o] = a1n;
                a.a2[o] = a2n;
                double el;
                u_element(a1n, a2n, a1o, a2o, a.EU_new + o, &el, a.eu + o, acc[c]);
                a.ElogU[o] = el;
                emn = fmin(emn, el); emx = fmax(emx, el);
            }
        }
        emn = warp_min(emn); emx = warp_max(emx);
        if (lane == 0) a.rowext[i] = make_double2(emn, emx);
        __syncwarp();
    }
    // flush per-lane partials
    double gd = 0, go = 0, en = 0, mn = 1e300, mx = -1e300;
#pragma unroll
    for (int c = 0; c < (KP + 31) / 32; ++c) {
        const int k = c * 32 + lane;
        if (k < a.K) {
            atomicAdd(a.cs_eu + k, acc[c].cs_eu);
            atomicAdd(a.cs_elog + k, acc[c].cs_elog);
        }
        gd += acc[c].gap_d; go += acc[c].gap_o; en += acc[c].ent;
        mn = fmin(mn, acc[c].mn); mx = fmax(mx, acc[c].mx);
    }
    gd = block_sum(gd, sh); go = block_sum(go, sh); en = block_sum(en, sh);
    mn = warp_min(mn); mx = warp_max(mx);
    if (threadIdx.x == 0) {
        atomicAdd(a.usc + 0, gd);
        atomicAdd(a.usc + 1, go);
        atomicAdd(a.usc + 2, en);
    }
    if (lane == 0 && mn <= mx) {
        atomicMin(a.guard_next + G_MINU, dbl_ordered(mn));
        atomicMax(a.guard_next + G_MAXU, dbl_ordered(mx));
    }
}

// Same pass when the gene passes have left q_ij = x_ij / T_ij for every non-zero (CSC order): the cell side is then a
// pure SpMM, EZ_j(i,.) = eu_i o sum_j q_ij evw_j -- one gathered K-vector per non-zero instead of two (ev for T, evw for
// the sum), no division, half the registers.  q is read through the CSR -> CSC position map; a NaN marks an entry whose
// weights need the reference's guarded formula (exact_terms), recomputed here exactly as k_estep_rows does.
// Used for the ZI / sparse models (evw != ev); the plain model gathers one vector either way.
template <int KP, typename VT>
__global__ void __launch_bounds__(256, 2) k_estep_rows_q(RowArgs a) {
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * warps_per_block + w;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    const VT *__restrict__ val = static_cast<const VT *>(a.val);
    UAcc acc[(KP + 31) / 32];

    constexpr int C = (KP + 31) / 32;
    for (int64_t i = warp0; i < a.n; i += nwarps) {
        double r[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) r[k] = 0.0;
        double sx[C];
#pragma unroll
        for (int c = 0; c < C; ++c) sx[c] = 0.0;
        const int64_t e0 = a.rowptr[i], e1 = a.rowptr[i + 1];
        for (int64_t eb = e0; eb < e1; eb += 32) {              // warp-uniform trip count: the exact path is cooperative
            const int64_t e = eb + lane;
            int j = 0; bool slow = false;
            if (e < e1) {
                j = __ldg(a.colidx + e);
                const uint32_t t = __ldg(a.csr2csc + e);
                double v[KP];
                load_vec<KP>(a.evw + (int64_t)j * KP, v);      // requested together with q, not after it (latency bound)
                const double q = __ldg(a.qcsc + t);
                if (q == q) {
#pragma unroll
                    for (int k = 0; k < KP; ++k) r[k] = fma(q, v[k], r[k]);
                } else slow = !a.allmasked[j];
            }
            unsigned need = __ballot_sync(0xffffffffu, slow);
            if (need && lane == 0) atomicAdd(a.dbg, (unsigned long long)__popc(need));
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const int js = __shfl_sync(0xffffffffu, j, src);
                const double xs = (double)__ldg(val + eb + src);
                double elu[C], elv[C], sd[C], ws[C], ex[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int k = c * 32 + lane;
                    const bool in = k < KP;
                    elu[c] = in ? a.ElogU[i * KP + k] : 0.0;
                    elv[c] = in ? a.ElogV[(int64_t)js * KP + k] : 0.0; sd[c] = in ? a.Sd[(int64_t)js * KP + k] : 0.0;
                    ws[c] = in ? a.wS[(int64_t)js * KP + k] : 0.0;
                }
                const double Tx = exact_terms_warp<C>(elu, elv, sd, a.K, lane, ex);
                const double qx = xs / (Tx > 0.0 ? Tx : 1.0);
#pragma unroll
                for (int c = 0; c < C; ++c) sx[c] = fma(qx * ex[c], ws[c], sx[c]);
            }
        }
        double mine[(KP + 31) / 32];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const double t = warp_sum(r[k]);
            if ((k & 31) == lane) mine[k >> 5] = t;
        }
        double emn = 1e300, emx = -1e300;
#pragma unroll
        for (int c = 0; c < (KP + 31) / 32; ++c) {
            const int k = c * 32 + lane;
            if (k < a.K) {
                const int64_t o = i * KP + k;
                const double ez = a.eu[o] * mine[c] + sx[c];
                const double a1o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.a1_old0[o] : a.a1[o]);
                const double a2o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.a2_old0[o] : a.a2[o]);
                const double a1n = (a.alpha1r ? a.alpha1r[o] : a.alpha1[k]) + ez;
                const double a2n = (a.alpha2r ? a.alpha2r[o] : a.alpha2[k]) + (a.PV ? a.PV[o] : a.rate_vec[k]);
                a.a1[o] = a1n;
                a.a2[o] = a2n;
                double el;
                u_element(a1n, a2n, a1o, a2o, a.EU_new + o, &el, a.eu + o, acc[c]);
                a.ElogU[o] = el;
                emn = fmin(emn, el); emx = fmax(emx, el);
            }
        }
        emn = warp_min(emn); emx = warp_max(emx);
        if (lane == 0) a.rowext[i] = make_double2(emn, emx);
        __syncwarp();
    }
    double gd = 0, go = 0, en = 0, mn = 1e300, mx = -1e300;
#pragma unroll
    for (int c = 0; c < (KP + 31) / 32; ++c) {
        const int k = c * 32 + lane;
        if (k < a.K) {
            atomicAdd(a.cs_eu + k, acc[c].cs_eu);
            atomicAdd(a.cs_elog + k, acc[c].cs_elog);
        }
        gd += acc[c].gap_d; go += acc[c].gap_o; en += acc[c].ent;
        mn = fmin(mn, acc[c].mn); mx = fmax(mx, acc[c].mx);
    }
    gd = block_sum(gd, sh); go = block_sum(go, sh); en = block_sum(en, sh);
    mn = warp_min(mn); mx = warp_max(mx);
    if (threadIdx.x == 0) {
        atomicAdd(a.usc + 0, gd);
        atomicAdd(a.usc + 1, go);
        atomicAdd(a.usc + 2, en);
    }
    if (lane == 0 && mn <= mx) {
        atomicMin(a.guard_next + G_MINU, dbl_ordered(mn));
        atomicMax(a.guard_next + G_MAXU, dbl_ordered(mx));
    }
}

// init: statistics of given (a1, a2) without an E-step (gap_factor_model.cpp:119, 140, 170: U_stats()).
struct UInitArgs {
    int64_t n; int K, KP;
    const double *a1, *a2;
    double *EU, *ElogU, *eu;
    long long *guard_next;
    double *cs_eu, *cs_elog, *usc;
    int ones;      // 1: the statistics keep the constructor's Ones (model.cpp:220-223) whatever a1, a2 say -- the state
                   // init_hyper_param alone leaves behind (wrapper_gap_factor.h:366-367); the entropy still comes from (a1, a2)
};
#ifdef PCMF_COMMON_KERNELS
__global__ void k_stats_u(UInitArgs a) {
    __shared__ double sh[32];
    __shared__ double scs[2][64];
    if (threadIdx.x < 64) { scs[0][threadIdx.x] = 0; scs[1][threadIdx.x] = 0; }
    __syncthreads();
    UAcc acc;
    const int64_t total = a.n * a.KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % a.KP);
        if (k < a.K) {
            UAcc one;
            u_element(a.a1[o], a.a2[o], 1.0, 1.0, a.EU + o, a.ElogU + o, a.eu + o, one);
            if (a.ones) { a.EU[o] = 1.0; a.ElogU[o] = 1.0; a.eu[o] = exp(1.0); one.cs_eu = 1.0; one.cs_elog = 1.0; one.mn = 1.0; one.mx = 1.0; }
            atomicAdd(&scs[0][k], one.cs_eu);
            atomicAdd(&scs[1][k], one.cs_elog);
            acc.ent += one.ent;
            acc.mn = fmin(acc.mn, one.mn); acc.mx = fmax(acc.mx, one.mx);
        } else {
            a.EU[o] = 0; a.ElogU[o] = 0; a.eu[o] = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x < a.K) {
        atomicAdd(a.cs_eu + threadIdx.x, scs[0][threadIdx.x]);
        atomicAdd(a.cs_elog + threadIdx.x, scs[1][threadIdx.x]);
    }
    const double en = block_sum(acc.ent, sh);
    const double mn = warp_min(acc.mn), mx = warp_max(acc.mx);
    if (threadIdx.x == 0) atomicAdd(a.usc + 2, en);
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        atomicMin(a.guard_next + G_MINU, dbl_ordered(mn));
        atomicMax(a.guard_next + G_MAXU, dbl_ordered(mx));
    }
}
#endif  // PCMF_COMMON_KERNELS

// ------------------------------------------------------------------------------------------------------------
// Gene-side non-zero pass over CSC: one block per gene.  B1(j,k) = sum_i q_ij e_ijk, B2(j,k) = sum_i q_ij e_ijk
// ElogU(i,k), xlogT(j) = sum_i x_ij log T_ij, with e_ijk = S_jk cexp(ElogU_ik + ElogV_jk - m), T = sum_k e, q = x / T'.
// Used three ways:
//   E-step gene side (old stats): EZ_i = wS o B1 (gap...cpp:491 and variants); B2 / xlogT feed the ELBO data term of
//     the previous iterate (gap...cpp:308-312, sparse...cpp:194-212);
//   sparse-probability pass (new stats, old mask): sparse...cpp:444-460, zi_sparse...cpp:433-449;
//   final ELBO data term.
// ------------------------------------------------------------------------------------------------------------
struct ColArgs {
    int32_t p; int K;
    const int64_t *colptr;
    const int32_t *rowidx;
    const void *val;
    const double *eu, *ElogU;         // n x KP
    const double *ev, *ElogV, *Sd;    // p x KP
    const double2 *colext;            // p: (min_k, max_k) of ElogV per gene
    const uint8_t *allmasked;         // p
    const uint8_t *reuse;             // p or null: 1 -> copy the gene's sums from B1_src/B2_src instead of recomputing
    const double *B1_src, *B2_src;    // p x KP: this rank's sums from the previous sparse-probability pass
    unsigned long long *dbg;          // [1] entries recomputed with the exact formula
    const long long *guard;           // global extremes (U part used)
    double *B1, *B2;                  // p x KP (plain stores: the block owns the gene)
    double *xlogT;                    // p
    double *qout;                     // nnz or null: q_ij = x_ij / T_ij of every entry visited, CSC order (NaN: exact-formula entry)
    // Cell blocks (null / 1: the whole gene in one go).  blkptr[b * p + j] = first CSC position of gene j with a cell >= b * rows
    // per block; blockIdx.y = b walks the blocks one after the other, so that the slice of the cell table a block gathers from
    // (eu [+ ElogU] of ~1e5 cells) stays in L2 instead of the whole n x K table streaming through it; sums are then added atomically.
    const int64_t *blkptr; int nblk;
    // tiled copy of X (xtile.cuh): q goes to qout[qmap[e]] (the cell-order position of CSC entry e) and an entry that needs
    // the guarded formula is marked -x instead of NaN
    const uint32_t *qmap;
};
#ifndef PCMF_GENE_MINB
#define PCMF_GENE_MINB 2
#endif
template <int KP, typename VT, bool WITH_B2, bool WITH_LOGT>
__global__ void __launch_bounds__(128, PCMF_GENE_MINB) k_gene_pass(ColArgs a) {
    __shared__ double sh[4][KP];
    __shared__ double sh2[4][KP];
    __shared__ double shl[4];
    __shared__ double sx1[4][KP], sx2[4][KP];   // exact-path contributions of the four warps
    constexpr int C = (KP + 31) / 32;
    const VT *__restrict__ val = static_cast<const VT *>(a.val);
    const double gminU = dbl_from_ordered(a.guard[G_MINU]), gmaxU = dbl_from_ordered(a.guard[G_MAXU]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool blocked = a.nblk > 1;
    const int blk = blocked ? (int)blockIdx.y : 0;
    for (int j = blockIdx.x; j < a.p; j += gridDim.x) {
        if (a.reuse && a.reuse[j]) {
            // same statistics, same mask row as the sparse-probability pass of the previous iteration
            // (sparse...cpp:448-451 vs :350-356): the sums are identical, copy them
            if (threadIdx.x < KP && blk == 0) {
                a.B1[(int64_t)j * KP + threadIdx.x] = a.B1_src[(int64_t)j * KP + threadIdx.x];
                if (WITH_B2) a.B2[(int64_t)j * KP + threadIdx.x] = a.B2_src[(int64_t)j * KP + threadIdx.x];
            }
            continue;
        }
        double v[KP], b1[KP], b2[WITH_B2 ? KP : 1];
        load_vec<KP>(a.ev + (int64_t)j * KP, v);
        // exact-path contributions: lane k of every warp owns factor k (and k + 32 for KP = 64)
        double x1[C], x2[C], elv[C], sdv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int k = c * 32 + lane;
            x1[c] = 0.0; x2[c] = 0.0;
            elv[c] = k < KP ? a.ElogV[(int64_t)j * KP + k] : 0.0; sdv[c] = k < KP ? a.Sd[(int64_t)j * KP + k] : 0.0;
        }
        const double2 ext = a.colext[j];
        const PathFlags pf = path_flags(ext.x, ext.y, gminU, gmaxU);
#pragma unroll
        for (int k = 0; k < KP; ++k) { b1[k] = 0.0; if (WITH_B2) b2[k] = 0.0; }
        double lt = 0.0;
        // all-masked gene: every r_ijk is 0 (sparse...cpp:350-356), nothing to accumulate
        const int64_t g0 = blocked ? a.blkptr[(int64_t)blk * a.p + j] : a.colptr[j];
        const int64_t g1 = blocked ? a.blkptr[(int64_t)(blk + 1) * a.p + j] : a.colptr[j + 1];
        const int64_t e0 = g0, e1 = a.allmasked[j] ? g0 : g1;
        if (a.qout && a.allmasked[j])
            for (int64_t e = e0 + threadIdx.x; e < g1; e += blockDim.x) a.qout[a.qmap ? (int64_t)a.qmap[e] : e] = 0.0;
        for (int64_t eb = e0 + w * 32; eb < e1; eb += blockDim.x) {    // warp-uniform trip count: the exact path is cooperative
            const int64_t e = eb + lane;
            int64_t i = 0; double x = 0.0; bool slow = false;
            if (e < e1) {
                i = __ldg(a.rowidx + e);
                x = (double)__ldg(val + e);
                // both rows of the cell are requested before anything is consumed: the pass is bound by the latency of these
                // gathers, and a second request issued only once T is known would double it
                double u[KP], el[WITH_B2 ? KP : 1];
                load_vec<KP>(a.eu + i * KP, u);
                if (WITH_B2) load_vec<KP>(a.ElogU + i * KP, reinterpret_cast<double (&)[KP]>(el));
                double T0 = 0.0, T1 = 0.0;
#pragma unroll
                for (int k = 0; k < KP; k += 2) { T0 = fma(u[k], v[k], T0); T1 = fma(u[k + 1], v[k + 1], T1); }
                const double T = T0 + T1;
                if (!pf.force_exact && T >= 1e-26) {
                    const double q = x / T;
                    if (a.qout) a.qout[a.qmap ? (int64_t)a.qmap[e] : e] = q;
                    if (WITH_LOGT) lt += x * log(T);
#pragma unroll
                    for (int k = 0; k < KP; ++k) b1[k] = fma(q, u[k], b1[k]);
                    if (WITH_B2) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) b2[k] = fma(q * u[k], el[k], b2[k]);
                    }
                } else {
                    slow = true;
                    // NaN (or -x for the tiled cell pass): the cell pass takes the exact formula too
                    if (a.qout) { if (a.qmap) a.qout[a.qmap[e]] = -x; else a.qout[e] = __longlong_as_double(0x7ff8000000000000LL); }
                }
            }
            unsigned need = __ballot_sync(0xffffffffu, slow);
            if (need && lane == 0) atomicAdd(a.dbg + 1, (unsigned long long)__popc(need));
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const int64_t is = __shfl_sync(0xffffffffu, i, src);
                const double xs = __shfl_sync(0xffffffffu, x, src);
                double elu[C], ex[C];
#pragma unroll
                for (int c = 0; c < C; ++c) elu[c] = (c * 32 + lane < KP) ? a.ElogU[is * KP + c * 32 + lane] : 0.0;
                const double Tx = exact_terms_warp<C>(elu, elv, sdv, a.K, lane, ex);
                const double q = xs / (Tx > 0.0 ? Tx : 1.0);
                if (WITH_LOGT && lane == 0) lt += xs * log(Tx);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    x1[c] = fma(q, ex[c], x1[c]);
                    if (WITH_B2) x2[c] = fma(q * ex[c], elu[c], x2[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (c * 32 + lane < KP) { sx1[w][c * 32 + lane] = x1[c]; sx2[w][c * 32 + lane] = x2[c]; }
        // block reduction: warp shuffles then 4 partial rows in shared memory
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const double t = warp_sum(b1[k]);
            if (lane == 0) sh[w][k] = t;
            if (WITH_B2) {
                const double t2 = warp_sum(b2[k]);
                if (lane == 0) sh2[w][k] = t2;
            }
        }
        if (WITH_LOGT) { lt = warp_sum(lt); if (lane == 0) shl[w] = lt; }
        __syncthreads();
        if (threadIdx.x < KP) {
            const int k = threadIdx.x;
            // fast path accumulated sum_i q eu_ik; e_ijk = eu_ik ev_jk
            const double evk = __ldg(a.ev + (int64_t)j * KP + k);
            const double s1 = (sh[0][k] + sh[1][k] + sh[2][k] + sh[3][k]) * evk + ((sx1[0][k] + sx1[1][k]) + (sx1[2][k] + sx1[3][k]));
            if (blocked) atomicAdd(a.B1 + (int64_t)j * KP + k, s1); else a.B1[(int64_t)j * KP + k] = s1;
            if (WITH_B2) {
                const double s2 = (sh2[0][k] + sh2[1][k] + sh2[2][k] + sh2[3][k]) * evk + ((sx2[0][k] + sx2[1][k]) + (sx2[2][k] + sx2[3][k]));
                if (blocked) atomicAdd(a.B2 + (int64_t)j * KP + k, s2); else a.B2[(int64_t)j * KP + k] = s2;
            }
        }
        if (WITH_LOGT && threadIdx.x == 0) {
            const double sl = shl[0] + shl[1] + shl[2] + shl[3];
            if (blocked) atomicAdd(a.xlogT + j, sl); else a.xlogT[j] = sl;
        }
        __syncthreads();
    }
}

// (min_k, max_k) of a padded row-major matrix, one thread per row: per-cell / per-gene extremes for path_flags
__global__ void k_row_extremes(int64_t rows, int K, int KP, const double *__restrict__ M, double2 *__restrict__ ext);

// ------------------------------------------------------------------------------------------------------------
// V-side M-step, one thread per (gene, factor): update_variational_gamma_param, factor V part
// (gap...cpp:535-537, zi...:340-342, sparse...:404-420, zi_sparse...:367-385) + V_stats, the V half of
// gap_between_iterates, the V entropy, and -- before the old statistics are overwritten -- the ELBO data term
// of the previous iterate (sparse...cpp:206-210).
// ------------------------------------------------------------------------------------------------------------
struct VArgs {
    int32_t p; int K, KP;
    int sparse, zi, first_iter, want_data, stats_only;
    const double *B1, *B2, *xlogT;        // from the E-step gene pass (all-reduced)
    const double *tmpMat;                 // p x KP: prob_D^T EU_new (ZI, all-reduced)
    const double *cs_eu;                  // KP: colsum(EU_new) (all-reduced)
    const double *beta1, *beta2;          // K-vectors
    const double *beta1r, *beta2r;        // or p x KP rows (see RowArgs::alpha1r)
    int ones;                             // stats_only: statistics are the constructor's Ones (see UInitArgs::ones)
    double *b1, *b2, *EV, *ElogV;
    const double *prob_S, *Sd, *colw;     // old sparse state / nnz weight of the gene
    double *ev, *evw, *EVS, *wS;          // refreshed here for the non-sparse models; ev (old mask) for the sparse pass
    long long *guard_next;
    double *cs_ev, *cs_elogv, *cs_evs;    // KP each
    double *scal;
    const double *b1_old0, *b2_old0;      // first_iter == 2 (see RowArgs)
};
#ifdef PCMF_COMMON_KERNELS
__global__ void __launch_bounds__(256) k_mstep_genes(VArgs a) {
    __shared__ double sh[32];
    __shared__ double scs[3][64];
    if (threadIdx.x < 64) { scs[0][threadIdx.x] = 0; scs[1][threadIdx.x] = 0; scs[2][threadIdx.x] = 0; }
    __syncthreads();
    double gd = 0, go = 0, en = 0, data = 0, mn = 1e300, mx = -1e300;
    const int64_t total = (int64_t)a.p * a.KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % a.KP);
        const int j = (int)(o / a.KP);
        if (k >= a.K) {
            a.EV[o] = 0; a.ElogV[o] = 0; a.ev[o] = 0; a.evw[o] = 0; a.EVS[o] = 0; a.wS[o] = 0;
            continue;
        }
        const double ps = a.sparse ? a.prob_S[o] : 1.0;
        const double cw = a.zi ? a.colw[j] : 1.0;
        double b1n, b2n, b1o, b2o;
        if (!a.stats_only) {
            if (a.want_data) {
                // ELBO data term of the state the E-step just read (old ElogV / prob_S / colw)
                if (a.sparse) data += ps * cw * (a.B2[o] + a.ElogV[o] * a.B1[o]);
                else if (k == 0) data += cw * a.xlogT[j];
            }
            b1o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.b1_old0[o] : a.b1[o]);
            b2o = a.first_iter == 1 ? 1.0 : (a.first_iter == 2 ? a.b2_old0[o] : a.b2[o]);
            b1n = (a.beta1r ? a.beta1r[o] : a.beta1[k]) + ps * cw * a.B1[o];
            b2n = (a.beta2r ? a.beta2r[o] : a.beta2[k]) + ps * (a.zi ? a.tmpMat[o] : a.cs_eu[k]);
            a.b1[o] = b1n;
            a.b2[o] = b2n;
            const double d1 = b1n - b1o, d2 = b2n - b2o;
            gd += d1 * d1 + d2 * d2;
            go += b1o * b1o + b2o * b2o;
        } else {
            b1n = a.b1[o];
            b2n = a.b2[o];
        }
        const double psi = digamma_pos(b1n);
        const double el = (a.stats_only && a.ones) ? 1.0 : psi - log(b2n);
        const double EVn = (a.stats_only && a.ones) ? 1.0 : b1n / b2n;
        a.EV[o] = EVn;
        a.ElogV[o] = el;
        en += gamma_entropy1(b1n, b2n, psi);
        mn = fmin(mn, el); mx = fmax(mx, el);
        atomicAdd(&scs[0][k], EVn);
        atomicAdd(&scs[1][k], el);
        // multinomial weights for the passes that follow.  Sparse models: old mask for the sparse-probability pass
        // (sparse...cpp:448-451); everything is refreshed again by k_sparse_genes once the new mask exists.
        const double sd = a.sparse ? a.Sd[o] : 1.0;
        const double evn = sd * exp(el);
        a.ev[o] = evn;
        a.wS[o] = ps * cw;
        a.evw[o] = evn * ps * cw;
        a.EVS[o] = EVn * ps;
        atomicAdd(&scs[2][k], EVn * ps);
    }
    __syncthreads();
    if (threadIdx.x < a.K) {
        atomicAdd(a.cs_ev + threadIdx.x, scs[0][threadIdx.x]);
        atomicAdd(a.cs_elogv + threadIdx.x, scs[1][threadIdx.x]);
        atomicAdd(a.cs_evs + threadIdx.x, scs[2][threadIdx.x]);
    }
    gd = block_sum(gd, sh); go = block_sum(go, sh); en = block_sum(en, sh); data = block_sum(data, sh);
    mn = warp_min(mn); mx = warp_max(mx);
    if (threadIdx.x == 0) {
        if (!a.stats_only) { atomicAdd(a.scal + S_GAPV_D, gd); atomicAdd(a.scal + S_GAPV_O, go); }
        atomicAdd(a.scal + S_ENTV, en);
        if (a.want_data) atomicAdd(a.scal + S_DATA, data);
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        atomicMin(a.guard_next + G_MINV, dbl_ordered(mn));
        atomicMax(a.guard_next + G_MAXV, dbl_ordered(mx));
    }
}
#endif  // PCMF_COMMON_KERNELS

// ELBO data term only (final loss / on-demand ELBO): same expression as in k_mstep_genes, nothing is updated.
struct DataArgs {
    int32_t p; int K, KP; int sparse, zi;
    const double *B1, *B2, *xlogT, *ElogV, *prob_S, *colw;
    double *scal;
};
#ifdef PCMF_COMMON_KERNELS
__global__ void k_data_term(DataArgs a) {
    __shared__ double sh[32];
    double data = 0;
    const int64_t total = (int64_t)a.p * a.KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % a.KP);
        const int j = (int)(o / a.KP);
        if (k >= a.K) continue;
        const double cw = a.zi ? a.colw[j] : 1.0;
        if (a.sparse) data += a.prob_S[o] * cw * (a.B2[o] + a.ElogV[o] * a.B1[o]);
        else if (k == 0) data += cw * a.xlogT[j];
    }
    data = block_sum(data, sh);
    if (threadIdx.x == 0) atomicAdd(a.scal + S_DATA, data);
}
#endif  // PCMF_COMMON_KERNELS

// ------------------------------------------------------------------------------------------------------------
// Spike-and-slab update, one thread per gene: update_variational_sparse_param (sparse...cpp:424-471,
// zi_sparse...:413-460) from the all-reduced gene-pass sums, S = prob_S >= sel_bound, then
// update_prior_sparse_param (sparse...cpp:474-476), the Bernoulli ELBO terms (sparse...cpp:172-173) and the refreshed
// multinomial weights for the next E-step.
// ------------------------------------------------------------------------------------------------------------
struct SparseArgs {
    int32_t p; int K, KP; int zi, update;
    double sel_bound;
    const double *B1, *B2;        // sparse pass (new stats, old mask), all-reduced
    const double *tmpMat;         // ZI: prob_D^T EU_new
    const double *cs_eu;          // non-ZI: colsum(EU_new)
    const double *EV, *ElogV, *colw;
    double *prob_S, *Sd, *prior_S;
    int32_t *S;
    double *ev, *evw, *EVS, *wS;
    double *cs_evs;
    uint8_t *allmasked;
    uint8_t *unchanged;   // p: 1 when the gene's mask row S(j,.) did not change in this update
    double *scal;
};
#ifdef PCMF_COMMON_KERNELS
__global__ void __launch_bounds__(128) k_sparse_genes(SparseArgs a) {
    __shared__ double sh[32];
    __shared__ double scs[64];
    if (threadIdx.x < 64) scs[threadIdx.x] = 0;
    __syncthreads();
    double bp = 0, bs = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.p; j += gridDim.x * blockDim.x) {
        const double pr = a.prior_S[j];
        const double cw = a.zi ? a.colw[j] : 1.0;
        const double lg = logit_clamped(pr);
        double rs = 0;
        int any = 0, chg = 0;
        for (int k = 0; k < a.K; ++k) {
            const int64_t o = (int64_t)j * a.KP + k;
            double ps;
            if (a.update) {
                if (pr == 1.0) ps = 1.0;
                else if (pr == 0.0) ps = 0.0;
                else {
                    const double t = -a.EV[o] * (a.zi ? a.tmpMat[o] : a.cs_eu[k]) + cw * (a.B2[o] + a.ElogV[o] * a.B1[o]);
                    ps = expit_clamped(lg + t);
                }
                a.prob_S[o] = ps;
            } else {
                ps = a.prob_S[o];
            }
            const int s = ps >= a.sel_bound;
            any |= s;
            chg |= (s != a.S[o]);
            a.S[o] = s;
            a.Sd[o] = (double)s;
            const double evn = (double)s * exp(a.ElogV[o]);
            a.ev[o] = evn;
            a.wS[o] = ps * cw;
            a.evw[o] = evn * ps * cw;
            const double evs = a.EV[o] * ps;
            a.EVS[o] = evs;
            atomicAdd(&scs[k], evs);
            rs += ps;
        }
        a.allmasked[j] = any ? 0 : 1;
        a.unchanged[j] = (a.update && !chg) ? 1 : 0;
        const double prn = rs / a.K;
        a.prior_S[j] = prn;
        // likelihood::bernoulli_loglike_vec(prob_S, prior_S) with the intended gene index + bernoulli_loglike(prob_S, prob_S)
        const bool gate = prn > 0.0 && prn < 1.0;
        const double l1 = gate ? log(prn) : 0.0, l0 = gate ? log(1.0 - prn) : 0.0;
        for (int k = 0; k < a.K; ++k) {
            const double ps = a.prob_S[(int64_t)j * a.KP + k];
            if (gate) bp += ps * l1 + (1.0 - ps) * l0;
            if (ps > 0.0 && ps < 1.0) bs += ps * log(ps) + (1.0 - ps) * log(1.0 - ps);
        }
    }
    __syncthreads();
    if (threadIdx.x < a.K) atomicAdd(a.cs_evs + threadIdx.x, scs[threadIdx.x]);
    bp = block_sum(bp, sh); bs = block_sum(bs, sh);
    if (threadIdx.x == 0) { atomicAdd(a.scal + S_BERNS_PRIOR, bp); atomicAdd(a.scal + S_BERNS_SELF, bs); }
}
#endif  // PCMF_COMMON_KERNELS

// ------------------------------------------------------------------------------------------------------------
// Zero-inflation, dense passes.  prob_D is never stored: P_ij = 1 where X_ij != 0 (bitmap), otherwise
// expit(logit(prior_D_j) - EU_i . EVS_j) from the triple saved when update_variational_zi_param ran
// (zi...cpp:346-367, zi_sparse...:389-410); whole-gene shortcuts for prior_D in {0,1} are per-gene flags.
// ------------------------------------------------------------------------------------------------------------
// per-gene preparation right before pass A: c_j = logit(prior_D_j), flag, nnz weight colw_j and the lgamma term
struct ZiPrepArgs {
    int32_t p; int init;
    const double *prior_D, *Lg;
    double *cz, *colw; int32_t *flag;
    double *scal;
};
#ifdef PCMF_COMMON_KERNELS
__global__ void k_zi_prepare(ZiPrepArgs a) {
    __shared__ double sh[32];
    double lgx = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.p; j += gridDim.x * blockDim.x) {
        int f = 0; double c, w = 1.0;
        if (a.init) { c = -1e300; }                    // init_zi_param: prob_D = (X > 0)  (zi...cpp:98-105)
        else {
            const double pr = a.prior_D[j];
            if (pr == 1.0) f = 1; else if (pr == 0.0) { f = 2; w = 0.0; }
            c = logit_clamped(pr);
        }
        a.cz[j] = c; a.flag[j] = f; a.colw[j] = w;
        lgx += w * a.Lg[j];                            // sum_ij prob_D_ij lgamma(X_ij + 1)  (zi...cpp:207-212)
    }
    lgx = block_sum(lgx, sh);
    if (threadIdx.x == 0) atomicAdd(a.scal + S_LGX, lgx);
}
#endif  // PCMF_COMMON_KERNELS

struct ZiAArgs {
    int64_t n; int32_t p; int K;
    const double *EU;            // n x KP (new)
    const double *EVS;           // p x KP (new)
    const double *cz; const int32_t *flag;
    const uint32_t *bitT;        // [ceil(p/32)][n]: bit c of word (jb, i) <-> X(i, 32 jb + c) != 0
    double *PV;                  // n x KP out: prob_D * EVS  (next iteration's a2, zi...cpp:336)
    double *colsumP;             // p (atomic): colsum(prob_D) -> prior_D (zi...cpp:370-372)
    double *zsc;                 // 3 doubles (atomic): sum prob_D o UVt, sum P log P + (1-P) log(1-P), [want_ll] zero part of the
                                 //   log-likelihood monitor   (all-reduced later)
    double *pack;                // scratch: fragment-ordered gene records for the DMMA kernel (zi_mma.cuh)
    uint32_t *pstore;            // or null: prob_D kept for the next iteration's gene pass, 32-bit fixed point in 8 x 8 tiles
                                 //   [ceil(n/8)][ceil(p/8)][gene in tile][cell in tile]  (zi_mma.cuh, k_zi_genes_stored)
    int want_ll;                 // the pass also evaluates the zero part of the log-likelihood monitor (zsc[2])
};
// Pass A: one thread per cell, genes streamed through shared memory in tiles of TJ, two genes per step.
// Per (cell, gene): lambda = EU_i . EVS_j (K DFMA), P = expit(c_j - lambda) (fast_exp + fast_rcp, ~22 FP64 ops),
// PV_i += P EVS_j (K DFMA).  The Bernoulli self term sum P log P + (1-P) log(1-P) = sum (1-P)(-z) - log(1+e^-z) keeps
// a running product of (1+e^-z) and takes one log per ~hundred entries.
template <int KP, int TJ>
__global__ void __launch_bounds__(128, 2) k_zi_pass_cells(ZiAArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_evs = reinterpret_cast<double *>(smem_raw);          // TJ x KP
    double *s_c = s_evs + TJ * KP;                                 // TJ
    double *s_cs = s_c + TJ;                                       // 4 x TJ per-warp gene sums
    double *s_P = s_cs + 4 * TJ;                                   // 4 x 32 x 33 transposition buffer
    int *s_f = reinterpret_cast<int *>(s_P + 4 * 32 * 33);         // TJ
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    double u[KP], pv[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) { u[k] = 0.0; pv[k] = 0.0; }
    if (live) load_vec<KP>(a.EU + i * KP, u);
    double sum_pl = 0, ent_lin = 0, ent_log = 0, prod = 1.0;
    for (int j0 = 0; j0 < a.p; j0 += TJ) {
        __syncthreads();
        for (int t = threadIdx.x; t < TJ * KP; t += blockDim.x) {
            const int jj = j0 + t / KP;
            s_evs[t] = jj < a.p ? a.EVS[(int64_t)j0 * KP + t] : 0.0;
        }
        for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
            const int jj = j0 + t;
            s_c[t] = jj < a.p ? a.cz[jj] : 0.0;
            s_f[t] = jj < a.p ? a.flag[jj] : 3;
        }
        for (int t = threadIdx.x; t < 4 * TJ; t += blockDim.x) s_cs[t] = 0.0;
        __syncthreads();
        for (int jb = 0; jb < TJ; jb += 32) {
            if (j0 + jb >= a.p) break;
            const uint32_t bits = live ? a.bitT[(int64_t)((j0 + jb) >> 5) * a.n + i] : 0u;
#pragma unroll 1
            for (int c = 0; c < 32; c += 2) {
                double v0[KP], v1[KP];
                const double2 *e0 = reinterpret_cast<const double2 *>(s_evs + (jb + c) * KP);
                const double2 *e1 = reinterpret_cast<const double2 *>(s_evs + (jb + c + 1) * KP);
                double l0[4] = {0.0, 0.0, 0.0, 0.0}, l1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
                    const double2 t0 = e0[k / 2], t1 = e1[k / 2];
                    v0[k] = t0.x; v0[k + 1] = t0.y; v1[k] = t1.x; v1[k + 1] = t1.y;
                    l0[k & 2] = fma(u[k], t0.x, l0[k & 2]); l0[(k & 2) + 1] = fma(u[k + 1], t0.y, l0[(k & 2) + 1]);
                    l1[k & 2] = fma(u[k], t1.x, l1[k & 2]); l1[(k & 2) + 1] = fma(u[k + 1], t1.y, l1[(k & 2) + 1]);
                }
                const double lam0 = (l0[0] + l0[1]) + (l0[2] + l0[3]), lam1 = (l1[0] + l1[1]) + (l1[2] + l1[3]);
                const int f0 = live ? s_f[jb + c] : 3, f1 = live ? s_f[jb + c + 1] : 3;
                const double z0 = s_c[jb + c] - lam0, z1 = s_c[jb + c + 1] - lam1;
                double y0, y1;
                bool in0, in1;
                double p0 = fast_expit(z0, y0, in0), p1 = fast_expit(z1, y1, in1);
                // exact entries: X != 0 -> 1 (zi...cpp:359-360); whole-gene shortcuts (zi...cpp:353-356)
                const bool nz0 = (bits >> c) & 1u, nz1 = (bits >> (c + 1)) & 1u;
                p0 = (f0 == 0) ? (nz0 ? 1.0 : p0) : ((f0 == 1) ? 1.0 : 0.0);
                p1 = (f1 == 0) ? (nz1 ? 1.0 : p1) : ((f1 == 1) ? 1.0 : 0.0);
                // -bernoulli_loglike(prob_D, prob_D) over 0 < P < 1 (1/(1+e^-z) is strictly inside (0,1) unless clamped)
                const bool g0 = (f0 == 0) && !nz0 && in0;
                const bool g1 = (f1 == 0) && !nz1 && in1;
                ent_lin = fma(g0 ? (1.0 - p0) : 0.0, -z0, ent_lin);
                ent_lin = fma(g1 ? (1.0 - p1) : 0.0, -z1, ent_lin);
                prod *= (g0 ? y0 : 1.0) * (g1 ? y1 : 1.0);
                sum_pl = fma(p0, lam0, sum_pl);                    // sum prob_D o UVt  (zi...cpp:182)
                sum_pl = fma(p1, lam1, sum_pl);
#pragma unroll
                for (int k = 0; k < KP; ++k) pv[k] = fma(p1, v1[k], fma(p0, v0[k], pv[k]));
                s_P[(w * 32 + c) * 33 + lane] = p0;
                s_P[(w * 32 + c + 1) * 33 + lane] = p1;
                if ((c & 6) == 6 && prod > 1e100) { ent_log += log(prod); prod = 1.0; }   // (1+e^30)^8 < 1e105
            }
            __syncwarp();
            double cs = 0.0;   // lane l sums gene (jb + l) over the warp's 32 cells
#pragma unroll 8
            for (int r = 0; r < 32; ++r) cs += s_P[(w * 32 + lane) * 33 + r];
            __syncwarp();
            s_cs[w * TJ + jb + lane] += cs;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
            if (j0 + t < a.p) atomicAdd(a.colsumP + j0 + t, s_cs[t] + s_cs[TJ + t] + s_cs[2 * TJ + t] + s_cs[3 * TJ + t]);
        }
    }
    if (live) {
        double2 *o = reinterpret_cast<double2 *>(a.PV + i * KP);
#pragma unroll
        for (int k = 0; k < KP; k += 2) o[k / 2] = make_double2(pv[k], pv[k + 1]);
    }
    double ent = ent_lin - (ent_log + log(prod));
    sum_pl = block_sum(sum_pl, sh); ent = block_sum(ent, sh);
    if (threadIdx.x == 0) { atomicAdd(a.zsc + 0, sum_pl); atomicAdd(a.zsc + 1, ent); }
}

struct ZiBArgs {
    int64_t n; int32_t p; int K;
    int64_t rows_per_chunk;
    const double *EUz;           // n x KP: EU when prob_D was defined
    const double *EUn;           // n x KP: EU_new
    const double *EVSz;          // p x KP
    const double *cz; const int32_t *flag;
    const uint32_t *bitB;        // [ceil(n/32)][p]: bit r of word (ib, j) <-> X(32 ib + r, j) != 0
    double *tmpMat;              // p x KP (atomic): prob_D^T EU_new (zi...cpp:341, zi_sparse...:369)
    double *pack;                // scratch: fragment-ordered cell records for the DMMA kernel (zi_mma.cuh)
    const uint32_t *pstore;      // or null: prob_D as the last cell pass stored it (ZiAArgs::pstore) -> no recomputation
};
// Pass B: one thread per TWO genes (j and j + 128), cells streamed through shared memory in tiles of TI (multiple
// of 32); every broadcast shared-memory load of a cell's EU feeds four DFMAs.
template <int KP, int TI>
__global__ void __launch_bounds__(128, 2) k_zi_pass_genes(ZiBArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_uz = reinterpret_cast<double *>(smem_raw);   // TI x KP
    double *s_un = s_uz + TI * KP;                         // TI x KP
    const int jA = blockIdx.x * 256 + threadIdx.x, jB = jA + 128;
    const bool liveA = jA < a.p, liveB = jB < a.p;
    double vA[KP], vB[KP], accA[KP], accB[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) { vA[k] = 0.0; vB[k] = 0.0; accA[k] = 0.0; accB[k] = 0.0; }
    double cA = 0.0, cB = 0.0; int fA = 3, fB = 3;
    if (liveA) { load_vec<KP>(a.EVSz + (int64_t)jA * KP, vA); cA = a.cz[jA]; fA = a.flag[jA]; }
    if (liveB) { load_vec<KP>(a.EVSz + (int64_t)jB * KP, vB); cB = a.cz[jB]; fB = a.flag[jB]; }
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_chunk;
    const int64_t r1 = min(a.n, r0 + a.rows_per_chunk);
    for (int64_t i0 = r0; i0 < r1; i0 += TI) {
        __syncthreads();
        for (int t = threadIdx.x; t < TI * KP; t += blockDim.x) {
            const int64_t ii = i0 + t / KP;
            const bool ok = ii < r1;
            s_uz[t] = ok ? a.EUz[i0 * KP + t] : 0.0;
            s_un[t] = ok ? a.EUn[i0 * KP + t] : 0.0;
        }
        __syncthreads();
        for (int ib = 0; ib < TI; ib += 32) {
            if (i0 + ib >= r1) break;
            const int64_t wrow = (i0 + ib) >> 5;
            const uint32_t bitsA = liveA ? a.bitB[wrow * a.p + jA] : 0u;
            const uint32_t bitsB = liveB ? a.bitB[wrow * a.p + jB] : 0u;
#pragma unroll 2
            for (int r = 0; r < 32; ++r) {
                const double2 *uz = reinterpret_cast<const double2 *>(s_uz + (ib + r) * KP);
                const double2 *un = reinterpret_cast<const double2 *>(s_un + (ib + r) * KP);
                double la0 = 0.0, la1 = 0.0, lb0 = 0.0, lb1 = 0.0;
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
                    const double2 t = uz[k / 2];
                    la0 = fma(vA[k], t.x, la0); la1 = fma(vA[k + 1], t.y, la1);
                    lb0 = fma(vB[k], t.x, lb0); lb1 = fma(vB[k + 1], t.y, lb1);
                }
                double y;
                bool in_;
                double pA = fast_expit(cA - (la0 + la1), y, in_), pB = fast_expit(cB - (lb0 + lb1), y, in_);
                pA = (fA == 0) ? (((bitsA >> r) & 1u) ? 1.0 : pA) : ((fA == 1) ? 1.0 : 0.0);
                pB = (fB == 0) ? (((bitsB >> r) & 1u) ? 1.0 : pB) : ((fB == 1) ? 1.0 : 0.0);
                if (i0 + ib + r >= r1) { pA = 0.0; pB = 0.0; }
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
                    const double2 t = un[k / 2];
                    accA[k] = fma(pA, t.x, accA[k]); accA[k + 1] = fma(pA, t.y, accA[k + 1]);
                    accB[k] = fma(pB, t.x, accB[k]); accB[k + 1] = fma(pB, t.y, accB[k + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (k < a.K) {
            if (liveA) atomicAdd(a.tmpMat + (int64_t)jA * KP + k, accA[k]);
            if (liveB) atomicAdd(a.tmpMat + (int64_t)jB * KP + k, accB[k]);
        }
    }
}

// update_prior_gamma_param (gap...cpp:541-548) when the rows of alpha / beta DIFFER (only reachable from user-supplied
// matrices, gap...cpp:126-132): alpha1(i,k) = digammaInv(log alpha2_old(i,k) + colmean(ElogU)_k), alpha2 = alpha1 / colmean(EU)_k,
// one thread per element, plus the Gamma term of the ELBO (likelihood.h:108-113), which has no closed form then.
#ifdef PCMF_COMMON_KERNELS
__global__ void k_hyper_rows(int64_t rows, int K, int KP, double count, int update, double *h1, double *h2, const double *cs_e,
                             const double *cs_elog, const double *E, const double *Elog, double *gout) {
    __shared__ double sh[32];
    double g = 0;
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % KP);
        if (k >= K) continue;
        double al1 = h1[o], al2 = h2[o];
        if (update) {
            al1 = digamma_inv(log(al2) + cs_elog[k] / count, 6);
            al2 = al1 / (cs_e[k] / count);
            h1[o] = al1; h2[o] = al2;
        }
        g += al1 * log(al2) - lgamma(al1) + (al1 - 1.0) * Elog[o] - al2 * E[o];
    }
    g = block_sum(g, sh);
    if (threadIdx.x == 0) atomicAdd(gout, g);
}
#endif  // PCMF_COMMON_KERNELS

// ------------------------------------------------------------------------------------------------------------
// Hyper-parameter update + iteration scalars, one block:
//   update_prior_gamma_param (gap...cpp:541-548) with internal::digammaInv(., 6); update_prior_zi_param
//   (zi...cpp:370-372); closed-form Gamma ELBO terms (likelihood.h:108-113 summed over identical rows);
//   gap_between_iterates (gap...cpp:255-262).
// ------------------------------------------------------------------------------------------------------------
struct HyperArgs {
    int64_t n_global; int32_t p; int K; int zi, sparse;
    double *alpha1, *alpha2, *beta1, *beta2;   // K
    const double *cs_eu, *cs_elog, *cs_ev, *cs_elogv, *cs_evs;
    const double *colsumP; double *prior_D;
    const double *usc;     // gap_d, gap_o, entropy of U (all-reduced)
    const double *zsc;     // sum prob_D o UVt, Bernoulli self term (all-reduced; ZI only)
    double lgx_const;      // sum_ij lgamma(X_ij + 1) for the non-ZI models
    double *scal;
    int update;            // 0: keep alpha / beta as they are (init_hyper_param alone runs no update_hyper_param)
    const double *rows_g;  // or null.  Hyper-parameter ROWS that differ (k_hyper_rows did the update): [0] Gamma term of U
                           //   (all-reduced), [1] of V; the K-vectors are not used then
};
#ifdef PCMF_COMMON_KERNELS
__global__ void __launch_bounds__(256) k_hyper(HyperArgs a) {
    __shared__ double sh[32];
    double gU = 0, gV = 0, closed = 0, bd = 0;
    if (threadIdx.x < a.K) {
        const int k = threadIdx.x;
        const double n = (double)a.n_global, p = (double)a.p;
        const bool upd = a.update && !a.rows_g;
        const double al1 = upd ? digamma_inv(log(a.alpha2[k]) + a.cs_elog[k] / n, 6) : a.alpha1[k];
        const double al2 = upd ? al1 / (a.cs_eu[k] / n) : a.alpha2[k];
        a.alpha1[k] = al1; a.alpha2[k] = al2;
        const double be1 = upd ? digamma_inv(log(a.beta2[k]) + a.cs_elogv[k] / p, 6) : a.beta1[k];
        const double be2 = upd ? be1 / (a.cs_ev[k] / p) : a.beta2[k];
        a.beta1[k] = be1; a.beta2[k] = be2;
        gU = n * (al1 * log(al2) - lgamma(al1)) + (al1 - 1.0) * a.cs_elog[k] - al2 * a.cs_eu[k];
        gV = p * (be1 * log(be2) - lgamma(be1)) + (be1 - 1.0) * a.cs_elogv[k] - be2 * a.cs_ev[k];
        closed = a.cs_eu[k] * a.cs_evs[k];     // sum_ij UVt = sum_k colsum(EU)_k colsum(EVS)_k  (gap...cpp:298)
    }
    if (a.zi) {
        const double n = (double)a.n_global;
        for (int j = threadIdx.x; j < a.p; j += blockDim.x) {
            const double c = a.colsumP[j];
            const double pr = c / n;
            a.prior_D[j] = pr;
            // bernoulli_loglike_vec(prob_D, prior_D), intended semantics
            if (pr > 0.0 && pr < 1.0) bd += c * log(pr) + (n - c) * log(1.0 - pr);
        }
    }
    gU = block_sum(gU, sh); gV = block_sum(gV, sh); closed = block_sum(closed, sh); bd = block_sum(bd, sh);
    if (threadIdx.x == 0) {
        if (a.rows_g) { gU = a.rows_g[0]; gV = a.rows_g[1]; }
        double *s = a.scal;
        s[S_GAMMAU] = gU; s[S_GAMMAV] = gV; s[S_SUMUVT_CLOSED] = closed; s[S_BERND_PRIOR] = bd;
        s[S_GAPU_D] = a.usc[0]; s[S_GAPU_O] = a.usc[1]; s[S_ENTU] = a.usc[2];
        if (a.zi) { s[S_SUM_PUVT] = a.zsc[0]; s[S_BERND_SELF] = a.zsc[1]; s[S_LL_ZERO] = a.zsc[2]; }
        else s[S_LGX] = a.lgx_const;
        const double diff = sqrt(s[S_GAPU_D] + s[S_GAPV_D]);
        const double nrm = sqrt(s[S_GAPU_O] + s[S_GAPV_O]);
        s[S_ABS_GAP] = diff;
        s[S_NORM_GAP] = diff / nrm;
        double e = gU + s[S_ENTU] + gV + s[S_ENTV];
        if (a.sparse) e += s[S_BERNS_PRIOR] - s[S_BERNS_SELF];
        if (a.zi) e += bd - s[S_BERND_SELF] - s[S_SUM_PUVT];
        else e -= closed;
        e -= s[S_LGX];
        s[S_ELBO_NODATA] = e;
    }
}
#endif  // PCMF_COMMON_KERNELS

// ------------------------------------------------------------------------------------------------------------
// one-off preprocessing kernels
// ------------------------------------------------------------------------------------------------------------
// occupancy bitmaps for the dense ZI passes, from CSR
#ifdef PCMF_COMMON_KERNELS
__global__ void k_build_bitmaps(int64_t n, int32_t p, const int64_t *rowptr, const int32_t *colidx, uint32_t *bitT,
                                uint32_t *bitB) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nw) {
        for (int64_t e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
            const int j = colidx[e];
            atomicOr(bitT + (int64_t)(j >> 5) * n + i, 1u << (j & 31));
            atomicOr(bitB + (i >> 5) * (int64_t)p + j, 1u << (i & 31));
        }
    }
}
#endif  // PCMF_COMMON_KERNELS
// per-gene constants: sum_i lgamma(X_ij + 1) (ELBO, gap...cpp:319-323), nnz count and sum of the gene (init, freq_D),
// sum_i X_ij log X_ij - X_ij (poisson_loglike(X, X) of the deviance monitors, likelihood.h:149-155)
template <typename VT>
__global__ void k_gene_constants(int32_t p, const int64_t *colptr, const void *val_, double *Lg, double *colsumX,
                                 double *colnnz, double *xlogx) {
    const VT *val = static_cast<const VT *>(val_);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < p; j += nw) {
        double lg = 0, sx = 0, cn = 0, xl = 0;
        for (int64_t e = colptr[j] + lane; e < colptr[j + 1]; e += 32) {
            const double x = (double)val[e];
            lg += lgamma(x + 1.0); sx += x; cn += (x != 0.0) ? 1.0 : 0.0;
            xl += x * custom_log(x) - x;
        }
        lg = warp_sum(lg); sx = warp_sum(sx); cn = warp_sum(cn); xl = warp_sum(xl);
        if (lane == 0) { Lg[j] = lg; colsumX[j] = sx; colnnz[j] = cn; xlogx[j] = xl; }
    }
}
template <typename VT>
__global__ void k_row_sums(int64_t n, const int64_t *rowptr, const void *val_, double *rowsumX) {
    const VT *val = static_cast<const VT *>(val_);
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nw) {
        double s = 0;
        for (int64_t e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) s += (double)val[e];
        s = warp_sum(s);
        if (lane == 0) rowsumX[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Monitors and factor ordering.  Everything the reference evaluates on the dense n x p rate
// matrix UVt (or prob_D o UVt) splits into a sum over the non-zeros of X -- sum x clog(rate) -- and the total rate,
// which has a closed form (gap...cpp:298) or comes out of the dense ZI pass (sum prob_D o UVt).
// ------------------------------------------------------------------------------------------------------------
// Non-zero pass over CSC, one block per gene: lambda_ij = EU_i . EVS_j.
//   out[0] += w_j sum_i x clog(w_j lambda)    rate prob_D o UVt of deviance / exp_deviance (zi...cpp:218-245); w_j is
//                                             prob_D on the gene's non-zeros (1, or 0 under the prior_D == 0 shortcut)
//   out[1] += sum_i x clog(lambda)            rate UVt of loglikelihood (gap...cpp:278-284, likelihood.h:197-212)
//   out[2] += sum_i lambda                    non-zero part of sum UVt
//   out[3] += #non-zeros with w_j = 0         log(prob_D) = -inf in zi_poisson_loglike (likelihood.h:206)
struct LamArgs {
    int32_t p; int K;
    const int64_t *colptr; const int32_t *rowidx; const void *val;
    const double *EU;        // n x KP
    const double *EVS;       // p x KP
    const int32_t *flag;     // p or null (non-ZI)
    double *out;             // 4 doubles (atomic)
};
template <int KP, typename VT>
__global__ void __launch_bounds__(128) k_gene_loglam(LamArgs a) {
    __shared__ double sh[32];
    const VT *__restrict__ val = static_cast<const VT *>(a.val);
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int j = blockIdx.x; j < a.p; j += gridDim.x) {
        double v[KP];
        load_vec<KP>(a.EVS + (int64_t)j * KP, v);
        const double w = (a.flag && a.flag[j] == 2) ? 0.0 : 1.0;
        double s1 = 0, s2 = 0;
        const int64_t e0 = a.colptr[j], e1 = a.colptr[j + 1];
        for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const int64_t i = __ldg(a.rowidx + e);
            const double x = (double)__ldg(val + e);
            double u[KP];
            load_vec<KP>(a.EU + i * KP, u);
            double l0 = 0.0, l1 = 0.0;
#pragma unroll
            for (int k = 0; k < KP; k += 2) { l0 = fma(u[k], v[k], l0); l1 = fma(u[k + 1], v[k + 1], l1); }
            const double lam = l0 + l1;
            s1 += x * custom_log(lam);
            s2 += lam;
        }
        t0 += w * s1; t1 += s1; t2 += s2;
        if (w == 0.0 && threadIdx.x == 0) t3 += (double)(e1 - e0);
    }
    t0 = block_sum(t0, sh); t1 = block_sum(t1, sh); t2 = block_sum(t2, sh); t3 = block_sum(t3, sh);
    if (threadIdx.x == 0) { atomicAdd(a.out + 0, t0); atomicAdd(a.out + 1, t1); atomicAdd(a.out + 2, t2); atomicAdd(a.out + 3, t3); }
}

// One round of the greedy factor ordering (model.cpp:137-187): for every factor c not chosen yet,
//   out[c] += sum_nnz x clog(w_j (base_ij + EU_ic EVS_jc)),   base_ij = sum_{k chosen} EU_ik EVS_jk
// which is the non-zero part of poisson_loglike(X, rate restricted to chosen + c) in partial_deviance
// (gap...cpp:420-450, zi...:264-286, sparse...:294-319, zi_sparse...:288-313).
struct OrderArgs {
    int32_t p; int K;
    const int64_t *colptr; const int32_t *rowidx; const void *val;
    const double *EU, *EVS; const int32_t *flag;
    const double *chosen;    // KP doubles: 1 when the factor is already chosen
    double *out;             // KP doubles (atomic)
    double *bound;           // FAST only: KP doubles (atomic), sum x (|log| + 1) per candidate -> error bound of out
    const float *EUf, *EVSf; // FAST only: FP32 copies of EU / EVS (half the bytes of the per-round gather of the E[U] rows)
};
// FAST: the K (K + 1) / 2 logarithms per non-zero of the greedy ordering in FP32 (lg2.approx) on FP32 copies of EU and EVS, with
// a rigorous error bound next to every score: a product u_k v_k carries three roundings (1.8e-7 relative), the sum over the
// chosen factors is taken in FP64 and rounded once (2.4e-7), the argument base + u_k v_k once more (3e-7, all terms >= 0);
// __logf <= 2^-21.4 absolute near 1 and 3 ulp elsewhere, so |l_fp32 - l| <= 6.6e-7 (|l| + 1) and |out_k - exact| <= 6.6e-7
// bound_k.  The host accepts a round's winner only when its score beats every other candidate's by more than the two bounds
// (with a factor 2 to spare) and otherwise repeats the round with the FP64 kernel, so the order is the reference's either way.
template <int KP, typename VT, bool FAST>
__global__ void __launch_bounds__(128, (FAST && KP <= 20) ? 3 : 1) k_gene_order(OrderArgs a) {
    __shared__ double sacc[KP];
    __shared__ double sbnd[KP];
    __shared__ double sch[KP];
    __shared__ __align__(16) float s_vf[KP];
    const VT *__restrict__ val = static_cast<const VT *>(a.val);
    if (threadIdx.x < KP) { sacc[threadIdx.x] = 0.0; sbnd[threadIdx.x] = 0.0; sch[threadIdx.x] = threadIdx.x < a.K ? a.chosen[threadIdx.x] : 1.0; }
    __syncthreads();
    double acc[KP];
    float accb[FAST ? KP : 1];
#pragma unroll
    for (int k = 0; k < KP; ++k) { acc[k] = 0.0; if (FAST) accb[k] = 0.f; }
    for (int j = blockIdx.x; j < a.p; j += gridDim.x) {
        if (a.flag && a.flag[j] == 2) continue;     // rate 0 on the whole gene: clog(0) = 0
        const int64_t e0 = a.colptr[j], e1 = a.colptr[j + 1];
        if (FAST) {
            // the gene's row in shared memory (registers are what limits the prefetch depth below)
            __syncthreads();
            if (threadIdx.x < KP) s_vf[threadIdx.x] = a.EVSf[(int64_t)j * KP + threadIdx.x];
            __syncthreads();
            // the loop is a dependent chain index -> E[U] row -> arithmetic per entry and the kernel waits on it (ncu: 51 % long
            // scoreboard at 12 warps per SM): the rows of the next two entries and the index of the third are in flight
            // (D rows per thread: 3 up to K = 20, 2 up to 32, 1 beyond -- the rows live in registers)
            constexpr int D = KP <= 20 ? 3 : (KP <= 32 ? 2 : 1);
            const int64_t step = blockDim.x;
            int64_t e = e0 + threadIdx.x;
            float4 r[D][KP / 4];
            double xs[D + 1];
            int64_t i_far = 0;
#pragma unroll
            for (int d = 0; d <= D; ++d) xs[d] = 0.0;
            auto fetch_row = [&](int64_t i, float4 (&dst)[KP / 4]) {
#pragma unroll
                for (int q4 = 0; q4 < KP / 4; ++q4) dst[q4] = __ldg(reinterpret_cast<const float4 *>(a.EUf + i * KP) + q4);
            };
#pragma unroll
            for (int d = 0; d < D - 1; ++d)
                if (e + d * step < e1) { xs[d] = (double)__ldg(val + e + d * step); fetch_row(__ldg(a.rowidx + e + d * step), r[d]); }
            if (e + (D - 1) * step < e1) { xs[D - 1] = (double)__ldg(val + e + (D - 1) * step); i_far = __ldg(a.rowidx + e + (D - 1) * step); }
            for (; e < e1; e += step) {
                if (e + (D - 1) * step < e1) fetch_row(i_far, r[D - 1]);
                if (e + D * step < e1) { i_far = __ldg(a.rowidx + e + D * step); xs[D] = (double)__ldg(val + e + D * step); }
                float4 (&r0)[KP / 4] = r[0];
                const double x0 = xs[0];
                const double x = x0;
                float pf[KP];
#pragma unroll
                for (int q4 = 0; q4 < KP / 4; ++q4) {
                    const float4 t = r0[q4];
                    const float4 vv = *reinterpret_cast<const float4 *>(s_vf + 4 * q4);
                    pf[4 * q4] = t.x * vv.x; pf[4 * q4 + 1] = t.y * vv.y; pf[4 * q4 + 2] = t.z * vv.z; pf[4 * q4 + 3] = t.w * vv.w;
                }
                double base = 0.0;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (sch[k] != 0.0) base += (double)pf[k];      // (uniform branch; pads have pf = 0)
                const float bf = (float)base, xf = (float)x;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (sch[k] == 0.0) {
                        const float arg = bf + pf[k];
                        const float l = arg > 0.f ? __logf(arg) : 0.f;
                        acc[k] = fma(x, (double)l, acc[k]);
                        accb[k] = fmaf(xf, fabsf(l) + 1.f, accb[k]);
                    }
#pragma unroll
                for (int d = 0; d + 1 < D; ++d)
#pragma unroll
                    for (int q4 = 0; q4 < KP / 4; ++q4) r[d][q4] = r[d + 1][q4];
#pragma unroll
                for (int d = 0; d < D; ++d) xs[d] = xs[d + 1];
            }
        } else {
            double v[KP];
            load_vec<KP>(a.EVS + (int64_t)j * KP, v);
            for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
                const int64_t i = __ldg(a.rowidx + e);
                const double x = (double)__ldg(val + e);
                double u[KP];
                load_vec<KP>(a.EU + i * KP, u);
                double base = 0.0;
#pragma unroll
                for (int k = 0; k < KP; ++k) { u[k] *= v[k]; base = fma(u[k], sch[k], base); }
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (sch[k] == 0.0) acc[k] = fma(x, custom_log(base + u[k]), acc[k]);
            }
        }
        if (FAST) {     // the FP32 bound sums are moved to FP64 once per gene (a gene's share of a thread is a few hundred entries)
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const double t = warp_sum((double)accb[k]);
                if ((threadIdx.x & 31) == 0 && t != 0.0) atomicAdd(&sbnd[k], t);
                accb[k] = 0.f;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const double t = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0 && t != 0.0) atomicAdd(&sacc[k], t);
    }
    __syncthreads();
    if (threadIdx.x < a.K && sacc[threadIdx.x] != 0.0) atomicAdd(a.out + threadIdx.x, sacc[threadIdx.x]);
    if (FAST && threadIdx.x < a.K && sbnd[threadIdx.x] != 0.0) atomicAdd(a.bound + threadIdx.x, sbnd[threadIdx.x]);
}

#ifdef PCMF_COMMON_KERNELS
__global__ void k_to_float(int64_t total, const double *__restrict__ src, float *__restrict__ dst) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) dst[o] = (float)src[o];
}
// likelihood::gamma_loglike(X, p1, p2) summed (likelihood.h:85-88): p1 log p2 + (p1 - 1) log X - p2 X - lgamma(p1)
__global__ void k_gamma_loglike(int64_t rows, int K, int KP, const double *X, const double *p1, const double *p2, double *out) {
    __shared__ double sh[32];
    double t = 0;
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        if ((int)(o % KP) >= K) continue;
        t += p1[o] * log(p2[o]) + (p1[o] - 1.0) * log(X[o]) - p2[o] * X[o] - lgamma(p1[o]);
    }
    t = block_sum(t, sh);
    if (threadIdx.x == 0) atomicAdd(out, t);
}
// out[k] += sum_r A(r,k) B(r,k)
__global__ void k_dot_cols(int64_t rows, int K, int KP, const double *A, const double *B, double *out) {
    __shared__ double sacc[64];
    if (threadIdx.x < 64) sacc[threadIdx.x] = 0.0;
    __syncthreads();
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % KP);
        if (k < K) atomicAdd(&sacc[k], A[o] * B[o]);
    }
    __syncthreads();
    if (threadIdx.x < K) atomicAdd(out + threadIdx.x, sacc[threadIdx.x]);
}
// dst(:, c) = src(:, order[c]) -- Eigen's `M *= perm` with perm.indices()[c] = order[c] (gap...cpp:344-366)
template <typename T>
__global__ void k_permute_cols(int64_t rows, int K, int KP, const int32_t *order, const T *src, T *dst) {
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % KP);
        dst[o] = k < K ? src[o - k + order[k]] : src[o];
    }
}
#endif  // PCMF_COMMON_KERNELS

// ------------------------------------------------------------------------------------------------------------
// Warm start from an explicit dense prob_D (init_zi_param(prob_D, prior_D), zi_gap_factor_model.cpp:110-114).  The
// matrix only lives until the first zero-inflation update overwrites it with the closed form, so these plain kernels
// run once or twice per fit: they replace the DMMA passes while prob_D is an arbitrary n x p matrix (column-major).
// ------------------------------------------------------------------------------------------------------------
#ifdef PCMF_COMMON_KERNELS
// values of the non-zeros weighted by prob_D (the multinomial sums are linear in x, zi...cpp:316-322), and
// sum_ij prob_D_ij lgamma(X_ij + 1) (zi...cpp:207-212)
template <typename VT>
__global__ void k_weight_vals_csr(int64_t n, const int64_t *rowptr, const int32_t *colidx, const void *val_, const double *P,
                                  double *out, double *lgx) {
    const VT *val = static_cast<const VT *>(val_);
    __shared__ double sh[32];
    double lg = 0;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nw)
        for (int64_t e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
            const double x = (double)val[e], w = P[i + (int64_t)colidx[e] * n];
            out[e] = w * x;
            lg += w * lgamma(x + 1.0);
        }
    lg = block_sum(lg, sh);
    if (threadIdx.x == 0 && lgx) atomicAdd(lgx, lg);
}
template <typename VT>
__global__ void k_weight_vals_csc(int32_t p, int64_t n, const int64_t *colptr, const int32_t *rowidx, const void *val_,
                                  const double *P, double *out) {
    const VT *val = static_cast<const VT *>(val_);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < p; j += nw)
        for (int64_t e = colptr[j] + lane; e < colptr[j + 1]; e += 32) out[e] = P[rowidx[e] + (int64_t)j * n] * (double)val[e];
}
// "pass A" on an explicit P: PV = P EVS, colsum(P), sum P o UVt, sum_{0<P<1} P log P + (1-P) log(1-P); thread per cell
__global__ void k_dense_zi_cells(int64_t n, int32_t p, int K, int KP, const double *P, const double *EU, const double *EVS,
                                 double *PV, double *colsumP, double *zsc) {
    __shared__ double sh[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    double u[64], pv[64];
    for (int k = 0; k < K; ++k) { u[k] = live ? EU[i * KP + k] : 0.0; pv[k] = 0.0; }
    double spl = 0, ent = 0;
    for (int j = 0; j < p; ++j) {
        const double pj = live ? P[i + (int64_t)j * n] : 0.0;
        double lam = 0;
        for (int k = 0; k < K; ++k) { const double v = EVS[(int64_t)j * KP + k]; lam = fma(u[k], v, lam); pv[k] = fma(pj, v, pv[k]); }
        spl = fma(pj, lam, spl);
        if (pj > 0.0 && pj < 1.0) ent += pj * log(pj) + (1.0 - pj) * log(1.0 - pj);
        const double cs = warp_sum(pj);
        if ((threadIdx.x & 31) == 0) atomicAdd(colsumP + j, cs);
    }
    if (live) for (int k = 0; k < KP; ++k) PV[i * KP + k] = k < K ? pv[k] : 0.0;
    spl = block_sum(spl, sh); ent = block_sum(ent, sh);
    if (threadIdx.x == 0) { atomicAdd(zsc + 0, spl); atomicAdd(zsc + 1, ent); }
}
// Monitors while prob_D still is the explicit matrix of a warm start (zi...cpp:160-167, :218-245; likelihood.h:149-155,
// :197-212), thread per cell, lambda = EU_i . EVS_j:
//   out[0] += sum_nnz x clog(P lambda)           deviance / exp_deviance, rate prob_D o UVt
//   out[1] += sum_nnz x clog(lambda)             zi_poisson_loglike, non-zero entries
//   out[2] += sum_nnz lambda
//   out[3] += #non-zeros with P = 0              log(prob_D) = -inf
//   out[5] += sum_zeros log(1 - P + P e^-lambda)
//   out[7] += sum_nnz log P                      (over P > 0)
template <typename VT>
__global__ void k_dense_zi_monitor(int64_t n, int32_t p, int K, int KP, const double *P, const double *EU, const double *EVS,
                                   const int64_t *rowptr, const int32_t *colidx, const void *val_, const uint32_t *bitT, double *out) {
    const VT *val = static_cast<const VT *>(val_);
    __shared__ double sh[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t5 = 0, t7 = 0;
    if (live) {
        for (int j = 0; j < p; ++j) {
            if ((bitT[(int64_t)(j >> 5) * n + i] >> (j & 31)) & 1u) continue;
            double lam = 0;
            for (int k = 0; k < K; ++k) lam = fma(EU[i * KP + k], EVS[(int64_t)j * KP + k], lam);
            const double pj = P[i + (int64_t)j * n];
            t5 += log(1.0 - pj + pj * exp(-lam));
        }
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int j = colidx[e];
            const double x = (double)val[e], pj = P[i + (int64_t)j * n];
            double lam = 0;
            for (int k = 0; k < K; ++k) lam = fma(EU[i * KP + k], EVS[(int64_t)j * KP + k], lam);
            t0 += x * custom_log(pj * lam); t1 += x * custom_log(lam); t2 += lam;
            if (pj > 0.0) t7 += log(pj); else t3 += 1.0;
        }
    }
    t0 = block_sum(t0, sh); t1 = block_sum(t1, sh); t2 = block_sum(t2, sh); t3 = block_sum(t3, sh); t5 = block_sum(t5, sh); t7 = block_sum(t7, sh);
    if (threadIdx.x == 0) {
        atomicAdd(out + 0, t0); atomicAdd(out + 1, t1); atomicAdd(out + 2, t2); atomicAdd(out + 3, t3); atomicAdd(out + 5, t5); atomicAdd(out + 7, t7);
    }
}
// "pass B" on an explicit P: tmpMat += P^T EU_new; thread per gene
__global__ void k_dense_zi_genes(int64_t n, int32_t p, int K, int KP, const double *P, const double *EUn, double *tmpMat) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    double acc[64];
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double pj = P[i + (int64_t)j * n];
        for (int k = 0; k < K; ++k) acc[k] = fma(pj, EUn[i * KP + k], acc[k]);
    }
    for (int k = 0; k < K; ++k) tmpMat[(int64_t)j * KP + k] += acc[k];
}
#endif  // PCMF_COMMON_KERNELS

#ifdef PCMF_COMMON_KERNELS
// conv_mode = 2: Gram matrices for matrix::rv_coeff (utils/matrix.h:100-104) of (A, B), rows x KP each:
// out[0..K^2) += A'B, out[K^2..2K^2) += A'A, out[2K^2..3K^2) += B'B (K <= 64); trace(A A' B B') = ||A'B||_F^2
__global__ void k_rv_gram(int64_t rows, int K, int KP, const double *A, const double *B, double *out) {
    const int kl = blockIdx.y * blockDim.y + threadIdx.y;     // (k, l) pair
    if (kl >= K * K) return;
    const int k = kl / K, l = kl % K;
    double ab = 0, aa = 0, bb = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const double ak = A[r * KP + k], al = A[r * KP + l], bk = B[r * KP + k], bl = B[r * KP + l];
        ab = fma(ak, bl, ab); aa = fma(ak, al, aa); bb = fma(bk, bl, bb);
    }
    ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
    if (threadIdx.x == 0) { atomicAdd(out + kl, ab); atomicAdd(out + K * K + kl, aa); atomicAdd(out + 2 * K * K + kl, bb); }
}
// dst = num / den element-wise (E[U] of the previous iterate from a1_old / a2_old), pads 0
__global__ void k_ratio(int64_t rows, int K, int KP, const double *num, const double *den, double *dst) {
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x)
        dst[o] = ((int)(o % KP) < K) ? num[o] / den[o] : 0.0;
}
#endif  // PCMF_COMMON_KERNELS

// column-major (R / Eigen) <-> padded row-major
#ifdef PCMF_COMMON_KERNELS
__global__ void k_colmajor_to_rowpad(int64_t rows, int K, int KP, const double *src, double *dst, double padval) {
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % KP);
        const int64_t r = o / KP;
        dst[o] = k < K ? src[r + (int64_t)k * rows] : padval;
    }
}
#endif  // PCMF_COMMON_KERNELS
#ifdef PCMF_COMMON_KERNELS
__global__ void k_rowpad_to_colmajor(int64_t rows, int K, int KP, const double *src, double *dst) {
    const int64_t total = rows * K;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = o % rows;
        const int k = (int)(o / rows);
        dst[o] = src[r * KP + k];
    }
}
#endif  // PCMF_COMMON_KERNELS
#ifdef PCMF_COMMON_KERNELS
__global__ void k_rowpad_to_colmajor_i32(int64_t rows, int K, int KP, const int32_t *src, int32_t *dst) {
    const int64_t total = rows * K;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = o % rows;
        const int k = (int)(o / rows);
        dst[o] = src[r * KP + k];
    }
}
#endif  // PCMF_COMMON_KERNELS
#ifdef PCMF_COMMON_KERNELS
__global__ void k_row_extremes(int64_t rows, int K, int KP, const double *__restrict__ M, double2 *__restrict__ ext) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        double mn = 1e300, mx = -1e300;
        for (int k = 0; k < K; ++k) { const double v = M[r * KP + k]; mn = fmin(mn, v); mx = fmax(mx, v); }
        ext[r] = make_double2(mn, mx);
    }
}
#endif  // PCMF_COMMON_KERNELS
#ifdef PCMF_COMMON_KERNELS
__global__ void k_fill(double *p, int64_t n, double v) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) p[o] = v;
}
#endif  // PCMF_COMMON_KERNELS
// dense prob_D output (n x p column-major) for the return object: ZI_param$prob_D (zi...cpp:251)
struct ProbDArgs {
    int64_t n; int32_t p; int K, KP;
    const double *EUz, *EVSz, *cz; const int32_t *flag; const uint32_t *bitT;
    double *out;
};
#ifdef PCMF_COMMON_KERNELS
__global__ void k_prob_d_dense(ProbDArgs a) {
    const int64_t total = a.n * (int64_t)a.p;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = o % a.n;
        const int j = (int)(o / a.n);
        const int f = a.flag[j];
        double pj;
        if (f == 1) pj = 1.0;
        else if (f == 2) pj = 0.0;
        else if ((a.bitT[(int64_t)(j >> 5) * a.n + i] >> (j & 31)) & 1u) pj = 1.0;
        else {
            double lam = 0.0;
            for (int k = 0; k < a.K; ++k) lam = fma(a.EUz[i * a.KP + k], a.EVSz[(int64_t)j * a.KP + k], lam);
            pj = expit_clamped(a.cz[j] - lam);
        }
        a.out[o] = pj;
    }
}
#endif  // PCMF_COMMON_KERNELS

}  // namespace pcmf
