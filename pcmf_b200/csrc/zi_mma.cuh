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

__device__ const double kPow2Tab64[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138, 1.0556451783605572,
    1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812, 1.189207115002721, 1.202156731452703,
    1.215247359980469, 1.22848053610687, 1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905,
    1.383909881963832, 1.3989796725383112, 1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, 1.5422108254079407, 1.559004400237837,
    1.5759808451078865, 1.593142151342267, 1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062, 1.7562521603732995, 1.7753764925265212,
    1.7947090750031072, 1.8142521755003989, 1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

// exp(-z) for |z| < 30 with a 64-entry table of 2^(i/64) (shared memory): m = rint(-64 z / ln 2), r = -z - m ln2/64
// (|r| <= ln2/128 = 0.0054), degree-4 Taylor (truncation r^5/120 <= 3.9e-14 relative), scaled through the exponent field.
// 8 FP64 instructions (fast_exp: 20).  Arguments outside the range give garbage that the caller discards.
// On this part the SM sub-partition's dispatch port is what the dense passes saturate (a DFMA holds it for 2 cycles, a
// DMMA for 16, nothing else issues meanwhile -- scripts/ubench_zi_core.cu), so every instruction removed here counts:
// round 1 ran a degree-5 polynomial with a two-term Cody-Waite reduction (1e-16); the second reduction term moves r by
// |m| 3.6e-19 <= 1e-15 and the fifth-order term by 3.9e-14, both far below the 1.2e-10 of the stored prob_D (PCMF_PSTORE_SCALE)
// and the 1e-8 the parity tests ask for, so they are gone (-DPCMF_EXPIT_FULL restores them).
//
// The table is REPLICATED 16 times, entry i of copy c at tab[16 i + c], and lane l reads copy l & 15: a 64-bit shared load is
// served per half-warp, and the 16 lanes of a half-warp then always touch 16 different bank pairs whatever their i -- the
// single 512-byte copy of round 1 cost ~3 wavefronts per lookup (ncu: 29 % of all shared wavefronts were bank conflicts).
constexpr int kTabRep = 16, kTabDoubles = 64 * kTabRep;
__device__ __forceinline__ double fast_exp_neg_tab(double z, const double *__restrict__ tab_lane) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    double t = fma(z, -92.33248261689366, MAGIC);
    const int m = __double2loint(t);
    t -= MAGIC;
    double r = fma(t, -0.010830424696249145, -z);
#ifdef PCMF_EXPIT_FULL
    r = fma(t, -3.623510646634843e-19, r);
    double p = 8.33333333333333333333e-03;                  // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);              // 1/4!
#else
    double p = 4.16666666666666666667e-02;                  // 1/4!
#endif
    p = fma(p, r, 1.66666666666666666667e-01);              // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // tab is PRE-BIASED (load_pow2_tab): entry i has (i << 14) taken off its high word, so that adding (m << 14) leaves
    // exactly ((m - i) << 14) = (m >> 6) << 20 on the exponent field -- one integer multiply-add instead of shift, mask, add
    const double tb = tab_lane[(m & 63) * kTabRep];
    const double sc = __hiloint2double(__double2hiint(tb) + (int)((unsigned)m << 14), __double2loint(tb));
    return p * sc;
}
// the replicated table of 2^(i/64) as the dense passes keep it in shared memory (see fast_exp_neg_tab); a thread passes
// pow2_tab_lane(s_tab) to the evaluation
__device__ __forceinline__ void load_pow2_tab(double *s_tab) {
    for (int t = threadIdx.x; t < kTabDoubles; t += blockDim.x) {
        const int i = t / kTabRep;
        const double v = kPow2Tab64[i];
        s_tab[t] = __hiloint2double(__double2hiint(v) - (i << 14), __double2loint(v));
    }
}
__device__ __forceinline__ const double *pow2_tab_lane(const double *s_tab) { return s_tab + (threadIdx.x & (kTabRep - 1)); }
// 1/y for normal y in the dense passes: MUFU.RCP64H seed (relative error e <= 2^-23) + one Newton step (error e^2 <= 1.4e-14)
__device__ __forceinline__ double fast_rcp2(double y) {
#ifdef PCMF_EXPIT_FULL
    return fast_rcp(y);
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    return fma(r, fma(-y, r, 1.0), r);
#endif
}
// internal::expit (internal.h:151-155): exactly 1 / 0 for z >= 30 / z <= -30.  Also returns y = 1 + exp(-z) (garbage when
// clamped) and the clamp flag.  Entries with X != 0 (zi...cpp:359-360) do not get a select of their own: the caller hands
// them z = +-1e300 (the value the packed record carries for the gene: +1e300, or -1e300 under the prior_D == 0 shortcut),
// which the clamp turns into exactly 1 / 0 and which takes them out of the entropy sums like every other clamped entry.
__device__ __forceinline__ double fast_expit_tab(double z, const double *__restrict__ tab_lane, double &y, bool &big, double &em) {
    const int hi = __double2hiint(z);
    big = (hi & 0x7fffffff) >= 0x403E0000;                 // |z| >= 30, also NaN / inf
    em = fast_exp_neg_tab(z, tab_lane);                    // e^-z (garbage when clamped)
    y = 1.0 + em;
    const double pr = fast_rcp2(y);
    const double sat = __hiloint2double((hi < 0) ? 0 : 0x3FF00000, 0);
    return big ? sat : pr;
}

__device__ __forceinline__ double fast_expit_tab(double z, const double *__restrict__ tab_lane, double &y, bool &big) {
    double em;
    return fast_expit_tab(z, tab_lane, y, big, em);
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// ---- bulk (TMA) copy + mbarrier plumbing --------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Fragment-ordered copy of a row-major [rows][KP] matrix, one record per group of 8 rows, so that a tile of the dense
// passes is ONE contiguous bulk copy and every fragment load is a conflict-free 256-byte shared-memory read:
//   record = [ f1: KS x 32 | f2: 2 x NT x 32 | (WITH_C) c: 8 | zf: 8 | e^-c: 8 ]
//   f1[s][lane]      = M1[8 grp + g][4 s + q]               first product, contraction over the K factors
//   f2[s'][nt][lane] = M2[8 grp + 2 q + s'][8 nt + g]       second product, contraction over the 8 rows (0 beyond KP)
//   c / zf: effective logit of the gene (+-1e300 for the whole-gene shortcuts, zi...cpp:353-356) and the z that stands for
//   its non-zero entries (+1e300 -> prob_D = 1, zi...cpp:359-360; -1e300 under the prior_D == 0 shortcut).  Rows beyond
//   `rows` are zero (c = zf = -1e300 -> P = 0).  e^-c feeds the zero part of the log-likelihood monitor when pass A
//   evaluates it on the way (k_zi_cells_mma<.., WITH_LL>).
template <int KP, bool WITH_C>
struct ZiRec {
    static constexpr int KS = KP / 4, NT = (KP + 7) / 8;
    static constexpr int F1 = KS * 32, F2 = 2 * NT * 32, REC = F1 + F2 + (WITH_C ? 24 : 0);
};
template <int KP, bool WITH_C>
__global__ void k_zi_pack(int64_t rows, const double *__restrict__ M1, const double *__restrict__ M2,
                          const double *__restrict__ cz, const int32_t *__restrict__ flag, double *__restrict__ out) {
    using R = ZiRec<KP, WITH_C>;
    const int64_t ngroups = (rows + 7) >> 3, total = ngroups * R::REC;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t grp = idx / R::REC;
        const int o = (int)(idx - grp * R::REC);
        double v = 0.0;
        if (o < R::F1) {
            const int s = o >> 5, lane = o & 31, g = lane >> 2, q = lane & 3;
            const int64_t row = grp * 8 + g;
            if (row < rows) v = M1[row * KP + 4 * s + q];
        } else if (o < R::F1 + R::F2) {
            const int o2 = o - R::F1, lane = o2 & 31, g = lane >> 2, q = lane & 3;
            const int t = o2 >> 5, sp = t / R::NT, nt = t - sp * R::NT;
            const int64_t row = grp * 8 + 2 * q + sp;
            const int k = 8 * nt + g;
            if (row < rows && k < KP) v = M2[row * KP + k];
        } else {
            const int o3 = o - R::F1 - R::F2, e = o3 & 7;
            const int64_t row = grp * 8 + e;
            const int f = row < rows ? flag[row] : 2;
            if (o3 < 8) v = (row < rows && f == 0) ? cz[row] : ((f == 1) ? 1e300 : -1e300);
            else if (o3 < 16) v = (f == 2) ? -1e300 : 1e300;
            else v = (row < rows && f == 0) ? exp(-cz[row]) : 1.0;
        }
        out[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pass A: rows = cells (8 MT per warp, fixed for the CTA), columns = genes streamed through shared memory.
// Outputs PV = prob_D . EVS (n x KP), colsum(prob_D), sum prob_D o UVt and the Bernoulli self term.
// ------------------------------------------------------------------------------------------------------------
// prob_D in 32-bit fixed point: q = rint(P (2^32 - 1)), so that 0 and 1 -- the values of every non-zero entry and of every
// clamped one -- are exact and the others carry an absolute error of 1.2e-10.
#define PCMF_PSTORE_SCALE 4294967295.0
// WITH_LL: also the zero part of likelihood::zi_poisson_loglike (likelihood.h:197-212) of the state this pass defines -- the
// log-likelihood monitor of the iteration (zi_gap_factor_model.cpp:160-167) -- from what is in registers anyway: with
// e = e^-z and lambda = c - z,  log(1 - P + P e^-lambda) = log(e + e^-c / e) - log(1 + e) = log(e^2 + e^-c) + z - log(1 + e),
// and sum log(1 + e) is the entropy's running product.  ~4 FP64 instructions per entry instead of a second dense pass.
template <int KP, int MT, int TJ, bool STORE_P, bool WITH_LL>
__global__ void __launch_bounds__(128, (KP <= 20) ? (MT <= 2 ? 4 : 3) : 2) k_zi_cells_mma(ZiAArgs a) {
    const double *__restrict__ pack = a.pack;
    using R = ZiRec<KP, true>;
    constexpr int KS = R::KS, NT = R::NT, NG = TJ / 8, BUF = NG * R::REC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_buf = reinterpret_cast<double *>(smem_raw);       // 2 x NG records, double buffered
    double *s_cs = s_buf + 2 * BUF;                             // 2 x 4 x TJ: per-warp column sums, double buffered (plain stores:
                                                                //   a shared-memory FP64 atomicAdd is a CAS spin loop)
    double *s_tab = s_cs + 8 * TJ;                              // kTabDoubles
    __shared__ double sh[32];
    __shared__ uint64_t bar[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    load_pow2_tab(s_tab);
    const double *tabl = pow2_tab_lane(s_tab);
    for (int t = threadIdx.x; t < 8 * TJ; t += blockDim.x) s_cs[t] = 0.0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncthreads();
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
    // rows beyond n (last CTA only) must contribute P = 0 everywhere: their EU fragment is 0 and the accumulator of the first
    // product starts at 1e305, above every effective logit (|c| <= 1e300), so z = c - lambda is clamped to P = 0 by the
    // expit itself -- no per-entry select on `live` in the inner loop
    double dinit[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) dinit[mt] = live[mt] ? 0.0 : 1e305;
    // stored prob_D: this lane's two words inside a tile, and the tile index of (cell group of mt = 0, gene group 0)
    // (tile index = cell group x gene groups + gene group, kept in 32 bits: the host stores prob_D only below 2^32 tiles)
    const unsigned st_ngg = (unsigned)((a.p + 7) >> 3);
    const unsigned st_tile0 = (unsigned)(cell0 >> 3) * st_ngg;
    uint32_t *const st_lane = STORE_P ? a.pstore + (2 * q) * 8 + g : nullptr;
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    double ent_lin = 0, ent_log = 0, prod = 1.0;
    double ll_lin = 0, ll_log = 0, ll_prod = 1.0;
    uint32_t bits[MT], bits_nx[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { bits[mt] = 0u; bits_nx[mt] = 0u; }

    // gene range of this CTA: blockIdx.y splits the gene tiles when there are too few cell blocks to fill the machine
    // for many waves (cell-sharded runs); PV and the scalars are then accumulated atomically over the chunks
    const int ngroups = (a.p + 7) >> 3, ntiles_all = (ngroups + NG - 1) / NG;
    const int tiles_per_chunk = (ntiles_all + (int)gridDim.y - 1) / (int)gridDim.y;
    const int tile_first = (int)blockIdx.y * tiles_per_chunk;
    const int ntiles = max(0, min(ntiles_all, tile_first + tiles_per_chunk) - tile_first);
    // every CTA walks its tiles in a rotated order (start = blockIdx.x mod ntiles): the resident CTAs then flush their
    // column sums to different genes instead of all hammering the same TJ addresses at the same moment
    const int rot = ntiles > 0 ? (int)(blockIdx.x % (unsigned)ntiles) : 0;
    auto tile_id = [&](int t) { const int u = t + rot; return tile_first + (u >= ntiles ? u - ntiles : u); };
    auto issue = [&](int t) {        // one thread: arm the barrier with the byte count, then one bulk copy
        const int g0 = tile_id(t) * NG, gn = min(NG, ngroups - g0);
        const unsigned bytes = (unsigned)(gn * R::REC * sizeof(double));
        mbar_expect_tx(&bar[t & 1], bytes);
        bulk_g2s(s_buf + (t & 1) * BUF, pack + (int64_t)g0 * R::REC, bytes, &bar[t & 1]);
    };
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
        bits_nx[mt] = (live[mt] && ntiles > 0) ? __ldg(a.bitT + (int64_t)((tile_id(0) * TJ) >> 5) * a.n + (cell0 + mt * 8 + g)) : 0u;
    if (threadIdx.x == 0 && ntiles > 0) issue(0);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int j0 = tile_id(tile) * TJ;
        const int j0_next = tile + 1 < ntiles ? tile_id(tile + 1) * TJ : a.p;
        const double *buf = s_buf + (tile & 1) * BUF;
        double *cs_buf = s_cs + (tile & 1) * 4 * TJ;
        if (threadIdx.x == 0 && tile + 1 < ntiles) issue(tile + 1);
        mbar_wait(&bar[tile & 1], (tile >> 1) & 1);
#pragma unroll 1
        for (int grp = 0; grp < NG; ++grp) {
            const int jg = j0 + grp * 8;
            if (jg >= a.p) break;
            const double *rec = buf + grp * R::REC;
            if ((grp & 3) == 0) {
                // occupancy words of this 32-gene block were requested one block ahead; request the next block now
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    bits[mt] = bits_nx[mt];
                    // next 32-gene block: in this tile if it has one (the last tile may be partial), else the next tile's first
                    const int jn = (grp + 4 < NG && jg + 32 < a.p) ? jg + 32 : j0_next;
                    bits_nx[mt] = (live[mt] && jn < a.p) ? __ldg(a.bitT + (int64_t)(jn >> 5) * a.n + (cell0 + mt * 8 + g)) : 0u;
                }
            }
            double d[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { d[mt][0] = dinit[mt]; d[mt][1] = dinit[mt]; }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = rec[s * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
            }
            const double2 c2 = *reinterpret_cast<const double2 *>(rec + R::F1 + R::F2 + 2 * q);
            const double2 n2 = *reinterpret_cast<const double2 *>(rec + R::F1 + R::F2 + 8 + 2 * q);
            double2 ec2 = make_double2(1.0, 1.0);
            if (WITH_LL) ec2 = *reinterpret_cast<const double2 *>(rec + R::F1 + R::F2 + 16 + 2 * q);
            const int sh0 = ((grp & 3) << 3) + 2 * q;
            const unsigned st_tile = st_tile0 + (unsigned)(jg >> 3);
            double cs0 = 0.0, cs1 = 0.0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const bool nz0 = (bits[mt] >> sh0) & 1u, nz1 = (bits[mt] >> (sh0 + 1)) & 1u;
                const double z0 = nz0 ? n2.x : c2.x - d[mt][0], z1 = nz1 ? n2.y : c2.y - d[mt][1];
                double y0, y1, em0, em1; bool big0, big1;
                const double p0 = fast_expit_tab(z0, tabl, y0, big0, em0), p1 = fast_expit_tab(z1, tabl, y1, big1, em1);
                if (WITH_LL) {
                    // interior entries (all of them zeros of X): log(e^2 + e^-c) + z; a zero entry with P = 1: -lambda
                    ll_prod *= (big0 ? 1.0 : fma(em0, em0, ec2.x)) * (big1 ? 1.0 : fma(em1, em1, ec2.y));
                    ll_lin += (big0 ? 0.0 : z0) + (big1 ? 0.0 : z1);
                    ll_lin -= ((big0 && !nz0 && z0 > 0.0) ? d[mt][0] : 0.0) + ((big1 && !nz1 && z1 > 0.0) ? d[mt][1] : 0.0);
                }
                // (rows beyond n: lambda starts at 1e305 and their occupancy words are 0, so P is the clamped 0 here)
                // -bernoulli_loglike(prob_D, prob_D) over 0 < P < 1: sum (1-P)(-z) - log(1 + e^-z)
                const bool g0 = !big0, g1 = !big1;
                ent_lin = fma(g0 ? (p0 - 1.0) : 0.0, z0, ent_lin);
                ent_lin = fma(g1 ? (p1 - 1.0) : 0.0, z1, ent_lin);
                prod *= (g0 ? y0 : 1.0) * (g1 ? y1 : 1.0);
                cs0 += p0; cs1 += p1;
                d[mt][0] = p0; d[mt][1] = p1;
                if (STORE_P) {   // (the store is padded to whole CTAs of cells: no test on rows beyond n)
                    // tile (cell group, gene group), element [gene 2q + s][cell g]: the accumulator layout of the gene pass
                    uint32_t *tp = st_lane + ((size_t)(st_tile + (unsigned)mt * st_ngg) << 6);
                    // rint(P (2^32 - 1)) read off the low mantissa word of 2^52 + P (2^32 - 1): one FMA per value (an F2I from
                    // FP64 cost ~20 issue cycles per entry here)
                    tp[0] = (uint32_t)__double2loint(fma(p0, PCMF_PSTORE_SCALE, 4503599627370496.0));
                    tp[8] = (uint32_t)__double2loint(fma(p1, PCMF_PSTORE_SCALE, 4503599627370496.0));
                }
            }
            if (prod > 1e100) { ent_log += log(prod); prod = 1.0; }     // (1 + e^30)^(2 MT) < 1e105 per group
            if (WITH_LL && (ll_prod > 1e60 || ll_prod < 1e-60)) { ll_log += log(ll_prod); ll_prod = 1.0; }   // (e^60 + e^28)^(2 MT) < 1e209
#pragma unroll
            for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = rec[R::F1 + (sp * NT + nt) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], d[mt][sp], b);
                }
            // column sums: reduce over the 8 row groups g (lane bits 2..4)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
            }
            if (g == 0) *reinterpret_cast<double2 *>(cs_buf + w * TJ + grp * 8 + 2 * q) = make_double2(cs0, cs1);
        }
        __syncthreads();   // every warp is done with this buffer: it may be refilled, its column sums flushed
        for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
            if (j0 + t < a.p) atomicAdd(a.colsumP + j0 + t, (cs_buf[t] + cs_buf[TJ + t]) + (cs_buf[2 * TJ + t] + cs_buf[3 * TJ + t]));
            cs_buf[t] = 0.0; cs_buf[TJ + t] = 0.0; cs_buf[2 * TJ + t] = 0.0; cs_buf[3 * TJ + t] = 0.0;
        }
    }
    // epilogue: PV, and sum_ij prob_D_ij UVt_ij = sum_ik EU_ik PV_ik (zi...cpp:182)
    double sum_pl = 0.0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        if (!live[mt]) continue;
        const int64_t i = cell0 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int k = nt * 8 + 2 * q;
            if (k < KP) {
                if (gridDim.y == 1) *reinterpret_cast<double2 *>(a.PV + i * KP + k) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                else { atomicAdd(a.PV + i * KP + k, acc[mt][nt][0]); atomicAdd(a.PV + i * KP + k + 1, acc[mt][nt][1]); }
                const double2 u = *reinterpret_cast<const double2 *>(a.EU + i * KP + k);
                sum_pl = fma(u.x, acc[mt][nt][0], fma(u.y, acc[mt][nt][1], sum_pl));
            }
        }
    }
    const double sum_log_y = ent_log + log(prod);
    double ent = ent_lin - sum_log_y;
    sum_pl = block_sum(sum_pl, sh); ent = block_sum(ent, sh);
    if (threadIdx.x == 0) { atomicAdd(a.zsc + 0, sum_pl); atomicAdd(a.zsc + 1, ent); }
    if (WITH_LL) {
        double ll = ll_lin + ll_log + log(ll_prod) - sum_log_y;
        ll = block_sum(ll, sh);
        if (threadIdx.x == 0) atomicAdd(a.zsc + 2, ll);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Zero-entry part of likelihood::zi_poisson_loglike (likelihood.h:197-212), monitor only:
//   sum over X_ij == 0 of log(1 - P_ij + P_ij exp(-lambda_ij)),  P the current prob_D, lambda = UVt.
// Same tiles as pass A (it reuses pass A's packed gene records), first product only.  With y = 1 + e^-z, P = 1/y:
//   1 - P + P e^-lambda = ((y - 1) + e^-lambda) / y;  clamped entries give -lambda (P = 1) or 0 (P = 0).
// ------------------------------------------------------------------------------------------------------------
template <int KP, int MT, int TJ>
__global__ void __launch_bounds__(128, (KP <= 20) ? 3 : 2) k_zi_loglik_mma(ZiAArgs a, double *out) {
    const double *__restrict__ pack = a.pack;
    using R = ZiRec<KP, true>;
    constexpr int KS = R::KS, NG = TJ / 8, BUF = NG * R::REC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_buf = reinterpret_cast<double *>(smem_raw);
    double *s_tab = s_buf + 2 * BUF + 2 * TJ;
    __shared__ double sh[32];
    __shared__ uint64_t bar[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    load_pow2_tab(s_tab);
    const double *tabl = pow2_tab_lane(s_tab);
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    const int64_t cell0 = (int64_t)blockIdx.x * (32 * MT) + w * (8 * MT);
    double a1[MT][KS];
    bool live[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int64_t i = cell0 + mt * 8 + g;
        live[mt] = i < a.n;
#pragma unroll
        for (int s = 0; s < KS; ++s) a1[mt][s] = live[mt] ? __ldg(a.EU + i * KP + 4 * s + q) : 0.0;
    }
    double lin = 0, lg = 0, pn = 1.0, pd = 1.0;
    uint32_t bits[MT];
    const int ngroups = (a.p + 7) >> 3, ntiles = (ngroups + NG - 1) / NG;
    auto issue = [&](int tile) {
        const int g0 = tile * NG, gn = min(NG, ngroups - g0);
        const unsigned bytes = (unsigned)(gn * R::REC * sizeof(double));
        mbar_expect_tx(&bar[tile & 1], bytes);
        bulk_g2s(s_buf + (tile & 1) * BUF, pack + (int64_t)g0 * R::REC, bytes, &bar[tile & 1]);
    };
    if (threadIdx.x == 0) issue(0);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int j0 = tile * TJ;
        const double *buf = s_buf + (tile & 1) * BUF;
        if (threadIdx.x == 0 && tile + 1 < ntiles) issue(tile + 1);
        mbar_wait(&bar[tile & 1], (tile >> 1) & 1);
#pragma unroll 1
        for (int grp = 0; grp < NG; ++grp) {
            const int jg = j0 + grp * 8;
            if (jg >= a.p) break;
            const double *rec = buf + grp * R::REC;
            if ((grp & 3) == 0) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    bits[mt] = live[mt] ? __ldg(a.bitT + (int64_t)(jg >> 5) * a.n + (cell0 + mt * 8 + g)) : 0xffffffffu;
            }
            double d[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { d[mt][0] = 0.0; d[mt][1] = 0.0; }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = rec[s * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
            }
            const double2 c2 = *reinterpret_cast<const double2 *>(rec + R::F1 + R::F2 + 2 * q);
            const int sh0 = ((grp & 3) << 3) + 2 * q;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double lam = d[mt][e], z = (e ? c2.y : c2.x) - lam;
                    const bool zero = !((bits[mt] >> (sh0 + e)) & 1u);      // dead rows carry all-ones words
                    const int hi = __double2hiint(z);
                    const bool big = (hi & 0x7fffffff) >= 0x403E0000;
                    const double em = fast_exp_neg_tab(z, tabl);               // y - 1
                    const double el = fast_exp_neg_tab(fmin(lam, 690.0), tabl);
                    const bool in_ = zero && !big;
                    pn *= in_ ? (em + el) : 1.0;
                    pd *= in_ ? (1.0 + em) : 1.0;
                    lin -= (zero && big && hi >= 0) ? lam : 0.0;
                }
            if (pn > 1e100 || pn < 1e-100) { lg += log(pn); pn = 1.0; }
            if (pd > 1e100) { lg -= log(pd); pd = 1.0; }
        }
        __syncthreads();
    }
    double t = lin + lg + log(pn) - log(pd);
    t = block_sum(t, sh);
    if (threadIdx.x == 0) atomicAdd(out, t);
}

// ------------------------------------------------------------------------------------------------------------
// Pass B: rows = genes (8 MT per warp), columns = cells of this CTA's chunk streamed through shared memory.
// Output tmpMat = prob_D^T . EU_new (p x KP), prob_D recomputed from the triple saved when it was defined.
// ------------------------------------------------------------------------------------------------------------
template <int KP, int MT, int TI>
__global__ void __launch_bounds__(128, (KP <= 20) ? (MT <= 2 ? 4 : 3) : 2) k_zi_genes_mma(ZiBArgs a) {
    const double *__restrict__ pack = a.pack;
    using R = ZiRec<KP, false>;
    constexpr int KS = R::KS, NT = R::NT, NG = TI / 8, BUF = NG * R::REC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_buf = reinterpret_cast<double *>(smem_raw);       // 2 x NG records [EUz first-product | EUn second-product]
    double *s_tab = s_buf + 2 * BUF;                            // kTabDoubles
    __shared__ uint64_t bar[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    load_pow2_tab(s_tab);
    const double *tabl = pow2_tab_lane(s_tab);
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    const int gene0 = blockIdx.x * (32 * MT) + w * (8 * MT);
    double a1[MT][KS], c[MT], zf[MT];
    bool live[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int j = gene0 + mt * 8 + g;
        live[mt] = j < a.p;
        c[mt] = -1e300; zf[mt] = -1e300;
        if (live[mt]) {
            const int f = a.flag[j];
            c[mt] = (f == 0) ? a.cz[j] : ((f == 1) ? 1e300 : -1e300);
            zf[mt] = (f == 2) ? -1e300 : 1e300;
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) a1[mt][s] = live[mt] ? __ldg(a.EVSz + (int64_t)j * KP + 4 * s + q) : 0.0;
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_chunk;        // multiple of 64
    const int64_t r1 = min(a.n, r0 + a.rows_per_chunk);
    uint32_t bits[MT], bits_nx[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        bits[mt] = 0u;
        bits_nx[mt] = (live[mt] && r0 < r1) ? __ldg(a.bitB + (r0 >> 5) * a.p + (gene0 + mt * 8 + g)) : 0u;
    }
    const int64_t grp0 = r0 >> 3, grp1 = (r1 + 7) >> 3;
    const int ntiles = (int)((grp1 - grp0 + NG - 1) / NG);
    auto issue = [&](int tile) {
        const int64_t g0 = grp0 + (int64_t)tile * NG;
        const int gn = (int)min((int64_t)NG, grp1 - g0);
        const unsigned bytes = (unsigned)(gn * R::REC * sizeof(double));
        mbar_expect_tx(&bar[tile & 1], bytes);
        bulk_g2s(s_buf + (tile & 1) * BUF, pack + g0 * R::REC, bytes, &bar[tile & 1]);
    };
    if (threadIdx.x == 0 && ntiles > 0) issue(0);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int64_t i0 = r0 + (int64_t)tile * TI;
        const double *buf = s_buf + (tile & 1) * BUF;
        if (threadIdx.x == 0 && tile + 1 < ntiles) issue(tile + 1);
        mbar_wait(&bar[tile & 1], (tile >> 1) & 1);
#pragma unroll 1
        for (int grp = 0; grp < NG; ++grp) {
            const int64_t ig = i0 + grp * 8;
            if (ig >= r1) break;
            const double *rec = buf + grp * R::REC;
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
                const double b = rec[s * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
            }
            const int sh0 = ((grp & 3) << 3) + 2 * q;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double y; bool big;
                const double p0 = fast_expit_tab(((bits[mt] >> sh0) & 1u) ? zf[mt] : c[mt] - d[mt][0], tabl, y, big);
                const double p1 = fast_expit_tab(((bits[mt] >> (sh0 + 1)) & 1u) ? zf[mt] : c[mt] - d[mt][1], tabl, y, big);
                d[mt][0] = p0; d[mt][1] = p1;      // cells beyond n have EU_new = 0 in the packed records
            }
#pragma unroll
            for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = rec[R::F1 + (sp * NT + nt) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], d[mt][sp], b);
                }
        }
        __syncthreads();   // every warp is done with this buffer before it is refilled
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

// ------------------------------------------------------------------------------------------------------------
// Pass B on the STORED prob_D.  The cell pass of the previous iteration evaluated exactly the prob_D this pass needs
// (same triple); with 180 GB of HBM it can be kept -- 4 bytes per entry, 80 GB at 1M x 20k -- and streamed back here, which
// leaves this pass with the second product only: no first product, no expit, no occupancy words.
// A tile is 64 words [gene in tile][cell in tile], i.e. lane (g, q) reads its accumulator pair d0, d1 = P[gene g][cells 2q, 2q+1]
// as one 8-byte word; the 16 gene groups of a CTA are contiguous (4 KB per cell group).
// ------------------------------------------------------------------------------------------------------------
template <int KP>
struct ZiRecF2 {
    static constexpr int NT = (KP + 7) / 8, REC = 2 * NT * 32;
};
template <int KP>
__global__ void k_zi_pack_f2(int64_t rows, const double *__restrict__ M2, double *__restrict__ out) {
    using R = ZiRecF2<KP>;
    const int64_t ngroups = (rows + 7) >> 3, total = ngroups * R::REC;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t grp = idx / R::REC;
        const int o2 = (int)(idx - grp * R::REC), lane = o2 & 31, g = lane >> 2, q = lane & 3;
        const int t = o2 >> 5, sp = t / R::NT, nt = t - sp * R::NT;
        const int64_t row = grp * 8 + 2 * q + sp;
        const int k = 8 * nt + g;
        out[idx] = (row < rows && k < KP) ? M2[row * KP + k] : 0.0;
    }
}
template <int KP, int MT, int TI>
__global__ void __launch_bounds__(128, 4) k_zi_genes_stored(ZiBArgs a) {
    const double *__restrict__ pack = a.pack;
    using R = ZiRecF2<KP>;
    constexpr int NT = R::NT, NG = TI / 8, BUF = NG * R::REC;
    static_assert(NG % 2 == 0, "the group loop is unrolled by two");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_buf = reinterpret_cast<double *>(smem_raw);       // 2 x NG records of EU_new second-product fragments
    __shared__ uint64_t bar[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    const int gene0 = blockIdx.x * (32 * MT) + w * (8 * MT);
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_chunk;        // multiple of 64
    const int64_t r1 = min(a.n, r0 + a.rows_per_chunk);
    const int64_t grp0 = r0 >> 3, grp1 = (r1 + 7) >> 3;
    const int ntiles = (int)((grp1 - grp0 + NG - 1) / NG);
    const int64_t ngg = (a.p + 7) >> 3;
    const uint32_t *__restrict__ pbase = a.pstore + ((int64_t)(gene0 >> 3) << 6) + 2 * lane;
    bool have[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) have[mt] = gene0 + mt * 8 < a.p;
    auto fetch = [&](int64_t cg, uint2 (&dst)[MT]) {       // the MT tiles of cell group cg (zeros beyond the chunk)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            dst[mt] = make_uint2(0u, 0u);
            if (have[mt] && cg < grp1) dst[mt] = __ldg(reinterpret_cast<const uint2 *>(pbase + ((cg * ngg + mt) << 6)));
        }
    };
    auto issue = [&](int tile) {
        const int64_t g0 = grp0 + (int64_t)tile * NG;
        const int gn = (int)min((int64_t)NG, grp1 - g0);
        const unsigned bytes = (unsigned)(gn * R::REC * sizeof(double));
        mbar_expect_tx(&bar[tile & 1], bytes);
        bulk_g2s(s_buf + (tile & 1) * BUF, pack + g0 * R::REC, bytes, &bar[tile & 1]);
    };
    auto second_product = [&](const uint2 (&pq)[MT], const double *rec) {
        double d[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { d[mt][0] = __uint2double_rn(pq[mt].x); d[mt][1] = __uint2double_rn(pq[mt].y); }
#pragma unroll
        for (int sp = 0; sp < 2; ++sp)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double b = rec[(sp * NT + nt) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], d[mt][sp], b);
            }
    };
    uint2 pa[MT], pb[MT];                                   // two cell groups in flight ahead of the one being multiplied
    fetch(grp0, pa);
    fetch(grp0 + 1, pb);
    if (threadIdx.x == 0 && ntiles > 0) issue(0);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int64_t cg0 = grp0 + (int64_t)tile * NG;
        const double *buf = s_buf + (tile & 1) * BUF;
        if (threadIdx.x == 0 && tile + 1 < ntiles) issue(tile + 1);
        mbar_wait(&bar[tile & 1], (tile >> 1) & 1);
#pragma unroll 1
        for (int grp = 0; grp < NG; grp += 2) {
            const int64_t cg = cg0 + grp;
            if (cg >= grp1) break;
            uint2 cur[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) cur[mt] = pa[mt];
            fetch(cg + 2, pa);
            second_product(cur, buf + grp * R::REC);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) cur[mt] = pb[mt];
            fetch(cg + 3, pb);
            if (cg + 1 < grp1) second_product(cur, buf + (grp + 1) * R::REC);
        }
        __syncthreads();   // every warp is done with this buffer before it is refilled
    }
    const double inv = 1.0 / PCMF_PSTORE_SCALE;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int j = gene0 + mt * 8 + g;
        if (j >= a.p) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = nt * 8 + 2 * q + e;
                if (k < a.K) atomicAdd(a.tmpMat + (int64_t)j * KP + k, acc[mt][nt][e] * inv);
            }
    }
}

}  // namespace pcmf
