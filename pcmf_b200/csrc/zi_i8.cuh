// Cell-side dense zero-inflation pass (pass A) of the FP64 mode with its SECOND product on the 5th-generation tensor cores --
// as EXACT integer arithmetic, so that the FP64 mode keeps its accuracy.
//
// k_zi_cells_mma (zi_mma.cuh) spends 6 of its 11 DMMA per 8 x 8 tile on PV_i += P_ij EVS_j (zi_gap_factor_model.cpp:336), and a
// DMMA holds the sub-partition's dispatch port for 16 cycles.  FP64 has no tcgen05 kind -- but that product does not need one:
//   * P is a probability and the pass already rounds it to 32-bit fixed point, Pq = rint(P (2^32 - 1)) (0 and 1 exact, absolute
//     error 1.2e-10 elsewhere), for the copy the gene-side pass streams back (PCMF_PSTORE_SCALE).  The four BYTES of that word are
//     four base-256 digits b_0..b_3, and a word written to tensor memory is read by tcgen05.mma kind::i8 as four consecutive K
//     elements: the A operand is the stored word itself,  A[cell][4 j + s] = b_s(Pq[cell][gene j]);  no repacking.
//   * EVS >= 0 is cut per factor column into ND = 6 base-256 digits e_0..e_5 (most significant first) of the 48-bit fixed-point
//     F = rint(EVS 2^(48 - ex_k)), 2^ex_k > max_j EVS_jk.  Pq F = sum_{s,t} b_s e_t 2^(8 (s - t + 5)): the digit products of equal
//     weight d = s - t share one accumulator,  B_d[k][4 j + s] = e_(s - d)[j][k],  d = 3 .. -2, and the six operands are stacked
//     along N: ONE MMA (M = 128 cells, N = 6 x 24, K = 32 bytes = 8 genes) per 8-gene group.  The dropped weights d <= -3 are
//     below 2^-46 of full scale (with five digits and weights, 2^-38: that truncation is a BIAS of every term and showed as
//     1.4e-7 on single entries of b2 where P EU is 1e4 below the column maximum).
//   * u8 x u8 -> s32 in tensor memory is exact: at most 4 digit pairs x 255^2 per gene and accumulator, so the accumulators are
//     read out (tcgen05.ld), scaled and added to PV in FP64 every 4096 genes, far below 2^31.
// The kernel is otherwise k_zi_cells_mma: first product lambda = EU_i . EVS_j as DMMA tiles (K = 20 is five m8n8k4 steps, as
// cheap as the FP64 pipe gets), exact-range FP64 expit, column sums, entropy, log-likelihood monitor, the stored prob_D.  The
// tcgen05.st shape 16x256b takes the mma.sync accumulator layout of two 8-row tiles as it is (row g / g + 8, columns 2q, 2q + 1).
// Checked bit for bit on the device by scripts/ubench_i8_ts.cu; SASS: UTCIMMA, STTM.16dp256bit, LDTM, UTCBAR, UBLKCP.
#pragma once
#include "zi_mma.cuh"
#include "zi_tc.cuh"

#ifndef PCMF_I8_UNROLL
#define PCMF_I8_UNROLL 1
#endif

namespace pcmf {

constexpr int I8_UNROLL = PCMF_I8_UNROLL;   // 8-gene groups of a tile unrolled together in the cell-side pass
constexpr int I8_TJ = 32;              // genes per shared-memory tile (one occupancy word per cell and tile)
constexpr int I8_STAGES = 3;           // tile buffers of the digit operands in shared memory (and slots of the per-tile counters / column sums)
constexpr int I8_FLUSH = 128;          // tiles between two read-outs of the integer accumulators (4096 genes)
constexpr int I8_TAIL = 128;           // doubles after the records: [colmax 64 | cfac 64]

constexpr int I8_STAGES_A = 4;         // tile buffers of the first-product records (released when every warp is through with a tile)

// Packed gene side of the pass, per group of 8 genes:  part A = [ f1: KS x 32 doubles (as ZiRec) | c: 8 | zf: 8 | e^-c: 8 ]  for the
// FP64 warps, part B = the ND digit operands B_3, B_2, .. stacked along N ([NTOT rows][32 bytes], K-major canonical; row
// di NFS + k = digit weight 3 - di, factor k), so that ONE MMA per 8-gene group fills all ND accumulators.
// The two parts are separate arrays (all A records, then all B records, then the tail [colmax 64 | cfac 64]) and travel through
// separate rings of shared-memory buffers: A is needed when a tile starts, B when it ends.
template <int KP>
struct ZiRecI8 {
    static_assert(KP <= 32, "the integer second product holds 6 x 32 accumulator columns and 2 x 32 of P in 256 TMEM columns");
    static constexpr int KS = KP / 4;
    static constexpr int NFS = (KP + 7) / 8 * 8;                          // accumulator columns per digit weight
    static constexpr int ND = 6;                                          // digits of the streamed factor matrix = digit weights kept: d = 3 .. 4 - ND
    static constexpr int NTOT = (ND * NFS + 15) / 16 * 16;                // N of the one MMA per 8-gene group: the ND operands stacked
    static constexpr int F1 = KS * 32, CZ = 24, RA = F1 + CZ;             // doubles per group, part A
    static constexpr int B8 = NTOT * 32;                                  // bytes per group, part B: [NTOT rows][32 bytes], K-major canonical
    static constexpr int COL_ACC = 0, COL_A = NTOT, TMEM_COLS = 256;
    static constexpr int PST = (TMEM_COLS - NTOT) / I8_TJ >= 3 ? 3 : 2;   // P buffers in tensor memory (tiles in flight towards the MMAs)
    static_assert(COL_A + PST * I8_TJ <= TMEM_COLS, "tensor-memory budget");
    static constexpr size_t SMEM = (size_t)(I8_TJ / 8) * (I8_STAGES_A * RA * 8 + I8_STAGES * B8) + sizeof(double) * (I8_STAGES * 4 * I8_TJ + kTabDoubles);
    // (the cell-side packing of EU_new for k_zi_genes_i8 has no part A)
    static size_t pack_doubles(int64_t groups, bool with_a = true) { return (size_t)groups * ((with_a ? RA : 0) + B8 / 8) + I8_TAIL; }
    __host__ __device__ static int64_t off_b(int64_t groups, bool with_a = true) { return with_a ? groups * RA : 0; }                 // in doubles
    __host__ __device__ static int64_t off_tail(int64_t groups, bool with_a = true) { return groups * ((with_a ? RA : 0) + B8 / 8); }   // in doubles
};

// per-column maximum of a non-negative row-major [rows][KP] matrix (bit patterns of non-negative doubles order as integers)
template <int KP>
__global__ void k_i8_colmax(int64_t rows, const double *__restrict__ M, unsigned long long *__restrict__ colmax) {
    __shared__ unsigned long long sm[KP];
    if (threadIdx.x < KP) sm[threadIdx.x] = 0ull;
    __syncthreads();
    const int64_t total = rows * KP;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const double v = M[idx];
        if (v > 0.0) atomicMax(&sm[(int)(idx % KP)], (unsigned long long)__double_as_longlong(v));
    }
    __syncthreads();
    if (threadIdx.x < KP && sm[threadIdx.x]) atomicMax(&colmax[threadIdx.x], sm[threadIdx.x]);
}

// canonical K-major, no-swizzle operand layout for 8-bit elements: core matrices of 8 rows x 16 bytes, 128 B between 8-row
// groups (SBO), (N / 8) x 128 B between 16-byte K chunks (LBO)
__host__ __device__ inline int i8_canon(int n, int kb, int N) { return (kb >> 4) * (N / 8) * 128 + (n >> 3) * 128 + (n & 7) * 16 + (kb & 15); }

// Writes both parts of the packed gene side (ZiRecI8) and, in the tail, cfac:  PV_ik = cfac_k sum_d 2^(8 (d + 1)) acc_d[i][k].
template <int KP, bool WITH_A>
__global__ void k_zi_pack_i8(int64_t rows, const double *__restrict__ EVS, const double *__restrict__ cz, const int32_t *__restrict__ flag,
                             double *__restrict__ out) {
    using R = ZiRecI8<KP>;
    const int64_t ngroups = (rows + 7) >> 3;
    const unsigned long long *colmax = reinterpret_cast<const unsigned long long *>(out + R::off_tail(ngroups, WITH_A));
    __shared__ double s_mult[KP];
    if (threadIdx.x < KP) {
        const double m = __longlong_as_double((long long)colmax[threadIdx.x]);
        int ex = 0;
        double mult = 0.0, cf = 0.0;
        if (m > 0.0 && m < 1e300) {
            frexp(m, &ex);                                                // m = f 2^ex, f in [0.5, 1)  ->  m < 2^ex
            mult = ldexp(1.0, 8 * R::ND - ex);
            cf = ldexp(1.0 / PCMF_PSTORE_SCALE, ex - 8 * R::ND + 24);
        }
        s_mult[threadIdx.x] = mult;
        if (blockIdx.x == 0) out[R::off_tail(ngroups, WITH_A) + 64 + threadIdx.x] = cf;
    }
    __syncthreads();
    // doubles of the record head, one per thread
    const int64_t head = WITH_A ? ngroups * R::RA : 0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < head; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t grp = idx / R::RA;
        const int o = (int)(idx - grp * R::RA);
        double v = 0.0;
        if (o < R::F1) {
            const int s = o >> 5, lane = o & 31, g = lane >> 2, q = lane & 3;
            const int64_t row = grp * 8 + g;
            if (row < rows) v = EVS[row * KP + 4 * s + q];
        } else {
            const int o3 = o - R::F1, e = o3 & 7;
            const int64_t row = grp * 8 + e;
            const int f = row < rows ? flag[row] : 2;
            if (o3 < 8) v = (row < rows && f == 0) ? cz[row] : ((f == 1) ? 1e300 : -1e300);
            else if (o3 < 16) v = (f == 2) ? -1e300 : 1e300;
            else v = (row < rows && f == 0) ? exp(-cz[row]) : 1.0;
        }
        out[idx] = v;
    }
    // digit operands: one thread per (group, operand row, gene in group) writes the row's 4 bytes (s = 0..3) for that gene.
    // Row r = di NFS + k of the stacked operand: digit weight d = 3 - di, factor k; rows beyond ND NFS are zero.
    const int64_t items = ngroups * R::NTOT * 8;
    unsigned char *b8all = reinterpret_cast<unsigned char *>(out + R::off_b(ngroups, WITH_A));
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t grp = idx / (R::NTOT * 8);
        const int rem = (int)(idx - grp * (R::NTOT * 8)), r = rem >> 3, jl = rem & 7;
        const int di = r / R::NFS, k = r - di * R::NFS, d = 3 - di;
        const int64_t row = grp * 8 + jl;
        uint32_t w = 0u;
        if (di < R::ND && row < rows && k < KP) {
            const double x = rint(EVS[row * KP + k] * s_mult[k]);
            constexpr unsigned long long FMAX = (1ull << (8 * R::ND)) - 1ull;
            const unsigned long long F = x <= 0.0 ? 0ull : (x >= (double)FMAX ? FMAX : (unsigned long long)x);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int t = s - d;                                      // digit e_t, t = 0 most significant: bits [8 (4 - t), +8) of F
                if (t >= 0 && t < R::ND) w |= (uint32_t)((F >> (8 * (R::ND - 1 - t))) & 255ull) << (8 * s);
            }
        }
        *reinterpret_cast<uint32_t *>(b8all + grp * R::B8 + i8_canon(r, 4 * jl, R::NTOT)) = w;
    }
}

__device__ __forceinline__ uint64_t i8_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor: D = S32 (2 at bit 4), A = B = U8 (0 at bits 7 / 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t i8_idesc(int m, int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void i8_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// CTA = 128 cells (the 128 lanes of tensor memory) = NW warps: NW = 4 (default), warp w owns the 32 cells of row block w as four
// 8-row tiles and all four 8-gene groups of every 32-gene tile; NW = 8 (PCMF_I8_NW), two warps per row block share its groups.
// No CTA-wide barrier in the main loop:
//   * a warp waits for the tile's first-product records (bar_fullA) and for the MMAs that last read the P buffer it is about to
//     overwrite (bar_mma of PST tiles ago), computes, stores its P words to tensor memory and its column sums to shared memory,
//     and counts itself in (s_cnt);
//   * the warp that counts in LAST owns the tile's tail (tile_tail): adds the column sums, starts the bulk copies of later tiles
//     (part A three tiles ahead, part B two tiles ahead once the MMAs of the tile before have released its buffer), issues the
//     tile's MMAs -- one per 8-gene group -- and their commit.
// Two CTAs per SM (tensor memory: 2 x 256 columns; shared memory: 2 x ~96 KB).
// Select-free expit for the pass without the log-likelihood monitor.  With za = min(|z|, 40), t = e^-za in (4e-18, 1], r = 1 / (1 + t):
//      P = z >= 0 ? r : t r            1 exactly for z >= 37 and below 2^-57 for z <= -40, where internal::expit (internal.h:151-155)
//                                      clamps to exactly 1 / 0 from |z| = 30 on: the two differ by at most e^-30 = 9.4e-14
//      -[P log P + (1 - P) log(1 - P)] = log(1 + t) + za t r     no cancellation, below 40 e^-40 = 1.7e-16 for the clamped entries
// so the entries X != 0 (z = +-1e300 from the packed record), the whole-gene shortcuts and the rows beyond n need no select of
// their own, and the running product of (1 + t) <= 2 is read out once per tile.  The clamp is two integer instructions on the
// high word (the low word rides along: |z| in [40, 40 + 2^-15)).
// (fast_exp_neg_tab with the table row addressed as one multiply-add on the 32-bit shared-memory address of the lane's copy)
__device__ __forceinline__ double fast_exp_neg_tab_a(double z, uint32_t tab_addr) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    double t = fma(z, -92.33248261689366, MAGIC);
    const int m = __double2loint(t);
    t -= MAGIC;
    const double r = fma(t, -0.010830424696249145, -z);
    double p = 4.16666666666666666667e-02;
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double tb;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tb) : "r"((uint32_t)(m & 63) * (uint32_t)(kTabRep * 8) + tab_addr));
    const double sc = __hiloint2double(__double2hiint(tb) + (int)((unsigned)m << 14), __double2loint(tb));
    return p * sc;
}
__device__ __forceinline__ double expit_sym(double z, uint32_t tab_addr, double &y, double &ent) {
    const int hi = __double2hiint(z);
    const int ah = min(hi & 0x7fffffff, 0x40440000);
    const double za = __hiloint2double(ah, __double2loint(z));
    const double t = fast_exp_neg_tab_a(za, tab_addr);
    y = 1.0 + t;
    const double r = fast_rcp2(y);
    const double tr = t * r;
    ent = fma(-za, tr, ent);          // the caller's running -sum za t r
    return hi >= 0 ? r : tr;
}

__device__ __forceinline__ void i8_mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// (A fifth warp that owns every tile's tail, so that no compute warp falls behind by the MMA issue, was built and measured: 75.2 ms
// against 73.0 ms at C4 -- the waits on the P buffers are the skew of equal warps, not the tail.)
template <int KP, int NW, bool STORE_P, bool WITH_LL>
#ifndef PCMF_I8_MINB
#define PCMF_I8_MINB 2
#endif
__global__ void __launch_bounds__(32 * NW, PCMF_I8_MINB) k_zi_cells_i8(ZiAArgs a) {
    static_assert(NW == 4 || NW == 8, "one or two warps per 32-cell row block");
    using R = ZiRecI8<KP>;
    constexpr int MT = 4, KS = R::KS, NFS = R::NFS, NTOT = R::NTOT, TJ = I8_TJ, NG = TJ / 8, BUFA = NG * R::RA, BUFB = NG * R::B8, GPW = NG / (NW / 4);
    const double *__restrict__ pack = a.pack;
#ifdef PCMF_I8_ABLATE
    const int abl = a.want_ll >> 8;
#else
    constexpr int abl = 0;
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_bufB = smem_raw;                           // I8_STAGES x NG digit-operand records (part B)
    double *s_bufA = reinterpret_cast<double *>(smem_raw + I8_STAGES * BUFB);   // I8_STAGES_A x NG first-product records (part A)
    double *s_cs = s_bufA + I8_STAGES_A * BUFA;                 // I8_STAGES x 4 x TJ: per-row-block column sums of a tile
    double *s_tab = s_cs + I8_STAGES * 4 * TJ;                  // kTabDoubles
    __shared__ double sh[32];
    __shared__ double s_cf[KP];
    __shared__ __align__(8) uint64_t bar_fullA[I8_STAGES_A], bar_fullB[I8_STAGES], bar_mma[R::PST];
    __shared__ uint32_t s_bits[NW][I8_STAGES][32];              // per warp: occupancy words of its 32 cells for a tile (cp.async)
    __shared__ int s_cnt[I8_STAGES];
    __shared__ uint32_t tmem_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3, rb = w & 3, gh = w >> 2;
    load_pow2_tab(s_tab);
    const double *tabl = pow2_tab_lane(s_tab);
    const uint32_t tab_addr = smem_u32(tabl);
    const int ngroups = (a.p + 7) >> 3;
    if (threadIdx.x < KP) s_cf[threadIdx.x] = pack[R::off_tail(ngroups) + 64 + threadIdx.x];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&bar_fullB[s], 1); s_cnt[s] = 0; }
#pragma unroll
        for (int s = 0; s < R::PST; ++s) mbar_init(&bar_mma[s], 1);
#pragma unroll
        for (int s = 0; s < I8_STAGES_A; ++s) mbar_init(&bar_fullA[s], 1);
        mbar_fence_init();
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(R::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tmem_base;
    const uint32_t idesc = i8_idesc(128, NTOT);
    const int64_t cell0 = (int64_t)blockIdx.x * 128 + rb * 32;
    double a1[MT][KS];
    bool live[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int64_t i = cell0 + mt * 8 + g;
        live[mt] = i < a.n;
#pragma unroll
        for (int s = 0; s < KS; ++s) a1[mt][s] = live[mt] ? __ldg(a.EU + i * KP + 4 * s + q) : 0.0;
    }
    // rows beyond n: lambda starts at 1e305, above every effective logit, so the expit itself clamps them to P = 0 (word 0)
    const unsigned st_ngg = (unsigned)((a.p + 7) >> 3);
    const unsigned st_tile0 = (unsigned)(cell0 >> 3) * st_ngg;
    // stored prob_D: tile (cell group, gene group) as two core matrices [cell >> 2][gene][cell & 3] -- the A operand of k_zi_genes_i8
    // (lanes g and g ^ 1 swap one word each, so that a lane holds one gene for two adjacent cells: one 8-byte store per row tile,
    //  a warp's store = the whole 256-byte tile)
    uint32_t *const st_lane = STORE_P ? a.pstore + (g >> 2) * 32 + (2 * q + (g & 1)) * 4 + (g & 2) : nullptr;
    double ent_lin = 0, ent_lin2 = 0, ent_log = 0, prod = 1.0;
    double ll_lin = 0, ll_log = 0, ll_prod = 1.0;

    const int ntiles_all = (ngroups + NG - 1) / NG;
    const int tiles_per_chunk = (ntiles_all + (int)gridDim.y - 1) / (int)gridDim.y;
    const int tile_first = (int)blockIdx.y * tiles_per_chunk;
    const int ntiles = max(0, min(ntiles_all, tile_first + tiles_per_chunk) - tile_first);
    const int rot = ntiles > 0 ? (int)(blockIdx.x % (unsigned)ntiles) : 0;
    auto tile_id = [&](int t) { const int u = t + rot; return tile_first + (u >= ntiles ? u - ntiles : u); };
    // one thread: arm the barrier with the byte count, then one bulk copy (part A / part B of tile t)
    auto issueA = [&](int t) {
        const int g0 = tile_id(t) * NG, gn = min(NG, ngroups - g0);
        const unsigned bytes = (unsigned)(gn * R::RA * 8);
        const int sa = t % I8_STAGES_A;
        mbar_expect_tx(&bar_fullA[sa], bytes);
        bulk_g2s(s_bufA + sa * BUFA, pack + (int64_t)g0 * R::RA, bytes, &bar_fullA[sa]);
    };
    const unsigned char *packB = reinterpret_cast<const unsigned char *>(pack + R::off_b(ngroups));
    auto issueB = [&](int t) {
        const int g0 = tile_id(t) * NG, gn = min(NG, ngroups - g0);
        const unsigned bytes = (unsigned)(gn * R::B8);
        const int sb = t % I8_STAGES;
        mbar_expect_tx(&bar_fullB[sb], bytes);
        bulk_g2s(s_bufB + sb * BUFB, packB + (int64_t)g0 * R::B8, bytes, &bar_fullB[sb]);
    };
    // occupancy words of this warp's 32 cells for tile t, one per lane, straight into shared memory (zeros beyond n)
    auto fetch_bits = [&](int t) {
        const int64_t i = cell0 + lane;
        const bool in = i < a.n;
        const uint32_t *src = a.bitT + (int64_t)tile_id(t) * a.n + (in ? i : 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(&s_bits[w][t % I8_STAGES][lane])), "l"(src), "r"(in ? 4 : 0) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // read-out of the integer accumulators: warp (rb, gh) takes lanes 32 rb.. and the column half gh; thread = one cell
    const int64_t my_cell = cell0 + lane;
    const uint32_t lane_addr = t0 + ((uint32_t)(rb * 32) << 16);
    double sum_pl = 0.0;
    bool pv_started = false;
    auto flush = [&]() {
        // 8-column chunks of the NFS columns per digit weight, dealt to the one or two warps of the row block
#pragma unroll
        for (int cb = 8 * gh; cb < NFS; cb += 8 * (NW / 4)) {
            uint32_t v[R::ND][8];
#pragma unroll
            for (int di = 0; di < R::ND; ++di)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[di][0]), "=r"(v[di][1]), "=r"(v[di][2]), "=r"(v[di][3]), "=r"(v[di][4]), "=r"(v[di][5]), "=r"(v[di][6]), "=r"(v[di][7])
                             : "r"(lane_addr + (uint32_t)(R::COL_ACC + di * NFS + cb)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (my_cell < a.n) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int k = cb + c;
                    if (k < KP) {
                        double x = (double)(int32_t)v[0][c];
#pragma unroll
                        for (int di = 1; di < R::ND; ++di) x = fma(x, 256.0, (double)(int32_t)v[di][c]);
                        x *= s_cf[k];
                        // sum_ij prob_D_ij UVt_ij = sum_ik EU_ik PV_ik (zi...cpp:182) is linear in the parts added here
                        sum_pl = fma(__ldg(a.EU + my_cell * KP + k), x, sum_pl);
                        double *dst = a.PV + my_cell * KP + k;
                        if (gridDim.y == 1) {
                            if (pv_started) x += *dst;
                            *dst = x;
                        } else atomicAdd(dst, x);
                    }
                }
            }
        }
        pv_started = true;
    };
    // the tail of tile `tile`: column sums over the CTA's 128 cells, the bulk copies of later tiles, the tile's MMAs and their commit
    auto tile_tail = [&](int tile) {
        const int st = tile % I8_STAGES, sp = tile % R::PST;
        const unsigned ph = (unsigned)((tile / I8_STAGES) & 1);
        const int j0 = tile_id(tile) * TJ;
        const int ng_valid = min(NG, ngroups - tile_id(tile) * NG);
        const double *cs_buf = s_cs + st * 4 * TJ;
        {   // one atomic per gene
            const int t = lane;
            if (t < ng_valid * 8 && j0 + t < a.p)
                atomicAdd(a.colsumP + j0 + t, (cs_buf[t] + cs_buf[TJ + t]) + (cs_buf[2 * TJ + t] + cs_buf[3 * TJ + t]));
        }
        if (lane == 0) {
            // part A of the tile three ahead goes where the tile before this one was: every warp is through with that
            if (tile + I8_STAGES_A - 1 < ntiles) issueA(tile + I8_STAGES_A - 1);
            // part B of the tile after next goes where the MMAs of the tile before this one read theirs
            if (tile + I8_STAGES - 1 < ntiles) {
                if (tile >= 1) mbar_wait(&bar_mma[(tile - 1) % R::PST], (unsigned)((((tile - 1) / R::PST) & 1)));
                issueB(tile + I8_STAGES - 1);
            }
            mbar_wait(&bar_fullB[st], ph);
            const bool fresh = (tile % I8_FLUSH) == 0;           // the accumulators were read out (or never written): overwrite
            const uint32_t bbase = smem_u32(s_bufB + st * BUFB);
            for (int grp = 0; grp < ng_valid && !(abl & 64); ++grp)     // one MMA per 8-gene group: M = 128 cells, N = NTOT, K = 32 bytes
                i8_mma_ts(t0 + R::COL_ACC, t0 + (uint32_t)(R::COL_A + sp * TJ + grp * 8),
                          i8_desc(bbase + (uint32_t)(grp * R::B8), (NTOT / 8) * 128, 128), idesc, !(fresh && grp == 0));
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma[sp])) : "memory");
        }
        __syncwarp();
    };
    if (threadIdx.x == 0) {
        for (int t = 0; t < min(ntiles, I8_STAGES_A - 1); ++t) issueA(t);
        for (int t = 0; t < min(ntiles, I8_STAGES - 1); ++t) issueB(t);
    }
    if (ntiles > 0) fetch_bits(0);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int st = tile % I8_STAGES;
        const unsigned ph = (unsigned)((tile / I8_STAGES) & 1);
        const int j0 = tile_id(tile) * TJ;
        const int sa = tile % I8_STAGES_A;
        const double *buf = s_bufA + sa * BUFA;
        double *cs_buf = s_cs + st * 4 * TJ;
        if (tile + 1 < ntiles) fetch_bits(tile + 1);
        // the MMAs of PST tiles ago are through with the P buffer this tile overwrites
        const int sp = tile % R::PST;
        const unsigned php = (unsigned)((tile / R::PST) & 1);
        if (tile >= R::PST) mbar_wait(&bar_mma[sp], php ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mbar_wait(&bar_fullA[sa], (unsigned)((tile / I8_STAGES_A) & 1));
        if (tile + 1 < ntiles) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        uint32_t bits[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) bits[mt] = s_bits[w][st][mt * 8 + g];
        const int ng_valid = min(NG, ngroups - tile_id(tile) * NG);
#pragma unroll (I8_UNROLL)
        for (int gi = 0; gi < GPW; ++gi) {
            const int grp = GPW * gh + gi;
            if (grp >= ng_valid) break;
            const int jg = j0 + grp * 8;
            const double *rec = buf + grp * R::RA;
            double d[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { d[mt][0] = live[mt] ? 0.0 : 1e305; d[mt][1] = d[mt][0]; }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = rec[s * 32 + lane];
                if (!(abl & 16)) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(d[mt][0], d[mt][1], a1[mt][s], b);
                } else {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) d[mt][0] += b;
                }
            }
            const double2 c2 = *reinterpret_cast<const double2 *>(rec + R::F1 + 2 * q);
            const double2 n2 = *reinterpret_cast<const double2 *>(rec + R::F1 + 8 + 2 * q);
            double2 ec2 = make_double2(1.0, 1.0);
            if (WITH_LL) ec2 = *reinterpret_cast<const double2 *>(rec + R::F1 + 16 + 2 * q);
            const int sh0 = (grp << 3) + 2 * q;
            const unsigned st_tile = st_tile0 + (unsigned)(jg >> 3);
            double cs0 = 0.0, cs1 = 0.0;
            uint32_t pw[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const bool nz0 = (bits[mt] >> sh0) & 1u, nz1 = (bits[mt] >> (sh0 + 1)) & 1u;
                const double z0 = nz0 ? n2.x : c2.x - d[mt][0], z1 = nz1 ? n2.y : c2.y - d[mt][1];
                double p0, p1;
                if (WITH_LL) {
                    double y0, y1, em0, em1; bool big0, big1;
                    p0 = fast_expit_tab(z0, tabl, y0, big0, em0); p1 = fast_expit_tab(z1, tabl, y1, big1, em1);
                    ll_prod *= (big0 ? 1.0 : fma(em0, em0, ec2.x)) * (big1 ? 1.0 : fma(em1, em1, ec2.y));
                    ll_lin += (big0 ? 0.0 : z0) + (big1 ? 0.0 : z1);
                    ll_lin -= ((big0 && !nz0 && z0 > 0.0) ? d[mt][0] : 0.0) + ((big1 && !nz1 && z1 > 0.0) ? d[mt][1] : 0.0);
                    // -bernoulli_loglike(prob_D, prob_D) over 0 < P < 1: sum (1-P)(-z) - log(1 + e^-z)
                    const bool g0 = !big0, g1 = !big1;
                    ent_lin = fma(g0 ? (p0 - 1.0) : 0.0, z0, ent_lin);
                    ent_lin = fma(g1 ? (p1 - 1.0) : 0.0, z1, ent_lin);
                    prod *= (g0 ? y0 : 1.0) * (g1 ? y1 : 1.0);
                } else {
                    double y0, y1;
                    p0 = expit_sym(z0, tab_addr, y0, ent_lin); p1 = expit_sym(z1, tab_addr, y1, ent_lin2);
                    prod *= y0 * y1;
                }
                cs0 += p0; cs1 += p1;
                // rint(P (2^32 - 1)) read off the low mantissa word of 2^52 + P (2^32 - 1)
                pw[mt][0] = (uint32_t)__double2loint(fma(p0, PCMF_PSTORE_SCALE, 4503599627370496.0));
                pw[mt][1] = (uint32_t)__double2loint(fma(p1, PCMF_PSTORE_SCALE, 4503599627370496.0));
                if (STORE_P && !(abl & 2)) {
                    uint32_t *tp = st_lane + ((size_t)(st_tile + (unsigned)mt * st_ngg) << 6);
                    const bool odd = g & 1;
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, odd ? pw[mt][0] : pw[mt][1], 4);
                    *reinterpret_cast<uint2 *>(tp) = odd ? make_uint2(got, pw[mt][1]) : make_uint2(pw[mt][0], got);
                }
            }
            // the words of two row tiles at a time into this tile's P buffer: lanes 32 rb + 16 h + {g, g + 8}, columns 8 grp + {2q, 2q + 1}
            if (!(abl & 1))
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t addr = t0 + ((uint32_t)(rb * 32 + h * 16) << 16) + (uint32_t)(R::COL_A + sp * TJ + grp * 8);
                asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(addr), "r"(pw[2 * h][0]), "r"(pw[2 * h][1]), "r"(pw[2 * h + 1][0]), "r"(pw[2 * h + 1][1]) : "memory");
            }
            if (WITH_LL && prod > 1e100) { ent_log += log(prod); prod = 1.0; }
            if (WITH_LL && (ll_prod > 1e60 || ll_prod < 1e-60)) { ll_log += log(ll_prod); ll_prod = 1.0; }
            if (!(abl & 4)) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
                    cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
                }
                if (g == 0) *reinterpret_cast<double2 *>(cs_buf + rb * TJ + grp * 8 + 2 * q) = make_double2(cs0, cs1);
            }
        }
        if (!WITH_LL && prod > 1e200) { ent_log += log(prod); prod = 1.0; }     // (1 + t)^(16 GPW) <= 2^64 per tile
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        __syncwarp();
        // count this warp in; the last one of the CTA owns the tail of the tile
        int last = 0;
        if (lane == 0) last = atomicAdd(&s_cnt[st], 1) == NW - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) { atomicExch(&s_cnt[st], 0); asm volatile("fence.acq_rel.cta;" ::: "memory"); }
            tile_tail(tile);
        }
        if (tile + 1 == ntiles || (tile + 1) % I8_FLUSH == 0) {
            // read the accumulators out before they can overflow (and at the end): every MMA up to this tile must have landed;
            // the MMAs of the next tile are issued only after every warp has counted itself in there, i.e. after its read-out
            mbar_wait(&bar_mma[sp], php);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            flush();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    if (ntiles == 0 && gridDim.y == 1 && gh == 0 && my_cell < a.n)
        for (int k = 0; k < KP; ++k) a.PV[my_cell * KP + k] = 0.0;
    const double sum_log_y = ent_log + log(prod);
    double ent = (ent_lin + ent_lin2) - sum_log_y;
    sum_pl = block_sum(sum_pl, sh); ent = block_sum(ent, sh);
    if (threadIdx.x == 0) { atomicAdd(a.zsc + 0, sum_pl); atomicAdd(a.zsc + 1, ent); }
    if (WITH_LL) {
        double ll = ll_lin + ll_log + log(ll_prod) - sum_log_y;
        ll = block_sum(ll, sh);
        if (threadIdx.x == 0) atomicAdd(a.zsc + 2, ll);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "n"(R::TMEM_COLS));
}

// ------------------------------------------------------------------------------------------------------------------------------
// Gene-side pass on the STORED prob_D (pass B, zi_gap_factor_model.cpp:341: tmpMat = prob_D^T EU_new) as the same exact integer
// product -- and with nothing but the tensor core and the copy engine on the data path:
//   * the cell-side pass keeps its P words in 8 x 8 tiles laid out as the MMA's shared-memory operand wants them: two core
//     matrices of [8 genes][4 cells x 4 digit bytes] (word (gene, cell) at (cell >> 2) 32 + gene 4 + (cell & 3)), tile index =
//     cell group x gene groups + gene group.  The 48 gene groups of a CTA are 12 KB contiguous per cell group: ONE bulk copy
//     brings them from HBM into the A operand of three M = 128 MMAs (K = 32 bytes = 8 cells; SBO = 256 B, LBO = 128 B);
//   * EU_new is cut into digits per cell group exactly like EVS per gene group (k_zi_pack_i8 on the n x KP matrix): 4 KB per
//     8 cells, shared by the three MMAs and, through L2, by the gene blocks of the same cell chunk (adjacent in the grid);
//   * a CTA owns 384 genes x <= 8192 cells, so its accumulators (3 x NTOT columns of tensor memory) cannot overflow and are read
//     out once, scaled and added to tmpMat in FP64.
// One producer thread (bulk copies), one MMA thread, 3 stages of 32 cells; the pass is bound by the HBM stream of P (4 B / entry).
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int I8B_CG = 4;              // cell groups (of 8) per stage
constexpr int I8B_STAGES = 3;
constexpr int I8B_MAXCELLS = 8192;     // cells per CTA: 4 digit pairs x 255^2 x 8192 < 2^31

template <int KP>
struct ZiGenesI8 {
    using R = ZiRecI8<KP>;
    static constexpr int GB = 512 / R::NTOT >= 3 ? 3 : 2;                  // 128-gene sub-blocks per CTA (accumulators: GB x NTOT <= 512 columns)
    static constexpr int A_CG = GB * 16 * 256;                             // bytes of P per cell group: 16 GB gene groups x 256 B
    static constexpr int B_CG = R::B8;                                     // bytes of EU digits per cell group
    static constexpr int STAGE = I8B_CG * (A_CG + B_CG);
    static constexpr size_t SMEM = (size_t)I8B_STAGES * STAGE;
    static constexpr int TMEM_COLS = 512;
    static_assert(GB * R::NTOT <= TMEM_COLS, "tensor-memory budget");
};

template <int KP>
__global__ void __launch_bounds__(128, 1) k_zi_genes_i8(ZiBArgs a) {
    using R = ZiRecI8<KP>;
    using G = ZiGenesI8<KP>;
    constexpr int NFS = R::NFS, NTOT = R::NTOT, I8B_GB = G::GB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[I8B_STAGES], bar_empty[I8B_STAGES], bar_done;
    __shared__ double s_cf[KP];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, w = tid >> 5;
    const int64_t ngroups_c = (a.n + 7) >> 3;                           // cell groups: records of the packed EU_new
    const int ngg = (a.p + 7) >> 3;
    const int gg0 = (int)blockIdx.x * (I8B_GB * 16), ngg_blk = min(I8B_GB * 16, ngg - gg0);
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_chunk, r1 = min(a.n, r0 + a.rows_per_chunk);     // rows_per_chunk: multiple of 32
    const int64_t cg0 = r0 >> 3, cg1 = (r1 + 7) >> 3;
    const int nst = (int)((cg1 - cg0 + I8B_CG - 1) / I8B_CG);
    if (tid < KP) s_cf[tid] = a.pack[R::off_tail(ngroups_c, false) + 64 + tid];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < I8B_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tmem_base;
    const unsigned char *packB = reinterpret_cast<const unsigned char *>(a.pack + R::off_b(ngroups_c, false));
    if (tid == 0) {
        // producer: per stage, the P tiles of up to I8B_CG cell groups (one bulk copy each) and their EU digit records (one copy)
        for (int s = 0; s < nst; ++s) {
            const int st = s % I8B_STAGES;
            if (s >= I8B_STAGES) mbar_wait(&bar_empty[st], (unsigned)(((s / I8B_STAGES) - 1) & 1));
            const int64_t c0 = cg0 + (int64_t)s * I8B_CG;
            const int nc = (int)min((int64_t)I8B_CG, cg1 - c0);
            unsigned char *sA = smem_raw + (size_t)st * G::STAGE, *sB = sA + I8B_CG * G::A_CG;
            mbar_expect_tx(&bar_full[st], (unsigned)(nc * (ngg_blk * 256 + G::B_CG)));
            for (int c = 0; c < nc; ++c)
                bulk_g2s(sA + c * G::A_CG, a.pstore + (((c0 + c) * ngg + gg0) << 6), (unsigned)(ngg_blk * 256), &bar_full[st]);
            bulk_g2s(sB, packB + c0 * G::B_CG, (unsigned)(nc * G::B_CG), &bar_full[st]);
        }
    } else if (tid == 32) {
        // MMA thread: per cell group one MMA per 128-gene sub-block (M = 128 genes, N = NTOT, K = 32 bytes = 8 cells)
        const uint32_t idesc = i8_idesc(128, NTOT);
        for (int s = 0; s < nst; ++s) {
            const int st = s % I8B_STAGES;
            mbar_wait(&bar_full[st], (unsigned)((s / I8B_STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t c0 = cg0 + (int64_t)s * I8B_CG;
            const int nc = (int)min((int64_t)I8B_CG, cg1 - c0);
            const uint32_t sA = smem_u32(smem_raw + (size_t)st * G::STAGE), sB = sA + I8B_CG * G::A_CG;
            for (int c = 0; c < nc; ++c)
#pragma unroll
                for (int b = 0; b < I8B_GB; ++b)
                    if (b * 16 < ngg_blk)
                        i8_mma_ss(t0 + b * NTOT, i8_desc(sA + c * G::A_CG + b * 4096, 128, 256), i8_desc(sB + c * G::B_CG, (NTOT / 8) * 128, 128),
                                  idesc, (s > 0) || (c > 0));
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_empty[st])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
    }
    __syncwarp();
    // epilogue: thread = TMEM lane = gene of each sub-block; digit weights recombined in FP64 (exact), scaled, added to tmpMat
    if (nst > 0) {
        mbar_wait(&bar_done, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t lane_addr = t0 + ((uint32_t)(w * 32) << 16);
#pragma unroll
        for (int b = 0; b < I8B_GB; ++b) {
            if (b * 16 >= ngg_blk) break;                              // (uniform over the CTA)
            const int j = gg0 * 8 + b * 128 + tid;
#pragma unroll
            for (int cb = 0; cb < NFS; cb += 8) {
                uint32_t v[R::ND][8];
#pragma unroll
                for (int di = 0; di < R::ND; ++di)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[di][0]), "=r"(v[di][1]), "=r"(v[di][2]), "=r"(v[di][3]), "=r"(v[di][4]), "=r"(v[di][5]), "=r"(v[di][6]), "=r"(v[di][7])
                                 : "r"(lane_addr + (uint32_t)(b * NTOT + di * NFS + cb)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (j < a.p) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int k = cb + c;
                        if (k < a.K) {
                            double x = (double)(int32_t)v[0][c];
#pragma unroll
                            for (int di = 1; di < R::ND; ++di) x = fma(x, 256.0, (double)(int32_t)v[di][c]);
                            atomicAdd(a.tmpMat + (int64_t)j * KP + k, x * s_cf[k]);
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "n"(G::TMEM_COLS));
}

}  // namespace pcmf
