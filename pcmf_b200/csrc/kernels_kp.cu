// One instantiation set of the K-templated kernels; compiled once per padded width with -DPCMF_KP=<n>.
#include <algorithm>
#include <cstdlib>
#include "launch.cuh"
#include "zi_mma.cuh"
#if PCMF_KP <= 32
#include "zi_i8.cuh"
#endif
#if PCMF_KP >= 16 && PCMF_KP <= 20
#include "xtile_kernels.cuh"
#endif

#ifndef PCMF_KP
#error "compile with -DPCMF_KP=<padded K>"
#endif

namespace pcmf {
namespace {

constexpr int KP = PCMF_KP;
constexpr int ZI_TJ = 128;                       // genes per shared-memory tile in the cell-threaded ZI pass
constexpr int ZI_TI = (KP <= 32) ? 64 : 32;      // cells per shared-memory tile in the gene-threaded ZI pass

void launch_estep_rows(bool f32, const RowArgs &a, int grid, int block, cudaStream_t st) {
    if (a.qcsc && a.evw != a.ev) {
        // stored-q variant: twice the resident warps, so twice the grid
        if (f32) k_estep_rows_q<KP, float><<<grid * 2, block, 0, st>>>(a);
        else k_estep_rows_q<KP, double><<<grid * 2, block, 0, st>>>(a);
        return;
    }
    if (f32) k_estep_rows<KP, float><<<grid, block, 0, st>>>(a);
    else k_estep_rows<KP, double><<<grid, block, 0, st>>>(a);
}

template <typename VT>
void launch_gene_t(int mode, const ColArgs &a, int gx, cudaStream_t st) {
    const dim3 grid(gx, a.nblk > 1 ? a.nblk : 1);
    switch (mode) {
        case GENE_B1: k_gene_pass<KP, VT, false, false><<<grid, 128, 0, st>>>(a); break;
        case GENE_B1_LOGT: k_gene_pass<KP, VT, false, true><<<grid, 128, 0, st>>>(a); break;
        default: k_gene_pass<KP, VT, true, false><<<grid, 128, 0, st>>>(a); break;
    }
}
void launch_gene_pass(bool f32, int mode, const ColArgs &a, int grid, cudaStream_t st) {
    if (f32) launch_gene_t<float>(mode, a, grid, st);
    else launch_gene_t<double>(mode, a, grid, st);
}

constexpr int MMA_MT = (KP <= 20) ? 4 : 2;          // 8-row tiles per warp in the DMMA ZI passes
constexpr int MMA_TJ = 64;                           // genes per shared-memory tile, pass A
constexpr int MMA_TI = 64;                          // cells per shared-memory tile, pass B

inline int current_device() { int d = 0; cudaGetDevice(&d); return d & 63; }
inline int pack_grid(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 148 * 16)); }

template <int MT>
int launch_zi_cells_mma_t(const ZiAArgs &a, cudaStream_t st) {
    using R = ZiRec<KP, true>;
    const size_t smem = sizeof(double) * (2 * (MMA_TJ / 8) * R::REC + 8 * MMA_TJ + kTabDoubles);
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_cells_mma<KP, MT, MMA_TJ, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_zi_cells_mma<KP, MT, MMA_TJ, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_zi_cells_mma<KP, MT, MMA_TJ, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_zi_cells_mma<KP, MT, MMA_TJ, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int64_t groups = (a.p + 7) / 8;
    k_zi_pack<KP, true><<<pack_grid(groups * R::REC), 256, 0, st>>>(a.p, a.EVS, a.EVS, a.cz, a.flag, a.pack);
    const int64_t gx = (a.n + 32 * MT - 1) / (32 * MT);
    // at least ~8 waves of CTAs (3 resident per SM): split the gene tiles over blockIdx.y when there are few cell blocks
    const int ntiles = (int)((groups + MMA_TJ / 8 - 1) / (MMA_TJ / 8));
    int gy = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles / 8, (148 * 3 * 8 + gx - 1) / gx));
    if (gx >= 148 * 3 * 8) gy = 1;
    if (gy > 1) cudaMemsetAsync(a.PV, 0, sizeof(double) * (size_t)a.n * KP, st);
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (a.want_ll) {
        if (a.pstore) k_zi_cells_mma<KP, MT, MMA_TJ, true, true><<<grid, 128, smem, st>>>(a);
        else k_zi_cells_mma<KP, MT, MMA_TJ, false, true><<<grid, 128, smem, st>>>(a);
    } else {
        if (a.pstore) k_zi_cells_mma<KP, MT, MMA_TJ, true, false><<<grid, 128, smem, st>>>(a);
        else k_zi_cells_mma<KP, MT, MMA_TJ, false, false><<<grid, 128, smem, st>>>(a);
    }
    return 2;
}
template <int MT>
int launch_zi_genes_stored_t(const ZiBArgs &a, int chunks, cudaStream_t st) {
    using R = ZiRecF2<KP>;
    const size_t smem = sizeof(double) * (2 * (MMA_TI / 8) * R::REC);
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_genes_stored<KP, MT, MMA_TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int64_t groups = (a.n + 7) / 8;
    k_zi_pack_f2<KP><<<pack_grid(groups * R::REC), 256, 0, st>>>(a.n, a.EUn, a.pack);
    dim3 grid((a.p + 32 * MT - 1) / (32 * MT), chunks);
    k_zi_genes_stored<KP, MT, MMA_TI><<<grid, 128, smem, st>>>(a);
    return 2;
}
template <int MT>
int launch_zi_genes_mma_t(const ZiBArgs &a, int chunks, cudaStream_t st) {
    if (a.pstore) return launch_zi_genes_stored_t<MT>(a, chunks, st);
    using R = ZiRec<KP, false>;
    const size_t smem = sizeof(double) * (2 * (MMA_TI / 8) * R::REC + kTabDoubles);
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_genes_mma<KP, MT, MMA_TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int64_t groups = (a.n + 7) / 8;
    k_zi_pack<KP, false><<<pack_grid(groups * R::REC), 256, 0, st>>>(a.n, a.EUz, a.EUn, nullptr, nullptr, a.pack);
    dim3 grid((a.p + 32 * MT - 1) / (32 * MT), chunks);
    k_zi_genes_mma<KP, MT, MMA_TI><<<grid, 128, smem, st>>>(a);
    return 2;
}
#if PCMF_KP <= 32
// pass A with the second product as exact integer arithmetic on tcgen05 (zi_i8.cuh): factor widths up to 32
template <int NW>
int launch_zi_cells_i8_t(const ZiAArgs &a_in, cudaStream_t st) {
    using R = ZiRecI8<KP>;
    ZiAArgs a = a_in;
#ifdef PCMF_I8_ABLATE
    static const int abl = getenv("PCMF_I8_ABL") ? atoi(getenv("PCMF_I8_ABL")) : 0;
    const int want_ll = a.want_ll & 1;
    a.want_ll = want_ll | (abl << 8);
#else
    const int want_ll = a.want_ll;
#endif
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_cells_i8<KP, NW, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::SMEM);
        cudaFuncSetAttribute(k_zi_cells_i8<KP, NW, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::SMEM);
        cudaFuncSetAttribute(k_zi_cells_i8<KP, NW, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::SMEM);
        cudaFuncSetAttribute(k_zi_cells_i8<KP, NW, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::SMEM);
        attr = true;
    }
    const int64_t groups = (a.p + 7) / 8;
    double *tail = a.pack + R::off_tail(groups);
    cudaMemsetAsync(tail, 0, sizeof(double) * I8_TAIL, st);
    k_i8_colmax<KP><<<pack_grid((int64_t)a.p * KP), 256, 0, st>>>(a.p, a.EVS, reinterpret_cast<unsigned long long *>(tail));
    k_zi_pack_i8<KP, true><<<pack_grid(groups * R::NTOT * 8), 256, 0, st>>>(a.p, a.EVS, a.cz, a.flag, a.pack);
    const int64_t gx = (a.n + 127) / 128;
    // at least ~8 waves of CTAs (2 resident per SM): split the gene tiles over blockIdx.y when there are few cell blocks
    const int ntiles = (int)((groups + I8_TJ / 8 - 1) / (I8_TJ / 8));
    int gy = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles / 16, (148 * 2 * 8 + gx - 1) / gx));
    if (gx >= 148 * 2 * 8) gy = 1;
    if (gy > 1) cudaMemsetAsync(a.PV, 0, sizeof(double) * (size_t)a.n * KP, st);
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (want_ll) {
        if (a.pstore) k_zi_cells_i8<KP, NW, true, true><<<grid, 32 * NW, R::SMEM, st>>>(a);
        else k_zi_cells_i8<KP, NW, false, true><<<grid, 32 * NW, R::SMEM, st>>>(a);
    } else {
        if (a.pstore) k_zi_cells_i8<KP, NW, true, false><<<grid, 32 * NW, R::SMEM, st>>>(a);
        else k_zi_cells_i8<KP, NW, false, false><<<grid, 32 * NW, R::SMEM, st>>>(a);
    }
    return 3;
}
// pass B on the prob_D the integer cell-side pass stored (k_zi_genes_i8): 384 genes x <= 8192 cells per CTA
int launch_zi_genes_i8(ZiBArgs a, cudaStream_t st) {
    using R = ZiRecI8<KP>;
    using G = ZiGenesI8<KP>;
    static bool attr_dev[64] = {false};
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_genes_i8<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        attr = true;
    }
    const int64_t groups = (a.n + 7) / 8;
    double *tail = a.pack + R::off_tail(groups, false);
    cudaMemsetAsync(tail, 0, sizeof(double) * I8_TAIL, st);
    k_i8_colmax<KP><<<pack_grid(a.n * KP), 256, 0, st>>>(a.n, a.EUn, reinterpret_cast<unsigned long long *>(tail));
    k_zi_pack_i8<KP, false><<<pack_grid(groups * R::NTOT * 8), 256, 0, st>>>(a.n, a.EUn, nullptr, nullptr, a.pack);
    // cell chunks of at most 8192 cells (accumulator range), multiples of 32 (one stage = 4 cell groups); the gene blocks of
    // one chunk are adjacent in the grid, so the digit records they share are read from HBM once
    int64_t chunks = std::max<int64_t>(1, (a.n + I8B_MAXCELLS - 1) / I8B_MAXCELLS);
    int64_t rpc = (a.n + chunks - 1) / chunks; rpc = (rpc + 31) / 32 * 32;
    chunks = (a.n + rpc - 1) / rpc;
    a.rows_per_chunk = rpc;
    const int gx = (a.p + G::GB * 128 - 1) / (G::GB * 128);
    k_zi_genes_i8<KP><<<dim3((unsigned)gx, (unsigned)chunks), 128, G::SMEM, st>>>(a);
    return 3;
}
int launch_zi_cells_i8(const ZiAArgs &a, cudaStream_t st) {
    // (one warp per 32-cell row block: 8 warps per SM with ~190 registers each interleave the eight expit chains of a tile
    //  group; two warps per row block at 128 registers measured 3.5 ms slower at C4.  PCMF_I8_NW=8 selects the latter.)
    static const int nw = getenv("PCMF_I8_NW") ? atoi(getenv("PCMF_I8_NW")) : 4;
    return nw == 4 ? launch_zi_cells_i8_t<4>(a, st) : launch_zi_cells_i8_t<8>(a, st);
}
#endif
int launch_zi_cells_mma(const ZiAArgs &a, int variant, cudaStream_t st) {
#if PCMF_KP <= 32
    if (!(variant & (8 | 8192))) return launch_zi_cells_i8(a, st);
#endif
    if ((variant & 8) && MMA_MT == 4) return launch_zi_cells_mma_t<2>(a, st);
    return launch_zi_cells_mma_t<MMA_MT>(a, st);
}
int launch_zi_genes_mma(const ZiBArgs &a, int chunks, int variant, cudaStream_t st) {
#if PCMF_KP <= 32
    // the stored prob_D has the tile layout of the cell-side kernel that wrote it: the same switch picks both
    if (a.pstore && !(variant & (8 | 8192))) return launch_zi_genes_i8(a, st);
#endif
    if ((variant & 8) && MMA_MT == 4) return launch_zi_genes_mma_t<2>(a, chunks, st);
    return launch_zi_genes_mma_t<MMA_MT>(a, chunks, st);
}
void launch_zi_loglik(const ZiAArgs &a, double *out, cudaStream_t st) {
    using R = ZiRec<KP, true>;
    const size_t smem = sizeof(double) * (2 * (MMA_TJ / 8) * R::REC + 2 * MMA_TJ + kTabDoubles);
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_loglik_mma<KP, MMA_MT, MMA_TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int64_t grid = (a.n + 32 * MMA_MT - 1) / (32 * MMA_MT);
    // (the gene records are packed here as well: in FP32 mode the cell pass does not leave them behind)
    k_zi_pack<KP, true><<<pack_grid(((a.p + 7) / 8) * R::REC), 256, 0, st>>>(a.p, a.EVS, a.EVS, a.cz, a.flag, a.pack);
    k_zi_loglik_mma<KP, MMA_MT, MMA_TJ><<<(unsigned)grid, 128, smem, st>>>(a, out);
}
void launch_gene_loglam(bool f32, const LamArgs &a, int grid, cudaStream_t st) {
    if (f32) k_gene_loglam<KP, float><<<grid, 128, 0, st>>>(a);
    else k_gene_loglam<KP, double><<<grid, 128, 0, st>>>(a);
}
void launch_gene_order(bool f32, const OrderArgs &a, int grid, cudaStream_t st) {
    if (a.bound) {       // FP32 logarithms with an error bound per score (the host falls back to the exact kernel on a close call)
        if (f32) k_gene_order<KP, float, true><<<grid, 128, 0, st>>>(a);
        else k_gene_order<KP, double, true><<<grid, 128, 0, st>>>(a);
        return;
    }
    if (f32) k_gene_order<KP, float, false><<<grid, 128, 0, st>>>(a);
    else k_gene_order<KP, double, false><<<grid, 128, 0, st>>>(a);
}
size_t zi_pack_doubles(int64_t rows, int cells_side) {
    const int64_t groups = (rows + 7) / 8;
#if PCMF_KP <= 32
    if (!cells_side) return std::max((size_t)groups * ZiRec<KP, true>::REC, ZiRecI8<KP>::pack_doubles(groups));
#endif
#if PCMF_KP <= 32
    if (cells_side) return std::max((size_t)groups * ZiRec<KP, false>::REC, ZiRecI8<KP>::pack_doubles(groups, false));
#endif
    return (size_t)groups * (cells_side ? ZiRec<KP, false>::REC : ZiRec<KP, true>::REC);
}

int launch_zi_cells(const ZiAArgs &a, int variant, cudaStream_t st) {
    if (!(variant & 4)) return launch_zi_cells_mma(a, variant, st);
    const size_t smem = sizeof(double) * (ZI_TJ * KP + ZI_TJ + 4 * ZI_TJ + 4 * 32 * 33) + sizeof(int) * ZI_TJ;
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_pass_cells<KP, ZI_TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int64_t grid = (a.n + 127) / 128;
    k_zi_pass_cells<KP, ZI_TJ><<<(unsigned)grid, 128, smem, st>>>(a);
    return 1;
}

int launch_zi_genes(const ZiBArgs &a, int chunks, int variant, cudaStream_t st) {
    if (!(variant & 4)) return launch_zi_genes_mma(a, chunks, variant, st);
    const size_t smem = sizeof(double) * 2 * ZI_TI * KP;
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_pass_genes<KP, ZI_TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((a.p + 255) / 256, chunks);
    k_zi_pass_genes<KP, ZI_TI><<<grid, 128, smem, st>>>(a);
    return 1;
}

int zi_genes_block(int variant) { return !(variant & 4) ? (((variant & 8) && MMA_MT == 4) ? 64 : 32 * MMA_MT) : 256; }

// ---- FP32 mode: tcgen05 dense passes ---------------------------------------------------------------------------
constexpr int TC_KF1 = (KP + 7) / 8 * 8;            // factors of the first product (tf32 MMA: K in steps of 8)
constexpr int TC_KF2 = (KP + 15) / 16 * 16;         // accumulator columns of the second product (N in steps of 16 at M = 128)
using TcC = TcCfg<TC_KF1, TC_KF2>;

size_t tc_pack_bytes(int64_t records) { return (size_t)((records + TC_TG - 1) / TC_TG) * TcC::TILE_BYTES; }

template <bool CELLS>
void tc_attr() {
    static bool attr_dev[64] = {false};       // function attributes are per device (pcmf_options.ngpus drives several)
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_zi_tc<TC_KF1, TC_KF2, CELLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcC::SMEM);
        attr = true;
    }
}
int launch_zi_cells_tc(ZiTcArgs a, const double *M1, const double *M2, unsigned char *pack, int sms, cudaStream_t st) {
    (void)sms;
    tc_attr<true>();
    const int64_t ntiles = (a.streamed + TC_TG - 1) / TC_TG;
    k_tc_pack<TC_KF1, TC_KF2><<<pack_grid(ntiles * TC_TG * TC_KF2), 256, 0, st>>>(a.streamed, KP, M1, M2, a.cz, a.flag, pack);
    a.tiles = pack;
    a.tiles_per_chunk = ntiles;
    k_zi_tc<TC_KF1, TC_KF2, true><<<(unsigned)((a.rows + 127) / 128), 128, TcC::SMEM, st>>>(a);
    return 2;
}
int launch_zi_genes_tc(ZiTcArgs a, const double *M1, const double *M2, unsigned char *pack, int sms, cudaStream_t st) {
    tc_attr<false>();
    const int64_t ntiles = (a.streamed + TC_TG - 1) / TC_TG;
    k_tc_pack<TC_KF1, TC_KF2><<<pack_grid(ntiles * TC_TG * TC_KF2), 256, 0, st>>>(a.streamed, KP, M1, M2, nullptr, nullptr, pack);
    a.tiles = pack;
    // (gene block, cell chunk) CTAs for ~8 waves at the resident CTA count; the gene blocks of one chunk are adjacent in the
    // grid, so the cell tiles they all stream are read from HBM once and from L2 afterwards
    const int64_t gx = (a.rows + 127) / 128;
    const int resident = (TcC::SMEM <= 110 * 1024) ? 2 : 1;
    int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(ntiles, ((int64_t)sms * resident * 8 + gx - 1) / gx));
    a.tiles_per_chunk = (ntiles + chunks - 1) / chunks;
    chunks = (ntiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
    dim3 grid((unsigned)gx, (unsigned)chunks);
    k_zi_tc<TC_KF1, TC_KF2, false><<<grid, 128, TcC::SMEM, st>>>(a);
    return 2;
}

// ---- non-zero passes on the tiled copy of X --------------------------------------------------------------------
// Factor widths 16 and 20 only.  Narrower blocks: the tiled kernels' per-entry work does not shrink with K (C5, 200k x 10k at
// K = 2 / 10: gene pass 3.7 / 4.0 ms on the tiles against 1.3 / 2.1 ms on CSC; with only the cell pass tiled the scatter of
// q into tile order ate the gain), so they keep the CSR / CSC passes.  Wider blocks do not fit shared memory twice over.
#if PCMF_KP >= 16 && PCMF_KP <= 20
template <bool B2, bool LOGT>
void launch_tile_genes_t(TileGeneArgs a, int sms, cudaStream_t st) {
    using C = TileGeneCfg<KP, B2, LOGT>;
    static bool attr_dev[64] = {false};
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_tile_genes<KP, B2, LOGT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        attr = true;
    }
    // (gene block, range of cell blocks) CTAs for ~32 waves of one CTA per SM; the gene blocks of one range are adjacent in
    // the grid, so the cell rows they all stream come from HBM once and from L2 afterwards
    const int nI = a.x.nI, nJ = a.x.nJ;
    int chunks = std::max(1, std::min(nI, (sms * 32 + nJ - 1) / nJ));
    a.blocks_per_chunk = (nI + chunks - 1) / chunks;
    chunks = (nI + a.blocks_per_chunk - 1) / a.blocks_per_chunk;
    k_tile_genes<KP, B2, LOGT><<<dim3((unsigned)nJ, (unsigned)chunks), 512, C::SMEM, st>>>(a);
}
void launch_tile_genes(int mode, TileGeneArgs a, int sms, cudaStream_t st) {
    switch (mode) {
        case GENE_B1: launch_tile_genes_t<false, false>(a, sms, st); break;
        case GENE_B1_LOGT: launch_tile_genes_t<false, true>(a, sms, st); break;
        default: launch_tile_genes_t<true, false>(a, sms, st); break;
    }
}
void launch_tile_cells(const TileCellArgs &a, cudaStream_t st) {
    using C = TileCellCfg<KP>;
    static bool attr_dev[64] = {false};
    bool &attr = attr_dev[current_device()];
    if (!attr) {
        cudaFuncSetAttribute(k_tile_cells<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        attr = true;
    }
    k_tile_cells<KP><<<(unsigned)((a.x.nI + C::RB - 1) / C::RB), 512, C::SMEM, st>>>(a);
}
#define PCMF_TILE_LAUNCHERS launch_tile_genes, launch_tile_cells
#else
#define PCMF_TILE_LAUNCHERS nullptr, nullptr
#endif

const Launchers table = {KP, launch_estep_rows, launch_gene_pass, launch_zi_cells, launch_zi_genes, zi_genes_block, zi_pack_doubles,
                         launch_zi_loglik, launch_gene_loglam, launch_gene_order, tc_pack_bytes, launch_zi_cells_tc, launch_zi_genes_tc,
                         PCMF_TILE_LAUNCHERS};

}  // namespace

#define PCMF_CAT_(a, b) a##b
#define PCMF_CAT(a, b) PCMF_CAT_(a, b)
const Launchers *PCMF_CAT(get_launchers_, PCMF_KP)() { return &table; }

}  // namespace pcmf
