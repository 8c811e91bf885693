// Per-KP kernel launch tables.  Every padded factor width KP is compiled in its own translation unit
// (kernels_kp.cu with -DPCMF_KP=<n>) so the build parallelises; the solver picks a table at run time.
#pragma once
#include "kernels.cuh"
#include "zi_tc.cuh"
#include "xtile.cuh"

namespace pcmf {

enum GeneMode : int { GENE_B1 = 0, GENE_B1_LOGT = 1, GENE_B1_B2 = 2 };

struct Launchers {
    int KP;
    void (*estep_rows)(bool f32, const RowArgs &, int grid, int block, cudaStream_t);
    void (*gene_pass)(bool f32, int mode, const ColArgs &, int grid, cudaStream_t);
    int (*zi_cells)(const ZiAArgs &, int variant, cudaStream_t);                // returns the number of launches
    int (*zi_genes)(const ZiBArgs &, int chunks, int variant, cudaStream_t);    // returns the number of launches
    int (*zi_genes_block)(int variant);   // genes per CTA of the gene-side ZI pass (for the chunk count)
    size_t (*zi_pack_doubles)(int64_t rows, int cells_side);
    void (*zi_loglik)(const ZiAArgs &, double *out, cudaStream_t);            // zero part of zi_poisson_loglike (monitor)
    void (*gene_loglam)(bool f32, const LamArgs &, int grid, cudaStream_t);   // deviance / log-likelihood non-zero pass
    void (*gene_order)(bool f32, const OrderArgs &, int grid, cudaStream_t);  // one round of the greedy factor ordering   // scratch size of ZiAArgs::pack (0) / ZiBArgs::pack (1)
    // FP32 mode: the dense passes on tcgen05 (zi_tc.cuh).  M1 / M2: row-major FP64 matrices of the streamed side for the first /
    // second product; the launcher packs them into ZiTcArgs::tiles (tc_pack_bytes(records) bytes) and runs the pass
    size_t (*tc_pack_bytes)(int64_t records);
    int (*zi_cells_tc)(ZiTcArgs, const double *M1, const double *M2, unsigned char *pack, int sms, cudaStream_t);
    int (*zi_genes_tc)(ZiTcArgs, const double *M1, const double *M2, unsigned char *pack, int sms, cudaStream_t);
    // non-zero passes on the 2-D tiled copy of X (xtile.cuh); null for factor widths that do not fit (KP > XT_MAXKP)
    void (*tile_genes)(int mode, TileGeneArgs, int sms, cudaStream_t);
    void (*tile_cells)(const TileCellArgs &, cudaStream_t);
};

const Launchers *get_launchers_4();
const Launchers *get_launchers_8();
const Launchers *get_launchers_12();
const Launchers *get_launchers_16();
const Launchers *get_launchers_20();
const Launchers *get_launchers_32();
const Launchers *get_launchers_64();

inline int padded_K(int K) {
    const int w[] = {4, 8, 12, 16, 20, 32, 64};
    for (int x : w) if (K <= x) return x;
    return -1;
}
inline const Launchers *get_launchers(int KP) {
    switch (KP) {
        case 4: return get_launchers_4();
        case 8: return get_launchers_8();
        case 12: return get_launchers_12();
        case 16: return get_launchers_16();
        case 20: return get_launchers_20();
        case 32: return get_launchers_32();
        case 64: return get_launchers_64();
    }
    return nullptr;
}

}  // namespace pcmf
