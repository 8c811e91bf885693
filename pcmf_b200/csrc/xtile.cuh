// 2-D tiled copy of X for the non-zero passes of the E-step (north_star item 1: fused SDDMM + SpMM with the U / V tiles
// staged in shared memory).
//
// The CSR / CSC passes of kernels.cuh gather one K-vector of the OTHER side per non-zero (160 B at K = 20 in FP64, twice
// that in the spike-and-slab pass) through L2, whose request rate caps them at ~9 TB/s -- 3-5 % of the HBM roofline for 8
// algorithmic bytes per non-zero (profiles/r01_*).  Here X is cut into tiles of XT_TI cells x XT_TJ genes; a CTA keeps the
// factor rows of both sides of a tile in shared memory (bulk-copied, double buffered) and the per-non-zero gathers become
// shared-memory reads.  What streams from HBM per non-zero is 8 B (gene pass: local cell, position, count) + 8 B (q written),
// resp. 9 B (cell pass: local gene, q read).
//
// Layout.  Tile t = I * nJ + J holds the non-zeros of cells [128 I, 128 I + 128) x genes [256 J, 256 J + 256), stored twice:
//   gene order (sorted by local gene, then local cell): grec[e] = { local cell | position of the entry in cell order << 8,
//       count as float };  gofs[t][c] = first entry of local gene c, gperm[t][.] = local genes by decreasing length
//   cell order (sorted by local cell, then local gene): ccol[e] = local gene, q[e] = x / T left by the last gene pass
//       (0: gene with every factor masked; -x: the entry needs the reference's guarded formula, gap_factor_model.cpp:481-487);
//       cofs[t][r] = first entry of local cell r, cperm[t][.] = local cells by decreasing length
// A warp works on 8 segments at a time, one per group of 4 lanes (the K factors are split over the 4 lanes, the partial dot
// products meet through two shuffles); the segments of a tile are handed out longest first, so the 8 of a round have about
// the same length and the last rounds of a tile are the short ones.
#pragma once
#include "kernels.cuh"

namespace pcmf {

constexpr int XT_TI = 128, XT_TJ = 256;
constexpr int XT_MINKP = 16, XT_MAXKP = 20;   // narrower factor blocks: no gain (kernels_kp.cu); wider: do not fit shared memory

struct XTile {
    int32_t nI, nJ;
    const int64_t *tptr;              // nI nJ + 1
    const uint2 *grec;                // nnz, gene order
    const uint16_t *gofs;             // nT x (XT_TJ + 1)
    const uint8_t *gperm;             // nT x XT_TJ
    const uint8_t *ccol;              // nnz, cell order
    double *q;                        // nnz, cell order
    const uint16_t *cofs;             // nT x (XT_TI + 1)
    const uint8_t *cperm;             // nT x XT_TI
};

struct TileGeneArgs { ColArgs c; XTile x; int64_t n; int blocks_per_chunk; const double2 *pairs; };   // pairs: (eu, eu o ElogU), n x KP (B2 mode)
struct TileCellArgs { RowArgs r; XTile x; int32_t p; };

#ifdef PCMF_COMMON_KERNELS
// ---- builder (once per fit) ---------------------------------------------------------------------------------
// Runs of equal block index inside a sorted segment (a cell's genes / a gene's cells): the count of every (segment, block)
// pair goes to slot [tile][local index + 1] of the offset table, which k_xt_scan then turns into offsets.
// ROWS: also runofs[J * nseg + i] = offset, inside cell i's CSR segment, of its first entry in gene block J (the gene-order
// scatter turns a CSR entry index into a cell-order position with it).
template <bool ROWS>
__global__ void k_xt_count(int64_t nseg, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, int nJ, uint16_t *__restrict__ ofs,
                           uint32_t *__restrict__ runofs) {
    constexpr int SH = ROWS ? 8 : 7, L = ROWS ? XT_TI : XT_TJ;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t sg = warp; sg < nseg; sg += nw) {
        const int64_t e0 = ptr[sg], e1 = ptr[sg + 1];
        int carryB = -1, carryCnt = 0;
        for (int64_t eb = e0; eb < e1; eb += 32) {
            const int64_t e = eb + lane;
            const bool valid = e < e1;
            const int B = valid ? (idx[e] >> SH) : -2;
            const unsigned mask = __match_any_sync(0xffffffffu, B);
            const int nextB = (eb + 32 < e1) ? (idx[eb + 32] >> SH) : -3;
            const int total = __popc(mask) + (B == carryB ? carryCnt : 0);
            const bool leader = lane == 31 - __clz(mask);
            if (valid && leader && B != nextB) {
                const int64_t t = ROWS ? (sg >> 7) * nJ + B : (int64_t)B * nJ + (sg >> 8);
                const int loc = ROWS ? (int)(sg & 127) : (int)(sg & 255);
                ofs[t * (L + 1) + loc + 1] = (uint16_t)total;
                if (ROWS) runofs[(int64_t)B * nseg + sg] = (uint32_t)(e + 1 - total - e0);      // the run ends at this entry
            }
            const int lastB = __shfl_sync(0xffffffffu, B, 31), lastTot = __shfl_sync(0xffffffffu, total, 31);
            if (lastB == nextB) { carryB = lastB; carryCnt = lastTot; } else { carryB = -1; carryCnt = 0; }
        }
    }
}
// counts -> offsets inside every tile (one warp per tile); the tile's total goes to tcnt
template <int L>
__global__ void k_xt_scan(int64_t nT, uint16_t *__restrict__ ofs, int64_t *__restrict__ tcnt) {
    constexpr int PER = L / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < nT; t += nw) {
        uint16_t *o = ofs + t * (L + 1);
        int v[PER], s = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = o[1 + lane * PER + i]; s += v[i]; }
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
        int run = incl - s;
        if (lane == 0) o[0] = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { run += v[i]; o[1 + lane * PER + i] = (uint16_t)run; }
        if (tcnt && lane == 31) tcnt[t] = incl;
    }
}
// segments of a tile by decreasing length (ties by index): perm[rank] = local index.  One warp per tile.
template <int L>
__global__ void k_xt_perm(int64_t nT, const uint16_t *__restrict__ ofs, uint8_t *__restrict__ perm) {
    __shared__ uint16_t len[4][L];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < nT; t += nw) {
        const uint16_t *o = ofs + t * (L + 1);
        __syncwarp();
        for (int c = lane; c < L; c += 32) len[w][c] = (uint16_t)(o[c + 1] - o[c]);
        __syncwarp();
        for (int c = lane; c < L; c += 32) {
            const int mine = len[w][c];
            int rank = 0;
            for (int d = 0; d < L; ++d) { const int other = len[w][d]; rank += (other > mine) || (other == mine && d < c); }
            perm[t * L + rank] = (uint8_t)c;
        }
    }
}
// cell order: the local gene of every CSR entry at its position inside its tile (tile offset of the cell + rank inside the
// run of the gene block)
__global__ void k_xt_scatter_rows(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, int nJ,
                                  const int64_t *__restrict__ tptr, const uint16_t *__restrict__ cofs, const uint32_t *__restrict__ runofs,
                                  uint8_t *__restrict__ ccol) {
    const int64_t total = rowptr[n];
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const int64_t e0 = rowptr[i], e1 = rowptr[i + 1];
        const int r = (int)(i & 127);
        for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const int j = colidx[e];
            const int64_t t = (i >> 7) * nJ + (j >> 8);
            const int pos = cofs[t * (XT_TI + 1) + r] + (int)(e - e0 - runofs[(int64_t)(j >> 8) * n + i]);
            ccol[tptr[t] + pos] = (uint8_t)(j & 255);
        }
    }
    (void)total;
}
// (eu, eu o ElogU) pairs of every cell: the rows the spike-and-slab gene pass streams (one 16-byte read per factor)
__global__ void k_eu_pairs(int64_t total, const double *__restrict__ eu, const double *__restrict__ ElogU, double2 *__restrict__ out) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const double u = eu[o];
        out[o] = make_double2(u, u * ElogU[o]);
    }
}
// gene order: the record of every CSC entry.  csr_of[f] is the CSR index of CSC entry f (the permutation of the CSC sort), so
// its cell-order position is the cell's tile offset + its distance from the start of the cell's run in this gene block --
// no CSR -> CSC map and no second scatter.  The tables are read with locality: a gene walks its cells upwards, and the 256
// genes of a block walk the same slices of cofs and runofs.
__global__ void k_xt_scatter_cols(int32_t p, int64_t n, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                  const float *__restrict__ val_csc, const int64_t *__restrict__ csr_of, const int64_t *__restrict__ rowptr,
                                  int nJ, const int64_t *__restrict__ tptr, const uint16_t *__restrict__ gofs,
                                  const uint16_t *__restrict__ cofs, const uint32_t *__restrict__ runofs, uint2 *__restrict__ grec,
                                  uint32_t *__restrict__ qmap) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < p; j += nw) {
        const int64_t e0 = colptr[j], e1 = colptr[j + 1];
        const int c = (int)(j & 255);
        const int64_t Jn = (j >> 8) * n;
        int carryB = -1, carryCnt = 0;
        for (int64_t eb = e0; eb < e1; eb += 32) {
            const int64_t e = eb + lane;
            const bool valid = e < e1;
            const int i = valid ? rowidx[e] : 0;
            const int B = valid ? (i >> 7) : -2;
            const unsigned mask = __match_any_sync(0xffffffffu, B);
            const int nextB = (eb + 32 < e1) ? (rowidx[eb + 32] >> 7) : -3;
            const int before = (B == carryB ? carryCnt : 0);
            const int rank = __popc(mask & ((1u << lane) - 1u)) + before;
            if (valid) {
                const int64_t t = (int64_t)B * nJ + (j >> 8);
                const int64_t tb = tptr[t];
                const unsigned pc = cofs[t * (XT_TI + 1) + (i & 127)] + (unsigned)(csr_of[e] - rowptr[i] - runofs[Jn + i]);
                if (grec) {       // (null: this factor width keeps the CSC gene pass, only the map is needed)
                    const int pos = gofs[t * (XT_TJ + 1) + c] + rank;
                    grec[tb + pos] = make_uint2((unsigned)(i & 127) | (pc << 8), __float_as_uint(val_csc[e]));
                }
                qmap[e] = (uint32_t)(tb + pc);        // where the cell pass reads the q of this CSC entry
            }
            const int lastB = __shfl_sync(0xffffffffu, B, 31), lastTot = __shfl_sync(0xffffffffu, __popc(mask) + before, 31);
            if (lastB == nextB) { carryB = lastB; carryCnt = lastTot; } else { carryB = -1; carryCnt = 0; }
        }
    }
}
#endif  // PCMF_COMMON_KERNELS

}  // namespace pcmf
