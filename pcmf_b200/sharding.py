"""Cell (row) sharding of the count matrix over the GPUs of one box.

The reference parallelises over genes with OpenMP thread-private accumulators that are summed at the end
(gap_factor_model.cpp:474-478).  Here cells are split into contiguous ranges, one per rank; everything indexed by a
cell (a1, a2, E[U], E[log U], the ZI row products) stays local, everything indexed by a gene is replicated and updated
identically on every rank from all-reduced sums.  Ranges are balanced by non-zeros, not by row count,
because the non-zero passes dominate the sparse models and real cells differ widely in depth.
"""
import numpy as np


def balanced_row_partition(rowptr, world):
    """Contiguous row ranges [(r0, r1), ...] whose non-zero counts are as equal as a contiguous split allows.
    `rowptr` is the CSR row-pointer array (len n+1).  Every rank gets at least one row when n >= world."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = len(rowptr) - 1
    if world <= 0:
        raise ValueError("world must be positive")
    if n < world:
        raise ValueError("fewer rows than ranks")
    nnz = rowptr[-1] - rowptr[0]
    bounds = [0]
    for r in range(1, world):
        target = rowptr[0] + nnz * r / world
        cut = int(np.searchsorted(rowptr, target, side="left"))
        cut = max(cut, bounds[-1] + 1)          # at least one row per rank
        cut = min(cut, n - (world - r))         # leave one row for every remaining rank
        bounds.append(cut)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def equal_row_partition(n, world):
    """Contiguous ranges with equal row counts (exchangeable synthetic cells)."""
    base, rem = divmod(n, world)
    out, r0 = [], 0
    for r in range(world):
        r1 = r0 + base + (1 if r < rem else 0)
        out.append((r0, r1))
        r0 = r1
    return out


def shard_csr(rowptr, colidx, val, r0, r1):
    """The CSR slice of rows [r0, r1) with a re-based row pointer."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    e0, e1 = rowptr[r0], rowptr[r1]
    return rowptr[r0:r1 + 1] - e0, colidx[e0:e1], val[e0:e1]
