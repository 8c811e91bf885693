"""Input formats and preprocessing either side of the iteration, through the C ABI on the GPU:
dgCMatrix (compressed-column) input, the prefilter() statistics pass, pCMF() on a sparse input, and the non-zero-only
initialisation of the zero-inflation state against the dense pass it replaces."""
import numpy as np
import pytest
import scipy.sparse as sp

import pcmf_b200
from oracle import pyoracle as po
from pcmf_b200 import datagen
from tests.test_gpu_parity import IDS, MODELS, RTOL, compare_state, make_pair, problem, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_dgcmatrix_input_matches_dense_input(zi, sparse):
    """A compressed-column matrix (with explicit zeros, unsorted-by-cell order is the format's nature) gives the same
    state as the dense matrix it stands for; both match the oracle."""
    X = problem(41, 130, 75)
    K = 4
    Xc = sp.csc_matrix(X.astype(np.float64))
    # plant explicit zeros: they must not be taken for non-null entries
    data = Xc.data.copy()
    data[::7] = 0.0
    Xz = sp.csc_matrix((data, Xc.indices.copy(), Xc.indptr.copy()), shape=Xc.shape)
    Xd = Xz.toarray()
    keep_rows, keep_cols = Xd.sum(1) > 0, Xd.sum(0) > 0
    assert keep_rows.all() and keep_cols.all()
    a1, a2, b1, b2 = datagen.random_init(Xd, K, seed=5)
    prob_S, prior_S = datagen.prior_S_heuristic(Xd, K)
    kw = dict(a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S)
    m = po.Model(Xd, K, zi, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(Xz, K, zi, sparse, **kw)
    assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
    for _ in range(4):
        m.update_param()
        m.prepare_next_iterate()
    s.iterate(4)
    compare_state(m, s, zi, sparse, "csc input")
    s.close()


def test_dgcmatrix_large_values_keep_fp64():
    """Values that are not exact in FP32 switch the value arrays to FP64 (same rule as the dense and CSR inputs)."""
    X = problem(42, 60, 40).astype(np.float64)
    X[X > 0] += 2.0 ** 25 + 1.0
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=6)
    m = po.Model(X, K, False, False)
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(sp.csc_matrix(X), K, False, False, a1=a1, a2=a2, b1=b1, b2=b2)
    for _ in range(2):
        m.update_param()
        m.prepare_next_iterate()
    s.iterate(2)
    compare_state(m, s, False, False, "csc fp64 values")
    s.close()


def test_dgcmatrix_bad_row_index_is_an_error():
    X = sp.csc_matrix(problem(43, 30, 20).astype(np.float64))
    X.indices[3] = 30
    with pytest.raises(pcmf_b200.PcmfError):
        pcmf_b200.Solver(X, 2, False, False)


@pytest.mark.parametrize("presel", [False, True])
def test_prefilter_sparse_matches_dense_restatement(presel):
    """prefilter() on CSR / device-resident input (GPU statistics pass) against the column-by-column restatement of
    prefiltering.R on the dense matrix."""
    X, _, _ = datagen.simulate(400, 300, 5, signal_U=8.0, signal_V=8.0, prob1=0.15, seed=9)
    X[:, 5] = 0
    X[:, 11] = 0
    X[3, 17] = 5000                                        # an extreme column for the quantile criterion
    for kw in (dict(), dict(prop=0.1), dict(quant_max=0.95), dict(prop=0.02, quant_max=0.9, threshold=0.4)):
        ref = po.prefilter(X, presel=presel, **kw)
        got = pcmf_b200.prefilter(sp.csr_matrix(X.astype(np.float64)), presel=presel, **kw)
        dense = pcmf_b200.prefilter(X, presel=presel, **kw)
        assert (got == ref).all(), kw
        assert (dense == ref).all(), kw
        assert 0 < ref.sum() < X.shape[1]
    s1, s2, cn, mx = pcmf_b200.csr_column_stats(sp.csr_matrix(X.astype(np.float64)), with_max=True)
    assert (mx == X.max(0)).all() and (cn == (X != 0).sum(0)).all() and (s1 == X.sum(0)).all()


def test_pcmf_on_sparse_input_matches_dense_input():
    """pCMF(X = sparse) derives prior_S from GPU column statistics; same fit as pCMF(X = dense)."""
    X = problem(44, 100, 60)
    kw = dict(verbose=False, iter_max=12, iter_min=6, seed=3, reorder_factor=True)
    rd = pcmf_b200.pCMF(X, 3, **kw)
    rs = pcmf_b200.pCMF(sp.csr_matrix(X.astype(np.float64)), 3, **kw)
    for grp in ("variational_params", "stats"):
        for k in rd[grp]:
            assert relerr(rs[grp][k], rd[grp][k]) < 1e-7, (grp, k)
    assert relerr(rs["sparse_param"]["prior_prob_S"], rd["sparse_param"]["prior_prob_S"]) < 1e-7


@pytest.mark.parametrize("sparse", [False, True], ids=["zi", "zi_sparse"])
@pytest.mark.parametrize("K", [3, 20, 30])
def test_zi_init_shortcut_matches_dense_pass(sparse, K):
    """init_zi_param's prob_D = (X > 0) realised on the non-zeros only (default) against the dense pass with c = -inf
    (kernel_variant bit 32): same ELBO at the initial state, same state after one and three iterations."""
    X = problem(45, 170, 90, K_true=5)
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=4)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S)
    s0 = pcmf_b200.Solver(X, K, True, sparse, **kw)
    s1 = pcmf_b200.Solver(X, K, True, sparse, kernel_variant=32, **kw)
    assert abs(s0.elbo() - s1.elbo()) < 1e-11 * abs(s1.elbo())
    for it in (1, 2):
        s0.iterate(it)
        s1.iterate(it)
        o0, o1 = s0.get_output(), s1.get_output()
        for k in ("a1", "a2", "b1", "b2", "prior_prob_D"):
            assert relerr(o0[k], o1[k]) < 1e-11, (it, k)
    s0.close()
    s1.close()


def test_malformed_csr_input_is_rejected():
    """host / device CSR is taken as it is, so what the kernels rely on is checked on the device first."""
    rp = np.array([0, 2, 3, 5], np.int64)
    ci = np.array([0, 3, 1, 2, 4], np.int32)
    vv = np.array([1, 2, 3, 4, 5], np.float32)
    ok = pcmf_b200.Solver((rp, ci, vv, 5), 2, False, False, seed=1)
    ok.close()
    for bad, msg in (((rp, np.array([0, 3, 1, 2, 7], np.int32), vv, 5), "gene index out of range"),
                     ((np.array([0, 3, 2, 5], np.int64), ci, vv, 5), "row pointers"),
                     ((rp, ci, np.array([1, 0, 3, 4, 5], np.float32), 5), "explicit zeros")):
        with pytest.raises(pcmf_b200.PcmfError, match=msg) as e:
            pcmf_b200.Solver(bad, 2, False, False, seed=1)
        assert e.value.code == 1
