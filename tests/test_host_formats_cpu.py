"""Host-side marshalling of the input formats and prefilter() on dense input (no GPU involved)."""
import numpy as np
import pytest
import scipy.sparse as sp

import pcmf_b200
from oracle import pyoracle as po
from pcmf_b200 import api, datagen


def test_csc_input_crosses_the_abi_as_dgcmatrix_slots():
    X = sp.random(40, 30, density=0.2, format="csc", random_state=1, data_rvs=lambda k: np.random.default_rng(2).integers(1, 9, k).astype(float))
    pr, keep, n, p = api._as_problem(X)
    assert (n, p) == (40, 30) and pr.n_local == 40 and pr.n_global == 40 and pr.p == 30
    assert pr.csc_colptr and pr.csc_rowidx and pr.csc_val_f64 and not pr.csr_rowptr and not pr.dense_f64 and not pr.dense_i32
    cp, ri, xv = keep
    assert cp.dtype == np.int32 and ri.dtype == np.int32 and xv.dtype == np.float64
    assert (cp == X.indptr).all() and (ri == X.indices).all() and (xv == X.data).all()


def test_sparse_inputs_are_made_canonical():
    rows, cols, vals = np.array([0, 0, 1, 2, 2]), np.array([1, 1, 0, 2, 2]), np.array([1.0, 2.0, 3.0, 4.0, 0.0])
    coo = sp.coo_matrix((vals, (rows, cols)), shape=(3, 3))
    csc = sp.csc_matrix((np.array([3.0, 1.0, 2.0, 4.0]), np.array([1, 0, 0, 2]), np.array([0, 1, 3, 4])), shape=(3, 3))   # (0,1) twice
    _, keep, _, _ = api._as_problem(csc)
    assert list(keep[0]) == [0, 1, 2, 3] and list(keep[2]) == [3.0, 3.0, 4.0]                     # duplicates summed
    pr, keep, _, _ = api._as_problem(coo)                                                       # -> CSR, duplicates summed, zeros dropped
    assert pr.csr_rowptr and list(keep[0]) == [0, 1, 2, 3] and list(keep[1]) == [1, 0, 2] and list(keep[2]) == [3.0, 3.0, 4.0]


def test_dense_integer_and_real_inputs_keep_their_r_types():
    Xi = np.arange(12, dtype=np.int64).reshape(4, 3)
    pr, keep, n, p = api._as_problem(Xi.astype(np.int32))
    assert pr.dense_i32 and not pr.dense_f64 and keep[0].flags.f_contiguous          # INTSXP
    pr, keep, n, p = api._as_problem(Xi.astype(float))
    assert pr.dense_f64 and not pr.dense_i32 and keep[0].flags.f_contiguous          # REALSXP
    with pytest.raises(pcmf_b200.PcmfError):
        api._as_problem(np.array([["a", "b"], ["c", "d"]]))


@pytest.mark.parametrize("kw", [dict(), dict(prop=0.2), dict(quant_max=0.9), dict(presel=True), dict(presel=True, threshold=0.5, prop=0.0)])
def test_prefilter_on_a_dense_matrix_matches_the_r_restatement(kw):
    X, _, _ = datagen.simulate(200, 120, 4, signal_U=6.0, signal_V=6.0, prob1=0.2, seed=3)
    X[:, 7] = 0
    X[5, 9] = 10000
    assert (pcmf_b200.prefilter(X, **kw) == po.prefilter(X, **kw)).all()


def test_prefilter_rejects_arguments_outside_the_unit_interval():
    X = np.ones((4, 3))
    for bad in (dict(prop=1.5), dict(quant_max=-0.1), dict(threshold=2)):
        with pytest.raises(pcmf_b200.PcmfError):
            pcmf_b200.prefilter(X, **bad)
