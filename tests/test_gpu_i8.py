"""The dense zero-inflation passes with their second product as exact integer arithmetic on tcgen05 (zi_i8.cuh) -- the
default of the FP64 mode for K <= 32 -- against the oracle and against the all-DMMA kernels (kernel_variant bit 8192).

prob_D enters those products as the 32-bit fixed-point word the cell-side pass stores (absolute error 1.2e-10 per entry), the
streamed factor matrix (E[V] o prob_S per gene, E[U] per cell) as six base-256 digits per factor column; everything summed
is an exact integer.  Tolerance against the oracle: RTOL = 1e-8 as everywhere in the FP64 mode (north_star: 1e-6); against
the all-DMMA path 1e-9.  Shapes are chosen to reach what small problems do not: the mid-pass read-out of the integer
accumulators (more than 4096 genes), several cell chunks in the gene-side pass (more than 8192 cells), ragged edges in both
directions, every padded factor width, whole-gene shortcuts, factor columns of very different scale.
"""
import numpy as np
import pytest

import pcmf_b200
from pcmf_b200 import datagen
from tests.test_gpu_parity import RTOL, compare_state, make_pair, problem, relerr

pytestmark = pytest.mark.gpu

I8, STORE, ALL_DMMA = 0, 64, 8192


def counts(seed, n, p, rate=2.5, keep=0.45):
    rng = np.random.default_rng(seed)
    X = rng.poisson(rate, (n, p)).astype(float) * (rng.uniform(size=(n, p)) < keep)
    X[X.sum(1) == 0, rng.integers(p)] = 1
    X[rng.integers(n), X.sum(0) == 0] = 1
    return X


@pytest.mark.parametrize("sparse", [False, True], ids=["zi", "zi_sparse"])
@pytest.mark.parametrize("n,p,K", [(131, 77, 2), (300, 190, 7), (257, 95, 12), (129, 250, 16), (200, 330, 20), (90, 140, 30), (64, 64, 32)])
def test_integer_passes_match_oracle_and_dmma(sparse, n, p, K):
    """Both integer passes (kernel_variant 64 keeps the stored prob_D on at these sizes) over five iterations."""
    X = counts(3 + n + p + K, n, p)
    m, s = make_pair(X, K, True, sparse, kernel_variant=I8 | STORE)
    _, sd = make_pair(X, K, True, sparse, kernel_variant=ALL_DMMA | STORE)
    for it in range(5):
        m.update_param()
        s.iterate(1)
        sd.iterate(1)
        compare_state(m, s, True, sparse, "integer passes, iter %d" % it)
        o, od = s.get_output(), sd.get_output()
        for k in ("a1", "a2", "b1", "b2", "prior_prob_D"):
            assert relerr(o[k], od[k]) < 1e-9, (it, k, relerr(o[k], od[k]))
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        assert abs(s.elbo() - sd.elbo()) < 1e-10 * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()
    sd.close()


def test_accumulator_readout_every_4096_genes():
    """More than 4096 genes: the cell-side pass reads its integer accumulators out mid-pass and adds the parts in FP64."""
    X = counts(5, 140, 4500, rate=1.5, keep=0.2)
    m, s = make_pair(X, 5, True, False, kernel_variant=STORE)
    for it in range(3):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, True, False, "4500 genes, iter %d" % it)
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()


def test_gene_side_pass_over_several_cell_chunks():
    """More than 8192 cells: a CTA of the gene-side pass owns at most 8192 (accumulator range), the chunks add atomically."""
    X = counts(6, 9100, 150, rate=1.5, keep=0.3)
    m, s = make_pair(X, 6, True, True, kernel_variant=STORE)
    for it in range(3):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, True, True, "9100 cells, iter %d" % it)
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()


def test_factor_columns_of_very_different_scale():
    """Digits are taken per factor column against the column's maximum: initial values whose columns differ by 1e8 and whose
    rows differ by 1e4 inside a column must come through at the FP64-mode tolerance."""
    n, p, K = 260, 210, 6
    X = problem(17, n, p, K_true=4)
    rng = np.random.default_rng(1)
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=2)
    a1 = a1 * (10.0 ** np.linspace(-4, 4, K))[None, :] * (10.0 ** rng.uniform(-2, 2, (n, 1)))
    b1 = b1 * (10.0 ** np.linspace(3, -3, K))[None, :] * (10.0 ** rng.uniform(-2, 2, (p, 1)))
    from oracle import pyoracle as po
    m = po.Model(X, K, True, False)
    m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X, K, True, False, a1=a1, a2=a2, b1=b1, b2=b2, kernel_variant=STORE)
    for it in range(4):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, True, False, "scaled columns, iter %d" % it)
        m.prepare_next_iterate()
    s.close()


def test_whole_gene_shortcuts_through_the_integer_passes():
    """Genes with prior_D = 1 (never zero) and genes that are zero almost everywhere: the +-1e300 logits of the packed records
    saturate through the clamped expit (zi_gap_factor_model.cpp:353-360)."""
    X = counts(9, 150, 90)
    X[:, :5] = np.maximum(X[:, :5], 1.0)          # prior_D = 1
    X[:, 5:9] = 0.0
    X[3, 5:9] = 1.0                               # a single non-zero each
    m, s = make_pair(X, 4, True, True, kernel_variant=STORE)
    for it in range(5):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, True, True, "shortcut genes, iter %d" % it)
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()
