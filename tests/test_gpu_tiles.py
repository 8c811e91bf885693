"""The non-zero passes of the E-step on the 2-D tiled copy of X (pcmf_b200/csrc/xtile.cuh, xtile_kernels.cuh), forced on
small problems with kernel_variant bit 2048 (the default switches them on from 65536 cells; tests/test_gpu_scale_parity.py
covers that).  Same oracle comparison as tests/test_gpu_parity.py: gap_factor_model.cpp:458-499 and the variants
zi...:289-330, sparse...:327-368, zi_sparse...:316-357; spike-and-slab pass sparse...:424-471."""
import numpy as np
import pytest

import pcmf_b200
from oracle import pyoracle as po
from pcmf_b200 import datagen
from tests.test_gpu_parity import IDS, MODELS, RTOL, compare_state, make_pair, problem, relerr

pytestmark = pytest.mark.gpu

TILES = 2048


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
# several cell blocks (128) and gene blocks (256), partial last blocks, both factor widths the tiles are compiled for (16, 20)
@pytest.mark.parametrize("n,p,K", [(60, 45, 13), (300, 530, 20), (517, 260, 16), (130, 257, 17), (260, 300, 14), (1000, 700, 20)])
def test_tiled_passes_match_oracle(zi, sparse, n, p, K):
    X = problem(3, n, p)
    m, s = make_pair(X, K, zi, sparse, kernel_variant=TILES)
    for it in range(5):
        m.update_param()
        crit = s.iterate(1)[0]
        _, ng = m.gap_between_iterates()
        assert abs(crit - ng) < RTOL * ng, (it, crit, ng)
        compare_state(m, s, zi, sparse, "iter %d" % it)
        e_gpu, terms = s.elbo(terms=True)
        e_cpu = m.elbo()
        ot = m.elbo_terms()
        for k in ot:
            assert abs(terms[k] - ot[k]) < RTOL * max(1.0, abs(ot[k])), (it, k, terms[k], ot[k])
        assert abs(e_gpu - e_cpu) < RTOL * abs(e_cpu), (it, e_gpu, e_cpu)
        m.prepare_next_iterate()
    s.close()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_tiled_passes_equal_csr_csc_passes(zi, sparse):
    """Same fit with the tiles and with the CSR / CSC passes (kernel_variant bit 4096), ELBO recorded every iteration (the
    late data term of the gene pass: B2 / x log T variants) -- agreement at rounding level."""
    X = problem(21, 700, 600, prob1=0.3)
    K = 20
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=4)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    outs = []
    for kv in (TILES, 4096):
        s = pcmf_b200.Solver(X, K, zi, sparse, a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S, kernel_variant=kv)
        crit, elbo = s.iterate(6, with_elbo=True)
        outs.append((np.array(crit), np.array(elbo), s.get_output(want_prob_D=False)))
        s.close()
    (c0, e0, o0), (c1, e1, o1) = outs
    assert relerr(c0, c1) < 1e-10 and relerr(e0, e1) < 1e-12
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(o0[k], o1[k]) < 1e-10, k
    if sparse:
        assert (o0["S"] == o1["S"]).all()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_tiled_guard_path_matches_oracle(zi, sparse):
    """The +-100 guards (gap...cpp:483, internal.h:121-127) inside the tiled kernels: entries that need the reference's
    per-entry formula are evaluated by their group of 4 lanes and marked for the cell pass (q = -x)."""
    n, p, K = 200, 300, 14
    X = problem(9, n, p)
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=2)
    a1[0, 0] = 1e60
    a1[1, :] = 1e-3
    a1[2, 1] = 5e-3
    b1[3, 2] = 2e-3
    a1[7, 13] = 1e55
    a1[150, :] = 1e-3
    b1[270, :] = 3e-3
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    m = po.Model(X, K, zi, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X, K, zi, sparse, a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S, kernel_variant=TILES)
    for it in range(3):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, zi, sparse, "guard iter %d" % it)
        e_gpu, e_cpu = s.elbo(), m.elbo()
        assert abs(e_gpu - e_cpu) < RTOL * abs(e_cpu), (it, e_gpu, e_cpu)
        m.prepare_next_iterate()
    s.close()


def test_tiled_run_driver_with_monitors_and_reorder():
    X = problem(5, 300, 280)
    K = 14
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=7)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=40, iter_min=10, epsilon=1e-2, additional_iter=3, a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S)
    ref = po.run(X, K, True, True, monitor=True, reorder_factor=True, **kw)
    res = pcmf_b200.run_zi_sparse_gap_factor(X, K, verbose=False, monitor=True, reorder_factor=True, kernel_variant=TILES, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert list(res["factor_order"]) == list(ref["factor_order"])
    for k in ("optim_criterion", "loglikelihood", "deviance"):
        assert relerr(res["monitor"][k], ref["monitor"][k]) < 1e-6, k
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k


@pytest.mark.parametrize("sparse", [False, True], ids=["zi", "zi_sparse"])
def test_tiled_warm_start_from_dense_prob_D_and_multi_init(sparse):
    """States the tiles must hand over to / take back from the CSR / CSC passes: the first iteration after a warm start from
    an explicit dense prob_D runs on the weighted non-zeros (wrapper_sparse_gap_factor.h:394-398), the closed form and the
    tiles from then on; a multi-init (wrapper_gap_factor.h:296-350) restores a saved state, after which the stored q is stale."""
    X = problem(25, 300, 280)
    K = 14
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=12)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    rng = np.random.default_rng(3)
    PD, prD = rng.uniform(0.05, 0.95, X.shape), np.full(X.shape[1], 0.5)
    kw = dict(iter_max=8, iter_min=8, epsilon=1e-9, additional_iter=2, a1=a1, a2=a2, b1=b1, b2=b2, prob_D=PD, prior_D=prD)
    ref = po.run(X, K, True, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    fn = pcmf_b200.run_zi_sparse_gap_factor if sparse else pcmf_b200.run_zi_gap_factor
    res = fn(X, K, verbose=False, monitor=True, reorder_factor=False, kernel_variant=TILES, **kw)
    for k in ("optim_criterion", "loglikelihood", "deviance"):
        assert relerr(res["monitor"][k], ref["monitor"][k]) < 1e-6, k
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k
    # multi-init from seeded random initialisations
    kw2 = dict(iter_max=10, iter_min=10, epsilon=1e-9, additional_iter=2, ninit=3, iter_init=4, seed=5)
    ref = po.run(X, K, True, sparse, monitor=True, reorder_factor=True, prob_S=prob_S, prior_S=prior_S, **kw2)
    if sparse:
        kw2.update(prob_S=prob_S, prior_S=prior_S)
    res = fn(X, K, verbose=False, monitor=True, reorder_factor=True, kernel_variant=TILES, **kw2)
    assert list(res["factor_order"]) == list(ref["factor_order"])
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k
    assert abs(res["loss"] - ref["loss"]) < 1e-6 * abs(ref["loss"])
