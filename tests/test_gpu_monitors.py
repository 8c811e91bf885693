"""GPU parity of the components around the iteration through the C ABI, against the oracle:
monitors (log-likelihood, deviance, explained deviance), the greedy factor ordering + reorder, and the multi-init loop.
Tolerance: 1e-8 relative on scalars (north_star allows 1e-6), exact equality on the chosen factor order."""
import numpy as np
import pytest

import pcmf_b200
from oracle import pyoracle as po
from pcmf_b200 import datagen
from tests.test_gpu_parity import IDS, MODELS, compare_state, make_pair, problem, relerr

pytestmark = pytest.mark.gpu

RUN = {(False, False): pcmf_b200.run_gap_factor, (True, False): pcmf_b200.run_zi_gap_factor,
       (False, True): pcmf_b200.run_sparse_gap_factor, (True, True): pcmf_b200.run_zi_sparse_gap_factor}


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
@pytest.mark.parametrize("K", [3, 20, 30])
def test_monitors_match_oracle(zi, sparse, K):
    """loglikelihood / deviance / exp_deviance of the state after 0, 1 and 4 iterations."""
    X = problem(21, 150, 90, K_true=5)
    m, s = make_pair(X, K, zi, sparse)
    done = 0
    for upto in (0, 1, 4):
        for _ in range(upto - done):
            m.update_param()
            m.prepare_next_iterate()
        s.iterate(upto - done)
        done = upto
        g = s.monitors()
        for name, ref in (("loglikelihood", m.loglikelihood()), ("deviance", m.deviance()), ("exp_dev", m.exp_deviance())):
            assert abs(g[name] - ref) < 1e-8 * max(1.0, abs(ref)), (upto, name, g[name], ref)
    s.close()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_order_factor_matches_oracle(zi, sparse):
    """Same greedy order, and every K-indexed array permuted like reorder_factor does; the solver stays consistent
    (one more iteration after the reorder still matches)."""
    X = problem(22, 120, 70, K_true=5)
    K = 5
    m, s = make_pair(X, K, zi, sparse)
    for _ in range(5):
        m.update_param()
        m.prepare_next_iterate()
    s.iterate(5)
    m.order_factor()
    s.order_factor()
    o = s.get_output()
    assert list(o["factor_order"]) == list(m.get("factor_order"))
    assert sorted(o["factor_order"]) == list(range(K))
    compare_state(m, s, zi, sparse, "after reorder")
    ref_dev = m.exp_deviance()
    assert abs(s.monitors(loglikelihood=False)["exp_dev"] - ref_dev) < 1e-8 * max(1.0, abs(ref_dev))
    m.update_param()
    m.prepare_next_iterate()
    s.iterate(1)
    compare_state(m, s, zi, sparse, "iteration after reorder")
    s.close()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_run_with_monitor_and_reorder(zi, sparse):
    """The full default call shape (monitor=TRUE, reorder_factor=TRUE): per-iteration log-likelihood and deviance
    vectors, exp_dev and the reordered factors."""
    X = problem(23, 90, 50)
    K = 4
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=9)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=25, iter_min=5, epsilon=1e-2, additional_iter=3, a1=a1, a2=a2, b1=b1, b2=b2)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=True, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, reorder_factor=True, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    for k in ("loglikelihood", "deviance", "optim_criterion", "norm_gap"):
        assert relerr(res["monitor"][k], ref["monitor"][k]) < 1e-6, k
    assert abs(res["exp_dev"] - ref["exp_dev"]) < 1e-8
    assert list(res["factor_order"]) == list(ref["factor_order"])
    for grp in ("variational_params", "hyper_params", "factor"):
        for k in res[grp]:
            assert relerr(res[grp][k], ref[grp][k]) < 1e-6, (grp, k)
    if sparse:
        assert (res["sparse_param"]["S"] == ref["sparse_param"]["S"]).all()


@pytest.mark.parametrize("zi,sparse", [(False, False), (True, True)], ids=["gap", "zi_sparse"])
def test_multi_init_matches_oracle(zi, sparse):
    """ninit = 3 seeded random initialisations of iter_init = 6 iterations, smallest loss kept (sic), then optimize
    continues from iteration iter_init: same kept run, same trajectory, same per-iteration vectors."""
    X = problem(24, 70, 40)
    K = 3
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=20, iter_min=8, epsilon=1e-2, additional_iter=3, seed=77, ninit=3, iter_init=6)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert relerr(res["convergence"]["conv_crit"], ref["convergence"]["conv_crit"]) < 1e-6
    assert relerr(res["monitor"]["optim_criterion"], ref["monitor"]["optim_criterion"]) < 1e-6
    assert abs(res["loss"] - ref["loss"]) < 1e-6 * abs(ref["loss"])
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k


@pytest.mark.parametrize("sparse", [False, True], ids=["zi", "zi_sparse"])
def test_warm_start_from_dense_prob_D_matches_oracle(sparse):
    """run_zi_*_gap_factor(..., prob_D = <dense n x p>, prior_D = <p>) as the reference's R tests call it
    (test-gap_factor_model.R:303-317): init_zi_param(prob_D, prior_D), first iteration on the explicit matrix, closed
    form from then on; freq_D keeps its constructor value (ones)."""
    X = problem(25, 64, 48)
    K = 4
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=12)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    rng = np.random.default_rng(3)
    PD, prD = rng.uniform(0.05, 0.95, X.shape), np.full(X.shape[1], 0.5)
    kw = dict(iter_max=15, iter_min=4, epsilon=1e-2, additional_iter=2, a1=a1, a2=a2, b1=b1, b2=b2, prob_D=PD, prior_D=prD)
    ref = po.run(X, K, True, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(True, sparse)](X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert relerr(res["convergence"]["conv_crit"], ref["convergence"]["conv_crit"]) < 1e-6
    for k in ("optim_criterion", "loglikelihood", "deviance"):
        assert relerr(res["monitor"][k], ref["monitor"][k]) < 1e-6, k
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k
    assert np.abs(res["ZI_param"]["prob_D"] - ref["ZI_param"]["prob_D"]).max() < 1e-8
    assert (res["ZI_param"]["freq_D"] == 1.0).all() and (ref["ZI_param"]["freq_D"] == 1.0).all()
    # state right after the initialisation: the explicit matrix is what get_output returns, ELBO matches
    m = po.Model(X, K, True, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    m.init_zi_param(PD, prD)
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X, K, True, sparse, a1=a1, a2=a2, b1=b1, b2=b2, prob_D=PD, prior_D=prD,
                         **(dict(prob_S=prob_S, prior_S=prior_S) if sparse else {}))
    assert abs(s.elbo() - m.elbo()) < 1e-8 * abs(m.elbo())
    o = s.get_output()
    assert np.abs(o["prob_D"] - PD).max() == 0.0
    assert relerr(o["prior_prob_D"], m.get("prior_D")) < 1e-12
    # monitors on the explicit matrix, before any iteration (zi_gap_factor_model.cpp:160-167, :218-245)
    mon = s.monitors()
    assert abs(mon["loglikelihood"] - m.loglikelihood()) < 1e-8 * abs(m.loglikelihood())
    assert abs(mon["deviance"] - m.deviance()) < 1e-8 * abs(m.deviance())
    assert abs(mon["exp_dev"] - m.exp_deviance()) < 1e-8 * max(1.0, abs(m.exp_deviance()))
    s.close()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_rv_coefficient_convergence_mode(zi, sparse):
    """conv_mode = 2: 1 - min(RV(EU, EU_old), RV(EV, EV_old)) (gap_factor_model.cpp:265-269, utils/matrix.h:100-104) as the
    stopping criterion; the gaps are still monitored (algorithm.h:432-437)."""
    X = problem(26, 70, 44)
    K = 4
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=13)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=30, iter_min=4, epsilon=2e-3, additional_iter=2, conv_mode=2, a1=a1, a2=a2, b1=b1, b2=b2)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert res["convergence"]["converged"] == ref["convergence"]["converged"]
    assert np.abs(res["convergence"]["conv_crit"] - ref["convergence"]["conv_crit"]).max() < 1e-10   # 1 - RV: absolute
    assert relerr(res["monitor"]["norm_gap"], ref["monitor"]["norm_gap"]) < 1e-6
    assert res["convergence"]["conv_mode"] == 2


@pytest.mark.parametrize("zi,sparse", [(False, False), (True, True)], ids=["gap", "zi_sparse"])
def test_init_from_factor_matches_oracle(zi, sparse):
    """run_*(X, K, U = ..., V = ...): init_from_factor (gap_factor_model.cpp:135-144)."""
    X = problem(27, 60, 40)
    K = 3
    rng = np.random.default_rng(8)
    U, V = rng.exponential(1.0, (60, K)) + 0.05, rng.exponential(1.0, (40, K)) + 0.05
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=8, iter_min=8, U=U, V=V)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k
    assert relerr(res["monitor"]["optim_criterion"], ref["monitor"]["optim_criterion"]) < 1e-6


@pytest.mark.parametrize("zi,sparse", [(False, False), (True, True)], ids=["gap", "zi_sparse"])
def test_multi_init_from_factors_perturbs_like_oracle(zi, sparse):
    """ninit = 3 with user-supplied U, V: every run starts from init_from_factor + perturb_param (uniform noise of
    max(U, V) / K off the one seeded stream, wrapper_gap_factor.h:312-314; pinned against the reference build in
    tests/test_ref_build.py), smallest loss kept, then optimize continues."""
    X = problem(28, 70, 40)
    K = 3
    rng = np.random.default_rng(11)
    U, V = rng.exponential(1.0, (70, K)) + 0.05, rng.exponential(1.0, (40, K)) + 0.05
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=16, iter_min=8, epsilon=1e-2, additional_iter=3, seed=91, ninit=3, iter_init=5, U=U, V=V)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert relerr(res["monitor"]["optim_criterion"], ref["monitor"]["optim_criterion"]) < 1e-6
    assert abs(res["loss"] - ref["loss"]) < 1e-6 * abs(ref["loss"])
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k
    assert np.abs(res["variational_params"]["a1"] - U).max() > 1e-3


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
@pytest.mark.parametrize("form", ["vectors", "same_rows", "rows_differ"])
def test_hyper_only_warm_start_matches_oracle(zi, sparse, form):
    """run_*(X, K, alpha1 =, alpha2 =, beta1 =, beta2 =) and nothing else: init_hyper_param alone (wrapper_gap_factor.h:362-363,
    gap_factor_model.cpp:126-132) -- Ones everywhere else, no draw, no hyper update; matrices whose rows differ are taken as
    given and stay row-dependent through update_prior_gamma_param (gap...cpp:541-548).  Pinned against the reference build
    in tests/test_ref_build.py."""
    X = problem(27, 70, 52)
    n, p, K = 70, 52, 4
    rng = np.random.default_rng(10)
    if form == "rows_differ":
        hy = dict(alpha1=rng.uniform(0.5, 2.0, (n, K)), alpha2=rng.uniform(0.5, 2.0, (n, K)),
                  beta1=rng.uniform(0.5, 2.0, (p, K)), beta2=rng.uniform(0.5, 2.0, (p, K)))
        gpu_hy = hy
    else:
        vec = {k: rng.uniform(0.5, 2.0, K) for k in ("alpha1", "alpha2", "beta1", "beta2")}
        hy = {k: np.tile(v, (n if k.startswith("alpha") else p, 1)) for k, v in vec.items()}
        gpu_hy = vec if form == "vectors" else hy
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=12, iter_min=4, epsilon=1e-2, additional_iter=2)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=True, prob_S=prob_S, prior_S=prior_S, **kw, **hy)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, reorder_factor=True, **kw, **gpu_hy)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert relerr(res["convergence"]["conv_crit"], ref["convergence"]["conv_crit"]) < 1e-6
    for k in ("optim_criterion", "loglikelihood", "deviance"):
        assert relerr(res["monitor"][k], ref["monitor"][k]) < 1e-6, k
    assert list(res["factor_order"]) == list(ref["factor_order"])
    for grp in ("variational_params", "hyper_params"):
        for k in res[grp]:
            assert relerr(res[grp][k], ref[grp][k]) < 1e-6, (grp, k)
    assert abs(res["loss"] - ref["loss"]) < 1e-6 * abs(ref["loss"])


@pytest.mark.parametrize("zi,sparse", [(False, False), (True, True)], ids=["gap", "zi_sparse"])
def test_hyper_parameters_are_ignored_next_to_variational_parameters(zi, sparse):
    """The reference's else-if chain never reaches init_all_param (wrapper_gap_factor.h:364-367 is dead code): with a1..b2
    (or U, V) given, alpha / beta arguments change nothing."""
    X = problem(28, 50, 40)
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=13)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(verbose=False, iter_max=6, iter_min=6, reorder_factor=False, a1=a1, a2=a2, b1=b1, b2=b2)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    plain = RUN[(zi, sparse)](X, K, **kw)
    hy = dict(alpha1=np.full(K, 3.0), alpha2=np.full(K, 0.5), beta1=np.full(K, 2.0), beta2=np.full(K, 4.0))
    with_hy = RUN[(zi, sparse)](X, K, **kw, **hy)
    for k in ("a1", "a2", "b1", "b2"):     # (atomic accumulation order: equal up to rounding, not bit for bit)
        assert relerr(plain["variational_params"][k], with_hy["variational_params"][k]) < 1e-12, k


@pytest.mark.parametrize("sparse", [False, True], ids=["zi", "zi_sparse"])
def test_loglikelihood_fused_into_the_cell_pass(sparse):
    """monitor = TRUE: the zero part of zi_poisson_loglike (likelihood.h:197-212, zi_gap_factor_model.cpp:160-167) comes out of
    the cell-side zero-inflation pass of the iteration (k_zi_cells_mma<.., WITH_LL>) instead of a second dense pass.  The
    per-iteration vector of run_* (fused) must equal the oracle's, and the last entry the stand-alone monitor pass."""
    X = problem(31, 333, 217, prob1=0.4)
    X[:, 5] = 0; X[0, 5] = 3                      # a gene that is almost never expressed
    X[:, 9] = np.maximum(X[:, 9], 1)              # a gene without zeros: prior_D == 1 shortcut (zi...cpp:353-356)
    K = 20
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=3)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=6, iter_min=6, epsilon=1e-9, additional_iter=3, a1=a1, a2=a2, b1=b1, b2=b2)
    ref = po.run(X, K, True, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = RUN[(True, sparse)](X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    assert relerr(res["monitor"]["loglikelihood"], ref["monitor"]["loglikelihood"]) < 1e-9
    # the same state, iterated step-wise: pcmf_monitors takes the stand-alone pass
    kw2 = {k: v for k, v in kw.items() if k in ("a1", "a2", "b1", "b2", "prob_S", "prior_S")}
    s = pcmf_b200.Solver(X, K, True, sparse, **kw2)
    s.iterate(6)
    alone = s.monitors()["loglikelihood"]
    s.close()
    assert abs(alone - res["monitor"]["loglikelihood"][-1]) < 1e-11 * abs(alone)


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
@pytest.mark.parametrize("K", [4, 20, 30, 64])
def test_order_factor_fp32_pass_with_bounds_gives_the_exact_order(zi, sparse, K):
    """The greedy ordering takes its K (K + 1) / 2 logarithms per non-zero in FP32 with an error bound per score and repeats a
    round in FP64 only when the winner does not clear every other candidate by the bounds: same order as the all-FP64 ordering
    (kernel_variant bit 16384) and as the oracle."""
    X = problem(31 + K, 400, 150, K_true=5)
    m, s = make_pair(X, K, zi, sparse)
    _, sx = make_pair(X, K, zi, sparse, kernel_variant=16384)
    for _ in range(4):
        m.update_param()
        m.prepare_next_iterate()
    s.iterate(4)
    sx.iterate(4)
    m.order_factor()
    s.order_factor()
    sx.order_factor()
    o, ox = s.get_output(), sx.get_output()
    assert list(o["factor_order"]) == list(ox["factor_order"]) == list(m.get("factor_order"))
    assert sx.phase_times()["order_factor_exact_rounds"][1] == 0          # (not counted when every round is exact anyway)
    compare_state(m, s, zi, sparse, "after reorder (fp32 pass)")
    s.close()
    sx.close()


def test_order_factor_tie_falls_back_to_the_exact_round():
    """Two factors with identical initial values stay identical, so their scores tie exactly: the FP32 pass cannot decide, the
    round is repeated in FP64 and the first of the two wins as in the reference (model.cpp:161-172)."""
    X = problem(5, 200, 90, K_true=3)
    K = 4
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=9)
    for A in (a1, a2, b1, b2):
        A[:, 3] = A[:, 1]
    from oracle import pyoracle as po
    m = po.Model(X, K, False, False)
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X, K, False, False, a1=a1, a2=a2, b1=b1, b2=b2)
    for _ in range(3):
        m.update_param()
        m.prepare_next_iterate()
    s.iterate(3)
    m.order_factor()
    s.order_factor()
    # (the two tied factors are interchangeable: which of them comes first is decided by the summation order of the reductions,
    #  in the OpenMP oracle as well -- the library itself takes the first, like a sequential run of the reference)
    same = lambda order: [1 if int(k) == 3 else int(k) for k in order]
    got = list(s.get_output()["factor_order"])
    assert same(got) == same(m.get("factor_order"))
    assert got.index(1) < got.index(3)
    assert s.phase_times()["order_factor_exact_rounds"][1] >= 1
    s.close()
