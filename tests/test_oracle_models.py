"""Cross-checks the dense C oracle (loop-for-loop restatement of the reference) against the numpy design twin
(sparse / closed-form identities) over several free-running iterations of all four models."""
import numpy as np
import pytest

from oracle import pyoracle as po
from pcmf_b200 import datagen
from tests.np_twin import Twin

MODELS = [(False, False), (True, False), (False, True), (True, True)]


def small_problem(seed=0, n=40, p=30, K_true=4, prob1=0.6):
    X, U, V = datagen.simulate(n, p, K_true, signal_U=20.0, signal_V=20.0, prob1=prob1, seed=seed)
    return X.astype(float)


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("zi,sparse", MODELS)
def test_oracle_matches_design_twin(zi, sparse):
    X = small_problem(seed=3)
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=11)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    m = po.Model(X, K, zi, sparse)
    t = Twin(X, K, zi, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    t.init(a1, a2, b1, b2, prob_S, prior_S)
    for it in range(6):
        m.update_param()
        t.iterate()
        for name in ("a1", "a2", "b1", "b2", "EU", "EV", "ElogU", "ElogV"):
            assert rel(getattr(t, name), m.get(name)) < 1e-9, (it, name)
        assert rel(t.alpha1, m.get("alpha1")[0]) < 1e-9
        assert rel(t.alpha2, m.get("alpha2")[0]) < 1e-9
        assert rel(t.beta1, m.get("beta1")[0]) < 1e-9
        assert rel(t.beta2, m.get("beta2")[0]) < 1e-9
        # hyper-parameter rows stay identical
        assert np.ptp(m.get("alpha1"), axis=0).max() == 0 and np.ptp(m.get("beta2"), axis=0).max() == 0
        if sparse:
            assert np.max(np.abs(t.prob_S - m.get("prob_S"))) < 1e-9
            assert (t.S.astype(int) == m.get("S")).all()
            assert rel(t.prior_S, m.get("prior_S")) < 1e-9
        if zi:
            assert np.max(np.abs(t.P_formula() - m.get("prob_D"))) < 1e-10
            assert rel(t.prior_D, m.get("prior_D")) < 1e-9
        ag, ng = m.gap_between_iterates()
        tg = t.gap()
        assert abs(tg[0] - ag) < 1e-9 * ag and abs(tg[1] - ng) < 1e-9 * ng
        e_o, e_t = m.elbo(), t.elbo()
        assert abs(e_o - e_t) < 1e-9 * abs(e_o), (it, e_o, e_t, m.elbo_terms())
        m.prepare_next_iterate()
        t.next()


def test_first_gap_is_against_ones():
    # algorithm.h:158-210 never calls prepare_next_iterate before the first update: old = Ones (gap...cpp:72-75)
    X = small_problem(seed=5)
    K = 2
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=1)
    m = po.Model(X, K, False, False)
    m.init_variational_param(a1, a2, b1, b2)
    m.update_param()
    ag, ng = m.gap_between_iterates()
    d = sum(((m.get(k) - 1.0) ** 2).sum() for k in ("a1", "a2", "b1", "b2"))
    nrm = 2 * (X.shape[0] + X.shape[1]) * K
    assert abs(ag - np.sqrt(d)) < 1e-12 * ag and abs(ng - np.sqrt(d / nrm)) < 1e-12 * ng


def test_optimize_driver_convergence_rule():
    # algorithm.h:412-450: strict > on both counters; vectors truncated to nb_iter
    X = small_problem(seed=7)
    K = 2
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=2)
    m = po.Model(X, K, False, False)
    m.init_variational_param(a1, a2, b1, b2)
    out = m.optimize(iter_max=400, iter_min=5, epsilon=1e-2, additional_iter=3, monitor=True)
    nb = out["nb_iter"]
    assert out["converged"] and 5 < nb < 400
    crit = out["conv_crit"]
    assert len(crit) == nb and (np.abs(crit[-4:]) < 1e-2).all()
    # converged exactly when the stability counter first exceeded additional_iter with iter > iter_min
    stab, first = 0, None
    for i, c in enumerate(crit):
        stab = stab + 1 if abs(c) < 1e-2 else 0
        if stab > 3 and i > 5:
            first = i
            break
    assert first == nb - 1
    assert np.isfinite(out["optim_criterion"]).all() and np.isfinite(out["loglikelihood"]).all()
    assert abs(out["loss"] - out["optim_criterion"][-1]) < 1e-9 * abs(out["loss"])
    # ELBO is non-decreasing up to the hyper-parameter / numerical noise for the plain GaP model
    el = out["optim_criterion"]
    assert el[-1] > el[0]


@pytest.mark.parametrize("zi,sparse", MODELS)
def test_run_returns_reference_shaped_object(zi, sparse):
    # return-object contract: algorithm.h:355-377, gap...cpp:369-393, sparse...cpp:278-286, zi...cpp:248-255
    X = small_problem(seed=9, n=30, p=20)
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=3)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    res = po.run(X, K, zi, sparse, iter_max=20, iter_min=10, a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S)
    assert res["factor"]["U"].shape == (30, K) and res["factor"]["V"].shape == (20, K)
    assert set(res["variational_params"]) == {"a1", "a2", "b1", "b2"}
    assert set(res["hyper_params"]) == {"alpha1", "alpha2", "beta1", "beta2"}
    assert set(res["stats"]) == {"EU", "EV", "ElogU", "ElogV"}
    assert isinstance(res["convergence"]["nb_iter"], int) and isinstance(res["convergence"]["converged"], bool)
    assert ("sparse_param" in res) == sparse and ("ZI_param" in res) == zi
    assert sorted(res["factor_order"]) == list(range(K))
    assert np.isfinite(res["exp_dev"]) and res["exp_dev"] < 1
    # seeded random init path runs too
    res2 = po.run(X, K, zi, sparse, iter_max=5, iter_min=2, seed=42, prob_S=prob_S, prior_S=prior_S)
    assert np.isfinite(res2["loss"])


def test_guard_paths_match_dense_definition():
    """Drive the +-100 guards (gap...cpp:483, internal.h:121-127): huge and tiny shapes so that
    max_k s_k >= 100 on some pairs and s_k - max <= -100 on others; the E-step must still be finite
    and r_ijk must sum to one over k on every non-masked pair."""
    rng = np.random.default_rng(0)
    n, p, K = 12, 10, 3
    X = rng.poisson(2.0, (n, p)).astype(float) + 1
    a1 = rng.uniform(0.5, 2, (n, K))
    a1[0, 0] = 1e60          # ElogU ~ +138  -> shift branch
    a1[1, :] = 1e-3          # ElogU ~ -1000 -> 3e-44 floor on every k
    a1[2, 1] = 5e-3          # one tiny component
    b1 = rng.uniform(0.5, 2, (p, K))
    m = po.Model(X, K, False, False)
    m.init_all_param(np.ones((n, K)), np.ones((n, K)), np.ones((p, K)), np.ones((p, K)),
                     a1, np.ones((n, K)), b1, np.ones((p, K)))
    m.phase("estep")
    EZ_j, EZ_i, T = m.get("EZ_j"), m.get("EZ_i"), m.get("T")
    assert np.isfinite(EZ_j).all() and np.isfinite(EZ_i).all() and (T > 0).all()
    # rows of r sum to one -> sum_k EZ_j[i,k] = sum_j X[i,j]
    assert np.allclose(EZ_j.sum(1), X.sum(1), rtol=1e-12)
    assert np.allclose(EZ_i.sum(1), X.sum(0), rtol=1e-12)
    # row 1: every exponent floored -> uniform split 1/K
    assert np.allclose(EZ_j[1], X[1].sum() / K, rtol=1e-12)


@pytest.mark.parametrize("zi,sparse", [(False, False), (True, False), (False, True), (True, True)], ids=["gap", "zi", "sparse", "zi_sparse"])
def test_sparse_aware_cpu_variant_is_the_same_algorithm(zi, sparse):
    """bench.py's second CPU figure times the oracle with set_sparse_aware(): the same loops, except that an entry with
    X = 0 -- whose multinomial weights are multiplied by X = 0 everywhere they are used -- skips the K exponentials.  The
    trajectories must coincide (summation order of the spike-and-slab term aside)."""
    from pcmf_b200 import datagen
    X, _, _ = datagen.simulate(60, 45, 4, signal_U=15.0, signal_V=15.0, prob1=0.4, seed=31)
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=5)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    ms = []
    for aware in (False, True):
        m = po.Model(X, K, zi, sparse)
        m.set_sparse_aware(aware)
        if sparse:
            m.init_sparse_param(0.5, prob_S, prior_S)
        if zi:
            m.init_zi_param()
        m.init_variational_param(a1, a2, b1, b2)
        ms.append(m)
    for it in range(6):
        for m in ms:
            m.update_param()
        e0, e1 = ms[0].elbo(), ms[1].elbo()
        assert abs(e0 - e1) < 1e-11 * abs(e0), (it, e0, e1)
        for k in ("a1", "a2", "b1", "b2"):
            assert np.max(np.abs(ms[0].get(k) - ms[1].get(k)) / np.abs(ms[0].get(k))) < 1e-11, (it, k)
        if sparse:
            assert np.max(np.abs(ms[0].get("prob_S") - ms[1].get("prob_S"))) < 1e-11
        for m in ms:
            m.prepare_next_iterate()
