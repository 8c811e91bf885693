"""GPU parity tests proper: the CUDA path, called through the C ABI (pcmf_b200.api -> libpcmf_b200.so), against the
CPU oracle on the same seeded inputs and the same explicit initialisation.

Tolerances (north_star: 1e-6 relative in FP64 mode): variational parameters, statistics and ELBO are compared at
RTOL = 1e-8 (observed agreement is ~1e-12; the slack covers summation order and exp(a)exp(b) vs exp(a+b)).
"""
import numpy as np
import pytest

import pcmf_b200
from oracle import pyoracle as po
from pcmf_b200 import datagen

pytestmark = pytest.mark.gpu

RTOL = 1e-8
MODELS = [(False, False), (True, False), (False, True), (True, True)]
IDS = ["gap", "zi", "sparse", "zi_sparse"]


def problem(seed, n, p, K_true=4, prob1=0.6, signal=20.0):
    X, _, _ = datagen.simulate(n, p, K_true, signal_U=signal, signal_V=signal, prob1=prob1, seed=seed)
    return X


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))) if a.size else 0.0


def make_pair(X, K, zi, sparse, seed=11, sel_bound=0.5, **kw):
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=seed)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    m = po.Model(X, K, zi, sparse)
    if sparse:
        m.init_sparse_param(sel_bound, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X, K, zi, sparse, sel_bound=sel_bound, a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S, **kw)
    return m, s


def compare_state(m, s, zi, sparse, tag):
    o = s.get_output()
    for name in ("a1", "a2", "b1", "b2", "EU", "EV", "alpha1", "alpha2", "beta1", "beta2"):
        assert relerr(o[name], m.get(name)) < RTOL, (tag, name, relerr(o[name], m.get(name)))
    for name in ("ElogU", "ElogV"):   # signed quantities crossing zero: absolute, scaled by magnitude
        assert np.max(np.abs(o[name] - m.get(name))) < RTOL * max(1.0, np.abs(m.get(name)).max()), (tag, name)
    if sparse:
        assert np.max(np.abs(o["prob_S"] - m.get("prob_S"))) < RTOL, tag
        assert (o["S"] == m.get("S")).all(), tag
        assert relerr(o["prior_prob_S"], m.get("prior_S")) < RTOL, tag
    if zi:
        assert np.max(np.abs(o["prob_D"] - m.get("prob_D"))) < RTOL, tag
        assert relerr(o["prior_prob_D"], m.get("prior_D")) < RTOL, tag
        assert relerr(o["freq_D"], m.get("freq_D")) < 1e-14, tag


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
# K = 3, 20, 5, 10, 14, 30, 64 -> padded widths 4, 20, 8, 12, 16, 32, 64: every compiled kernel set (BASELINE config 5 sweeps K)
@pytest.mark.parametrize("n,p,K", [(60, 45, 3), (97, 130, 20), (40, 33, 5), (70, 90, 10), (66, 70, 14), (75, 140, 30),
                                   (131, 100, 64)])
def test_free_running_iterations_match_oracle(zi, sparse, n, p, K):
    X = problem(3, n, p)
    m, s = make_pair(X, K, zi, sparse)
    compare_state(m, s, zi, sparse, "init")
    e_gpu, terms = s.elbo(terms=True)
    e_cpu = m.elbo()
    assert abs(e_gpu - e_cpu) < RTOL * abs(e_cpu), ("init elbo", e_gpu, e_cpu, terms, m.elbo_terms())
    for it in range(6):
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
def test_run_matches_oracle_end_to_end(zi, sparse):
    """run_* through the whole driver (optimize + stopping rule + monitor vectors + loss) vs the oracle's driver."""
    X = problem(5, 80, 60)
    K = 4
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=7)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    kw = dict(iter_max=60, iter_min=10, epsilon=1e-2, additional_iter=3, a1=a1, a2=a2, b1=b1, b2=b2)
    ref = po.run(X, K, zi, sparse, monitor=True, reorder_factor=False, prob_S=prob_S, prior_S=prior_S, **kw)
    fn = {(False, False): pcmf_b200.run_gap_factor, (True, False): pcmf_b200.run_zi_gap_factor,
          (False, True): pcmf_b200.run_sparse_gap_factor, (True, True): pcmf_b200.run_zi_sparse_gap_factor}[(zi, sparse)]
    if sparse:
        kw.update(prob_S=prob_S, prior_S=prior_S)
    res = fn(X, K, verbose=False, monitor=True, reorder_factor=False, **kw)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    assert res["convergence"]["converged"] == ref["convergence"]["converged"]
    assert relerr(res["convergence"]["conv_crit"], ref["convergence"]["conv_crit"]) < 1e-6
    assert relerr(res["monitor"]["optim_criterion"], ref["monitor"]["optim_criterion"]) < 1e-6   # ELBO trajectory
    assert relerr(res["monitor"]["abs_gap"], ref["monitor"]["abs_gap"]) < 1e-6
    assert abs(res["loss"] - ref["loss"]) < 1e-6 * abs(ref["loss"])
    for grp in ("variational_params", "hyper_params", "factor"):
        for k in res[grp]:
            assert relerr(res[grp][k], ref[grp][k]) < 1e-6, (grp, k)
    if sparse:
        assert (res["sparse_param"]["S"] == ref["sparse_param"]["S"]).all()
        assert np.max(np.abs(res["sparse_param"]["prob_S"] - ref["sparse_param"]["prob_S"])) < 1e-6
    if zi:
        assert np.max(np.abs(res["ZI_param"]["prob_D"] - ref["ZI_param"]["prob_D"])) < 1e-6
    assert res.cls == ("spCMF" if sparse else "pCMF")


def test_pcmf_entry_point_readme_shape():
    """BASELINE config 1 (README example shape: n=100, p=500, ZI 40%, K=2, ZI + sparse) through pCMF() itself, seeded
    random init on both sides (same MT19937 + inversion stream in the oracle and the library)."""
    X, _, _ = datagen.simulate(100, 500, 20, ngroup_U=2, ngroup_V=2, signal_U=120.0, signal_V=120.0, prob1=0.4, seed=250)
    K = 2
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    ref = po.run(X, K, True, True, iter_max=40, iter_min=20, seed=123, reorder_factor=False, prob_S=prob_S, prior_S=prior_S)
    res = pcmf_b200.pCMF(X, K, verbose=False, iter_max=40, iter_min=20, seed=123, reorder_factor=False)
    assert res["convergence"]["nb_iter"] == ref["convergence"]["nb_iter"]
    for k in ("a1", "a2", "b1", "b2"):
        assert relerr(res["variational_params"][k], ref["variational_params"][k]) < 1e-6, k
    assert relerr(res["monitor"]["optim_criterion"], ref["monitor"]["optim_criterion"]) < 1e-6
    assert res["seed"] == 123 and res.cls == "spCMF"
    # return-object contract (test-gap_factor_model.R:35-75): matrices, integer nb_iter, logical converged
    assert res["factor"]["U"].shape == (100, K) and res["factor"]["V"].shape == (500, K)
    assert isinstance(res["convergence"]["nb_iter"], int) and isinstance(res["convergence"]["converged"], bool)
    assert res["ZI_param"]["prob_D"].shape == (100, 500) and res["sparse_param"]["S"].dtype == np.int32
    assert pcmf_b200.getU(res).shape == (100, K) and pcmf_b200.getV(res).shape == (500, K)


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_guard_path_matches_oracle(zi, sparse):
    """The +-100 guards of the multinomial update (gap...cpp:483, internal.h:121-127): shapes that make
    E[log U] huge / tiny force the exact (non-hoisted) kernel path; results must still match the dense oracle."""
    rng = np.random.default_rng(0)
    n, p, K = 48, 40, 3
    X = problem(9, n, p)
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=2)
    a1[0, 0] = 1e60
    a1[1, :] = 1e-3
    a1[2, 1] = 5e-3
    b1[3, 2] = 2e-3
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    m = po.Model(X, K, zi, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X, K, zi, sparse, a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S)
    for it in range(3):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, zi, sparse, "guard iter %d" % it)
        e_gpu, e_cpu = s.elbo(), m.elbo()
        assert abs(e_gpu - e_cpu) < RTOL * abs(e_cpu), (it, e_gpu, e_cpu)
        m.prepare_next_iterate()
    s.close()


def test_input_formats_agree():
    """INTSXP / REALSXP dense (wrapper_gap_factor.h:204-211), scipy CSR, host CSR tuple and device CSR give the same fit."""
    import scipy.sparse as sp
    X = problem(13, 70, 50)
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=5)
    kw = dict(a1=a1, a2=a2, b1=b1, b2=b2)
    outs = []
    Xc = sp.csr_matrix(X.astype(np.float64))
    variants = [X.astype(np.int32), X.astype(np.float64), Xc,
                (Xc.indptr.astype(np.int64), Xc.indices.astype(np.int32), Xc.data.astype(np.float32), X.shape[1])]
    for Xv in variants:
        s = pcmf_b200.Solver(Xv, K, True, False, **kw)
        s.iterate(4)
        outs.append(s.get_output())
        s.close()
    for o in outs[1:]:
        for k in ("a1", "a2", "b1", "b2", "prob_D"):
            assert np.allclose(o[k], outs[0][k], rtol=1e-9, atol=1e-12), k   # atomics reorder sums; 4 iterations amplify
    # non-integer "counts" (the reference does not check, wrapper_gap_factor.h:204-211) take the FP64 value path
    Xr = X.astype(np.float64) * 1.1
    m = po.Model(Xr, K, False, False)
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(Xr, K, False, False, **kw)
    for _ in range(3):
        m.update_param()
        m.prepare_next_iterate()
    s.iterate(3)
    compare_state(m, s, False, False, "real-valued X")
    s.close()


def test_device_generator_and_conservation_at_scale():
    """Size-independent properties at a size the dense oracle cannot hold (n=200k, p=2k, ~5% nnz):
    the multinomial split conserves counts -- sum_k (a1 - alpha1)[i,k] = sum_j X[i,j] and
    sum_k (b1 - beta1)[j,k] = sum_i X[i,j] for the plain GaP model (gap...cpp:489-493 with sum_k r_ijk = 1)."""
    n, p, Kt, K = 200_000, 2_000, 8, 8
    rng = np.random.default_rng(1)
    U = rng.exponential(0.12, (n, Kt))
    V = rng.exponential(0.12, (p, Kt))
    csr = pcmf_b200.generate_counts_device(n, p, Kt, U, V, prob1=0.4, seed=3)
    dens = csr.nnz / (n * p)
    assert 0.01 < dens < 0.2
    rp, ci, vv = csr.to_host()
    assert rp[0] == 0 and rp[-1] == csr.nnz and (np.diff(rp) >= 0).all()
    assert ci.min() >= 0 and ci.max() < p and (vv >= 1).all() and (vv == np.round(vv)).all()
    # expected density of the generative model: P(keep) * P(Poisson(lambda) > 0), checked on a row sample
    lam = U[:2000] @ V.T
    exp_d = 0.4 * (1 - np.exp(-lam)).mean()
    got_d = (rp[2000] - rp[0]) / (2000 * p)
    assert abs(got_d - exp_d) < 0.05 * exp_d
    rowsum = np.add.reduceat(vv.astype(np.float64), rp[:-1])
    rowsum[np.diff(rp) == 0] = 0
    colsum = np.bincount(ci, weights=vv.astype(np.float64), minlength=p)
    live = rowsum > 0
    a1 = np.where(live[:, None], rng.exponential(1.0, (n, K)), 1.0) + 0.1
    b1 = rng.exponential(1.0, (p, K)) + 0.1
    s = pcmf_b200.Solver(csr, K, False, False, a1=a1, a2=np.ones((n, K)), b1=b1, b2=np.ones((p, K)))
    o0 = s.get_output()
    s.iterate(1)
    o = s.get_output()
    # alpha used by the update is the one computed at init (o0)
    ez_j = (o["a1"] - o0["alpha1"]).sum(1)
    ez_i = (o["b1"] - o0["beta1"]).sum(1)
    assert np.max(np.abs(ez_j - rowsum) / np.maximum(rowsum, 1)) < 1e-10
    assert np.max(np.abs(ez_i - colsum) / np.maximum(colsum, 1)) < 1e-10
    assert np.isfinite(s.elbo())
    s.close()
    csr.free()


def test_full_size_c4_properties():
    """BASELINE config 4 at FULL size (1M cells x 20k genes, ~1e9 non-zeros, K = 20, ZI + sparse V) -- far beyond what the
    dense oracle can hold -- through size-independent properties of two iterations:
      * the cell-side pass (stored q, CSR order) and the gene-side pass (CSC order) split the SAME weighted counts:
        sum_ik (a1 - alpha1) = sum_jk (b1 - beta1) = sum_ij x_ij sum_k r_ijk prob_S_jk prob_D_ij  (zi_sparse...cpp:340-352);
      * prior_D_j = colmean(prob_D) lies in [freq_D_j, 1] (prob_D = 1 on non-zeros, in [0,1] elsewhere; zi...cpp:359-372);
      * prior_S = rowmean(prob_S) in [0,1], S = (prob_S >= sel_bound) exactly (sparse...cpp:470-476);
      * the normalised gap decreases and the ELBO is finite; few entries need the guarded formula."""
    import bench
    cfg = bench.CONFIGS["c4"]
    n, p, K = cfg["n"], cfg["p"], cfg["K"]
    U, V = bench.factor_matrices(cfg, 3)
    csr = pcmf_b200.generate_counts_device(n, p, cfg["K_true"], U, V, prob1=cfg["prob1"], seed=3)
    assert 0.045 < csr.nnz / (n * p) < 0.055
    s = pcmf_b200.Solver(csr, K, True, True, seed=1003, prob_S=np.full((p, K), 0.5), prior_S=np.full(p, 0.5), iter_max=10)
    o0 = s.get_output(want_prob_D=False)
    crit1 = s.iterate(1)[0]
    o1 = s.get_output(want_prob_D=False)
    ez_cells = float((o1["a1"] - o0["alpha1"]).sum())
    ez_genes = float((o1["b1"] - o0["beta1"]).sum())
    assert abs(ez_cells - ez_genes) < 1e-9 * abs(ez_genes), (ez_cells, ez_genes)
    crit2 = s.iterate(1)[0]
    o2 = s.get_output(want_prob_D=False)
    assert crit2 < crit1
    ez_cells = float((o2["a1"] - o1["alpha1"]).sum())
    ez_genes = float((o2["b1"] - o1["beta1"]).sum())
    assert abs(ez_cells - ez_genes) < 1e-9 * abs(ez_genes), (ez_cells, ez_genes)
    assert (o2["prior_prob_D"] >= o2["freq_D"] - 1e-12).all() and (o2["prior_prob_D"] <= 1 + 1e-12).all()
    assert (o2["prior_prob_S"] >= 0).all() and (o2["prior_prob_S"] <= 1).all()
    assert ((o2["prob_S"] >= 0.5) == (o2["S"] == 1)).all()
    assert np.isfinite(s.elbo())
    # right after a random initialisation some cells have exp(E[log U]) underflowing (tiny Gamma shapes): those entries
    # take the reference's guarded formula in both passes; they must stay a small minority
    ph = s.phase_times()
    assert ph["exact_fallback_entries_cells"][1] < 0.02 * 2 * csr.nnz
    s.close()
    csr.free()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
@pytest.mark.parametrize("n,p,K", [(9, 5, 2), (33, 7, 1), (5, 40, 3)])
def test_degenerate_shapes_match_oracle(zi, sparse, n, p, K):
    """Shapes smaller than every tile (fewer genes than one 8-gene group, fewer cells than one warp, K = 1)."""
    rng = np.random.default_rng(100 + n + p)
    X = rng.poisson(2.0, (n, p)).astype(float) * (rng.uniform(size=(n, p)) < 0.7)
    X[X.sum(1) == 0, 0] = 1
    X[0, X.sum(0) == 0] = 1
    m, s = make_pair(X, K, zi, sparse)
    for it in range(4):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, zi, sparse, "iter %d" % it)
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()


def _fuzz_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for c in range(36):
        n = int(rng.choice([1, 2, 7, 8, 9, 31, 32, 33, 63, 65, 127, 129, 200, 257, 300]))
        p = int(rng.choice([1, 2, 7, 8, 9, 31, 33, 63, 64, 65, 95, 129, 190]))
        K = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 12, 13, 16, 17, 20, 21, 32, 33, 48, 64]))
        zi, sparse = MODELS[c % 4]
        if sparse and n == 1:
            n = 2                                   # the prior_S heuristic needs a standard deviation (pCMF.R:228)
        if c % 3 == 0:                              # a third of the cases keep the drawn K: widen the matrix instead
            n, p = max(n, K), max(p, K)
        K = min(K, n, p)                            # K <= min(n, p) is an argument check (wrapper_gap_factor.h:210-213)
        cases.append((c, n, p, K, zi, sparse, float(rng.choice([0.15, 0.5, 0.9]))))
    return cases


@pytest.mark.parametrize("c,n,p,K,zi,sparse,dens", _fuzz_cases(), ids=lambda v: str(v))
def test_random_shapes_match_oracle(c, n, p, K, zi, sparse, dens):
    """Seeded sweep over shapes around every tile boundary of the kernels (8-row fragments, 32-gene occupancy words,
    64-wide tiles, 128-cell CTAs, each padded factor width) and over densities: three free-running iterations, every
    state array and the ELBO against the oracle."""
    rng = np.random.default_rng(1000 + c)
    X = rng.poisson(3.0, (n, p)).astype(float) * (rng.uniform(size=(n, p)) < dens)
    X[X.sum(1) == 0, rng.integers(p)] = 1
    X[rng.integers(n), X.sum(0) == 0] = 1
    m, s = make_pair(X, K, zi, sparse, seed=c)
    for it in range(3):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, zi, sparse, "case %d iter %d" % (c, it))
        assert abs(s.elbo() - m.elbo()) < RTOL * max(1.0, abs(m.elbo()))
        m.prepare_next_iterate()
    s.close()


@pytest.mark.parametrize("sparse", [False, True], ids=["zi", "zi_sparse"])
@pytest.mark.parametrize("n,p,K", [(300, 190, 3), (65, 33, 20), (129, 95, 12), (2500, 40, 4), (70, 130, 64), (9, 5, 2), (200, 1100, 5)])
def test_stored_prob_D_path_matches_oracle(sparse, n, p, K):
    """The gene-side ZI pass on the prob_D that the previous cell pass stored (32-bit fixed point, 8 x 8 tiles) instead of
    recomputing it -- the default from 5e7 entries on, forced here with kernel_variant bit 64.  Same oracle comparison as
    the recomputing path at the same tolerance (stored values carry an absolute error of 1.2e-10), and the two paths agree
    with each other to 1e-9; (2500, 40) splits the cells of a gene block over several CTAs, (200, 1100) the genes of a cell
    block."""
    rng = np.random.default_rng(7 + n + p)
    X = rng.poisson(2.5, (n, p)).astype(float) * (rng.uniform(size=(n, p)) < 0.45)
    X[X.sum(1) == 0, rng.integers(p)] = 1
    X[rng.integers(n), X.sum(0) == 0] = 1
    m, s = make_pair(X, K, True, sparse, kernel_variant=64)
    _, s0 = make_pair(X, K, True, sparse, kernel_variant=128)
    for it in range(5):
        m.update_param()
        s.iterate(1)
        s0.iterate(1)
        compare_state(m, s, True, sparse, "stored prob_D, iter %d" % it)
        o, o0 = s.get_output(), s0.get_output()
        for k in ("a1", "a2", "b1", "b2"):
            assert relerr(o[k], o0[k]) < 1e-9, (it, k)
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()
    s0.close()


@pytest.mark.parametrize("zi,sparse", MODELS, ids=IDS)
def test_cell_blocked_gene_passes_match_oracle(zi, sparse):
    """The gene passes walk the cells in blocks (so that the slice of the cell table they gather from stays in L2) and add the
    per-block sums; forced to 64-cell blocks here (kernel_variant bit 1024; one block per ~1e5 cells by default)."""
    X = problem(61, 300, 190, K_true=5)
    m, s = make_pair(X, 6, zi, sparse, kernel_variant=1024)
    for it in range(5):
        m.update_param()
        s.iterate(1)
        compare_state(m, s, zi, sparse, "blocked gene pass, iter %d" % it)
        assert abs(s.elbo() - m.elbo()) < RTOL * abs(m.elbo())
        m.prepare_next_iterate()
    s.close()
