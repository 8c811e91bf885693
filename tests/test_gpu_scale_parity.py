"""Oracle parity at the sizes and through the code paths the benchmark runs ("next round" item 1).

Every case uses the DEFAULT heuristics (kernel_variant = 0): the stored prob_D switches on from 5e7 entries, the gene
passes walk the cells in L2-sized blocks from ~1e5 cells on, grids span several waves, 32-bit tile indices are exercised
at realistic sizes.  The oracle is the dense OpenMP restatement of the reference (zi_sparse_gap_factor_model.cpp:316-460
and the three other models); three free-running iterations, then every state array and the ELBO are compared.

Tolerance: 1e-7 relative on a1, a2, b1, b2, hyper-parameters, prob_S, prior_D, ELBO (north_star: 1e-6 in FP64 mode; the
stored prob_D carries an absolute error of 1.2e-10 per entry), the selection mask S exactly.
"""
import numpy as np
import pytest

import pcmf_b200
from oracle import pyoracle as po
from pcmf_b200 import datagen

pytestmark = pytest.mark.gpu

RTOL = 1e-7


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))) if a.size else 0.0


def dense_from_csr(rp, ci, vv, p):
    n = len(rp) - 1
    X = np.zeros((n, p))
    rows = np.repeat(np.arange(n), np.diff(rp))
    X[rows, ci] = vv
    return X


def run_pair(X, K, zi, sparse, niter=3, prob_S=None, prior_S=None, seed=11, gpu_input=None):
    """3 free-running iterations on both sides from the same explicit initial values; returns (oracle model, solver)."""
    n, p = X.shape
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=seed)
    if sparse and prob_S is None:
        prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    m = po.Model(X, K, zi, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    s = pcmf_b200.Solver(X if gpu_input is None else gpu_input, K, zi, sparse, a1=a1, a2=a2, b1=b1, b2=b2,
                         prob_S=prob_S if sparse else None, prior_S=prior_S if sparse else None)
    crit, elbo = s.iterate(niter, with_elbo=True)
    for it in range(niter):
        m.update_param()
        _, ng = m.gap_between_iterates()
        assert abs(crit[it] - ng) < RTOL * ng, ("conv_crit", it, crit[it], ng)
        e = m.elbo()
        assert abs(elbo[it] - e) < RTOL * abs(e), ("elbo", it, elbo[it], e)
        m.prepare_next_iterate()
    return m, s


def check_state(m, s, zi, sparse, want_prob_D=False):
    o = s.get_output(want_prob_D=want_prob_D)
    for name in ("a1", "a2", "b1", "b2", "EU", "EV", "alpha1", "alpha2", "beta1", "beta2"):
        assert relerr(o[name], m.get(name)) < RTOL, (name, relerr(o[name], m.get(name)))
    if sparse:
        assert np.max(np.abs(o["prob_S"] - m.get("prob_S"))) < RTOL
        assert (o["S"] == m.get("S")).all()
        assert relerr(o["prior_prob_S"], m.get("prior_S")) < RTOL
    if zi:
        assert relerr(o["prior_prob_D"], m.get("prior_D")) < RTOL
        if want_prob_D:
            assert np.max(np.abs(o["prob_D"] - m.get("prob_D"))) < RTOL
    return o


def test_c2_full_dense_gap():
    """BASELINE config 2 in full: dense Gamma-Poisson 10k x 2k, K = 10, no ZI, no sparsity (gap_factor_model.cpp:458-548)."""
    X, _, _ = datagen.simulate(10_000, 2_000, 10, signal_U=60.0, signal_V=60.0, prob1=None, seed=1)
    m, s = run_pair(X, 10, False, False)
    check_state(m, s, False, False)
    s.close()


def test_c3_shape_stored_prob_d_default():
    """C3-shaped ZI + sparse-V, K = 20, 12.5k x 5k = 6.25e7 entries: the stored-prob_D path is on BY DEFAULT (solver.cu
    alloc_state: n p >= 5e7) and the dense passes run multi-wave grids."""
    X, _, _ = datagen.simulate(12_500, 5_000, 20, signal_U=20.0, signal_V=20.0, prob1=0.3, seed=2)
    m, s = run_pair(X, 20, True, True, prob_S=np.full((5_000, 20), 0.5), prior_S=np.full(5_000, 0.5))
    ph = s.phase_times()
    assert ph["stored_prob_D_bytes"][1] > 0, "the default heuristics did not keep prob_D at this size"
    check_state(m, s, True, True)
    s.close()


def test_many_cells_blocked_gene_passes_default():
    """250k cells x 280 genes, ZI + sparse-V, K = 20: more than one L2-sized cell block in the gene passes by default
    (solver.cu build_csc: ~92k cells per block at K = 20), atomically combined per-block sums, 7e7 entries -> stored prob_D."""
    X, _, _ = datagen.simulate(250_000, 280, 20, signal_U=20.0, signal_V=20.0, prob1=0.3, seed=3)
    m, s = run_pair(X, 20, True, True, prob_S=np.full((280, 20), 0.5), prior_S=np.full(280, 0.5))
    ph = s.phase_times()
    assert ph["gene_pass_cell_blocks"][1] > 1, "the default heuristics did not block the gene passes at this size"
    assert ph["stored_prob_D_bytes"][1] > 0
    assert ph["tiled_nonzero_passes"][1] == 1, "the non-zero passes did not take the 2-D tiled copy of X at this size"
    check_state(m, s, True, True)
    s.close()


@pytest.mark.parametrize("zi,sparse,K", [(False, False, 14), (True, False, 20), (False, True, 16)], ids=["gap", "zi", "sparse"])
def test_tiled_nonzero_passes_are_the_default_from_65536_cells(zi, sparse, K):
    """70k cells x 330 genes: above the threshold (solver.cu build_csc) the E-step non-zero passes run on the tiled copy of X
    (xtile.cuh) without any kernel_variant bit -- 547 cell blocks x 2 gene blocks, last blocks partial; all three gene-pass
    variants (B1, B1 + x log T for the ELBO, B1 + B2) against the oracle (gap_factor_model.cpp:458-499, sparse...:327-368)."""
    X, _, _ = datagen.simulate(70_000, 330, 20, signal_U=20.0, signal_V=20.0, prob1=0.3, seed=8)
    X[X.sum(1) == 0, 0] = 1
    X[0, X.sum(0) == 0] = 1
    m, s = run_pair(X, K, zi, sparse)
    assert s.phase_times()["tiled_nonzero_passes"][1] == 1
    check_state(m, s, zi, sparse)
    s.close()


def test_tiled_nonzero_passes_are_the_default_for_dense_mid_size_matrices():
    """BASELINE config 3 is 50k x 5k at ~30 % non-zeros: fewer than 65 536 cells, but 7.5e7 non-zeros, and the tiled passes win
    there (long tile segments).  The default takes them from 16 384 cells and 5e7 non-zeros on: 17k x 4.5k at ~70 %, zero-inflated
    + sparse-V, K = 20, against the oracle."""
    rng = np.random.default_rng(12)
    n, p = 17_000, 4_500
    X = rng.poisson(1.2, (n, p)).astype(float)
    assert (X != 0).sum() >= 50_000_000
    X[X.sum(1) == 0, 0] = 1
    X[0, X.sum(0) == 0] = 1
    m, s = run_pair(X, 20, True, True, prob_S=np.full((p, 20), 0.5), prior_S=np.full(p, 0.5))
    assert s.phase_times()["tiled_nonzero_passes"][1] == 1
    check_state(m, s, True, True)
    s.close()


@pytest.mark.parametrize("K", [2, 10, 30, 64])
def test_c5_shape_latent_dimension_sweep(K):
    """BASELINE config 5 (ZI, K = 2 / 10 / 30 / 64) on a 20k x 3k sub-sample: 6e7 entries, every padded width of the sweep
    through the default dense-pass configuration (stored prob_D, gene-tile split, MT per width)."""
    X, _, _ = datagen.simulate(20_000, 3_000, 20, signal_U=20.0, signal_V=20.0, prob1=0.05, seed=4)
    X[X.sum(1) == 0, 0] = 1
    X[0, X.sum(0) == 0] = 1
    m, s = run_pair(X, K, True, False)
    assert s.phase_times()["stored_prob_D_bytes"][1] > 0
    check_state(m, s, True, False)
    s.close()


def _bench_c4_rows(nrows, prob1, mean_rate, seed=3):
    """First `nrows` cells of the benchmark's own C4 generator output (bench.py: factor_matrices + device generator)."""
    import bench
    cfg = dict(bench.CONFIGS["c4"], prob1=prob1, mean_rate=mean_rate)
    U, V = bench.factor_matrices(cfg, seed)
    csr = pcmf_b200.generate_counts_device(nrows, cfg["p"], cfg["K_true"], U[:nrows], V, prob1=prob1, seed=seed, row_offset=0)
    rp, ci, vv = csr.to_host()
    X = dense_from_csr(rp, ci, vv, cfg["p"])
    return cfg, csr, X


def test_bench_c4_generator_subsample():
    """The benchmark's own parametrisation (prob1 = 0.05, mean rate 20, prob_S = prior_S = 0.5, K = 20, p = 20k) on the
    first 3000 cells of the C4 generator output, device-resident CSR in, as bench.py feeds it."""
    cfg, csr, X = _bench_c4_rows(3000, 0.05, 20.0)
    p, K = cfg["p"], cfg["K"]
    assert X.sum(1).min() > 0
    live = X.sum(0) > 0                      # a 3000-cell sample leaves a few genes empty; the model needs every gene seen
    if not live.all():
        X = X[:, live]
        csr.free()
        m, s = run_pair(X, K, True, True, prob_S=np.full((X.shape[1], K), 0.5), prior_S=np.full(X.shape[1], 0.5))
    else:
        m, s = run_pair(X, K, True, True, prob_S=np.full((p, K), 0.5), prior_S=np.full(p, 0.5), gpu_input=csr)
    o = check_state(m, s, True, True)
    frac = o["S"].mean()
    assert 0.2 < frac < 0.8, frac            # all genes stay alive, a good share of the loadings selected
    s.close()


def test_survey_8d_parametrisation_kills_every_loading():
    """DESIGN.md section 6 claims that with SURVEY 8d's parametrisation of C4 (prob1 = 0.4, mean Poisson rate 0.13, the
    pCMF.R:228 prior_S heuristic) the REFERENCE ALGORITHM masks (almost) every loading at once, which is why bench.py does not
    use it.  Checked here with the oracle, and the CUDA path must agree with it on that degenerate trajectory too."""
    cfg, csr, X = _bench_c4_rows(3000, 0.4, 0.13)
    csr.free()
    live = X.sum(0) > 0
    X = X[:, live]
    X[X.sum(1) == 0, 0] = 1
    K = cfg["K"]
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    assert (prior_S < 0.5).mean() > 0.9      # the heuristic starts ~all genes below sel_bound
    m, s = run_pair(X, K, True, True, niter=2, prob_S=prob_S, prior_S=prior_S)
    assert m.get("S").mean() < 0.02          # the reference's update leaves (nearly) nothing selected
    check_state(m, s, True, True)
    s.close()
