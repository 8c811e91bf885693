"""Synthetic count data following the reference package's own generative model.

Mirrors `generate_factor_matrix` (reference pkg/R/data_generation.R:114-275) and `generate_count_matrix`
(:393-463): block-structured factor matrices, Lambda = U V^T, X = Poisson(Lambda) * Bernoulli(prob1_j) where
prob1 is the probability of NOT dropping out (:331-335, :443).  R's RNG stream is not reproducible outside R, so
the draws come from numpy's Philox generator; only the distribution is the same.

Host (numpy) generation is meant for the small parity configurations; the large benchmark configurations
(1M x 20k) are generated on the GPU straight into CSR by `pcmf_b200.api.generate_counts_device`.
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def generate_factor_matrix(nfeature, K, ngroup=1, average_signal=1.0, disp=None, group_separation=0.5,
                           distribution="uniform", shuffle_feature=True, prop_noise_feature=0.0, noise_level=None,
                           seed=None, rng=None):
    """data_generation.R:114-275.  Returns dict(factor_matrix, feature_order, feature_label)."""
    if not (isinstance(nfeature, (int, np.integer)) and nfeature > 0):
        raise ValueError("'nfeature' input parameter should be a positive integer")
    if not (isinstance(K, (int, np.integer)) and K > 0):
        raise ValueError("'K' input parameter should be a positive integer")
    if not (0 < ngroup <= K):
        raise ValueError("'ngroup' input parameter should be a positive integer <= 'K'")
    if not (0 <= group_separation <= 1):
        raise ValueError("'group_separation' input parameter should be a real value in [0,1]")
    if distribution not in ("uniform", "gamma", "exponential"):
        raise ValueError("'distribution' input parameter should be \"uniform\" or \"gamma\"")
    if not (0 <= prop_noise_feature <= 1):
        raise ValueError("'prop_noise_feature' input parameter should be a real value in [0,1]")
    rng = rng if rng is not None else _rng(seed)
    avg = np.broadcast_to(np.asarray(average_signal, dtype=float), (ngroup,)).copy()
    if disp is None:
        disp = 1.0
    # block parameter matrix, ngroup x ngroup (:175-184)
    param_block = np.diag(avg)
    disp_block = np.eye(ngroup)
    if ngroup > 1:
        param_block = param_block + avg.mean() * (1 - group_separation) * (1 - np.eye(ngroup))
        disp_block = disp_block + (1 - np.eye(ngroup))
    col_order = rng.permutation(ngroup)          # shuffle groups (:186-188)
    param_block = param_block[:, col_order]
    disp_block = disp_block[:, col_order]
    nfeature0 = int(np.floor((1 - prop_noise_feature) * nfeature))
    sqK = np.sqrt(K)
    mats = []
    id_rows = np.zeros(0, dtype=int)
    if nfeature0 > 0:
        if ngroup > 1:
            # sort(rep(id_block, length=...)) (:204-207)
            id_rows = np.sort(np.resize(np.arange(ngroup), nfeature0))
            id_cols = np.sort(np.resize(np.arange(ngroup), K))
            param = param_block[np.ix_(id_rows, id_cols)]
            dispm = disp_block[np.ix_(id_rows, id_cols)]
        else:
            id_rows = np.zeros(nfeature0, dtype=int)
            param = np.full((nfeature0, K), avg[0])
            dispm = np.full((nfeature0, K), float(disp))
        if distribution == "uniform":
            mat0 = rng.uniform(0.0, 1.0, (nfeature0, K)) * (2 * param / sqK)
        elif distribution == "gamma":
            mat0 = rng.gamma(shape=param / sqK, scale=1.0 / dispm)
        else:  # exponential: rgamma(shape=1, rate=sqrt(K)/param) (:235)
            mat0 = rng.exponential(1.0, (nfeature0, K)) * (param / sqK)
        mats.append(mat0)
    nnoise = nfeature - nfeature0
    if noise_level is None:
        noise_level = avg.mean() * (1 - group_separation)
    if nnoise > 0:
        if distribution == "uniform":
            mn = rng.uniform(0.0, 2 * noise_level / sqK, (nnoise, K))
        elif distribution == "gamma":
            mn = rng.gamma(shape=noise_level / sqK, scale=1.0 / np.min(disp), size=(nnoise, K))
        else:
            mn = rng.exponential(1.0, (nnoise, K)) * (noise_level / sqK)
        mats.append(mn)
    mat = np.vstack(mats)
    order = rng.permutation(nfeature) if shuffle_feature else np.arange(nfeature)
    mat = mat[order]
    label = np.concatenate([id_rows + 1, np.zeros(nnoise, dtype=int)])[order]
    return {"factor_matrix": mat, "feature_order": order, "feature_label": label}


def generate_count_matrix(n, p, K, U, V, ZI=False, prob1=None, rate0=None, seed=None, rng=None):
    """data_generation.R:393-463.  Returns dict(X, U, V, n, p, K, ZI, prob1, rate0, Xnzi, ZIind)."""
    U = np.asarray(U, dtype=float)
    V = np.asarray(V, dtype=float)
    if U.shape != (n, K) or (U < 0).any():
        raise ValueError("'U' input parameter should be a positive-valued matrix of dimension 'n' x 'K'")
    if V.shape != (p, K) or (V < 0).any():
        raise ValueError("'V' input parameter should be a positive-valued matrix of dimension 'p' x 'K'")
    if ZI and prob1 is None and rate0 is None:
        raise ValueError("It is required to use a zero-inflated model to generate data ('ZI' input parameter is set "
                         "to TRUE) however neither 'prob1' nor 'rate0' input parameters are given")
    rng = rng if rng is not None else _rng(seed)
    Lambda = U @ V.T
    Xnzi = rng.poisson(Lambda).astype(np.int64)
    if ZI:
        if rate0 is not None:
            meanX = Xnzi.mean(axis=0)
            prob1 = 1 - np.exp(-np.asarray(rate0) * meanX ** 2)
        prob1 = np.asarray(prob1, dtype=float)
        if prob1.shape != (p,):
            raise ValueError("if used, 'prob1' input parameter should be a vector of size 'p'")
        D = (rng.uniform(size=(n, p)) < prob1[None, :]).astype(np.int64)
    else:
        D = np.ones((n, p), dtype=np.int64)
    X = Xnzi * D
    return {"X": X, "U": U, "V": V, "n": n, "p": p, "K": K, "ZI": ZI, "prob1": prob1, "rate0": rate0,
            "Xnzi": Xnzi if ZI else None, "ZIind": D if ZI else None}


def simulate(n, p, K_true, ngroup_U=3, ngroup_V=2, signal_U=60.0, signal_V=60.0, group_separation=0.8,
             prop_noise_feature=0.6, prob1=None, seed=0, redraw_empty=True):
    """The benchmark's synthetic input: block exponential factors -> Poisson -> drop-out.
    Empty rows/columns are given one count so the reference's random init (sqrt(K/mean)) stays finite."""
    rng = _rng(seed)
    U = generate_factor_matrix(n, K_true, ngroup=ngroup_U, average_signal=signal_U, group_separation=group_separation,
                               distribution="exponential", shuffle_feature=True, rng=rng)["factor_matrix"]
    V = generate_factor_matrix(p, K_true, ngroup=ngroup_V, average_signal=signal_V, group_separation=group_separation,
                               distribution="exponential", shuffle_feature=True,
                               prop_noise_feature=prop_noise_feature, rng=rng)["factor_matrix"]
    zi = prob1 is not None
    pr = np.full(p, float(prob1)) if zi else None
    X = generate_count_matrix(n, p, K_true, U, V, ZI=zi, prob1=pr, rng=rng)["X"]
    if redraw_empty:
        for i in np.where(X.sum(1) == 0)[0]:
            X[i, rng.integers(p)] = 1
        for j in np.where(X.sum(0) == 0)[0]:
            X[rng.integers(n), j] = 1
    return X, U, V


def random_init(X, K, seed=0):
    """Initial variational parameters with the reference's random-init *distribution*
    (gap_factor_model.cpp:147-168): a2=b2=1, a1[i,.] ~ Gamma(1, rate sqrt(K/rowmean_i)),
    b1[j,.] ~ Gamma(1, rate sqrt(K/colmean_j)).  Passed explicitly to both the oracle and the GPU path."""
    rng = _rng(seed)
    X = np.asarray(X)
    n, p = X.shape
    rm = X.mean(1)
    cm = X.mean(0)
    a1 = rng.exponential(1.0, (n, K)) / np.sqrt(K / rm)[:, None]
    b1 = rng.exponential(1.0, (p, K)) / np.sqrt(K / cm)[:, None]
    return a1, np.ones((n, K)), b1, np.ones((p, K))


def prior_S_heuristic(X, K):
    """pCMF.R:228-229: prior_S = 1 - exp(-sd_j / mean(X[X != 0])) (R's sd is the n-1 estimator), prob_S = rep(prior_S, K)."""
    X = np.asarray(X, dtype=float)
    sd = X.std(axis=0, ddof=1)
    prior_S = 1 - np.exp(-sd / X[X != 0].mean())
    prob_S = np.repeat(prior_S[:, None], K, axis=1)
    return prob_S, prior_S
