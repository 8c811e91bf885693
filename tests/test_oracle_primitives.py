"""Pins the oracle's primitives against every known-answer test the reference holds for this path
 and against scipy for the restated Boost special functions."""
import numpy as np
import pytest
from scipy import special as sp

from oracle import pyoracle as po

L = po.lib()


def arr(x):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    return a, a.ctypes.data_as(po.dp)


# ---- reference test_internal.cpp:35-41 -------------------------------------------------
def test_digamma_inverse_golden():
    assert abs(L.orc_digammaInv(2.0, 5) - 7.8834286312) < 1e-8
    assert abs(L.orc_digammaInv(10.0, 5) - 22026.9657929151) < 1e-8


# ---- test_internal.cpp:43-49 -----------------------------------------------------------
def test_variance_golden():
    a, p = arr(np.arange(1, 10))
    assert abs(L.orc_variance(p, 9) - 7.5) < 1e-8


# ---- test_internal.cpp:51-56 -----------------------------------------------------------
def test_custom_log():
    assert L.orc_custom_log(0.0) == 0.0
    assert L.orc_custom_log(7.5) == np.log(7.5)


def test_custom_exp():
    # internal.h:121-127
    assert L.orc_custom_exp(-100.0) == 3e-44
    assert L.orc_custom_exp(-150.0) == 3e-44
    assert L.orc_custom_exp(-99.0) == np.exp(-99.0)
    assert L.orc_custom_exp(1.5) == np.exp(1.5)


# ---- test_internal.cpp:58-68 -----------------------------------------------------------
def test_logit_golden():
    assert L.orc_logit(0.5) == 0
    assert L.orc_logit(0.0) == -30
    assert L.orc_logit(1.0) == 30
    assert abs(L.orc_logit(0.2) - np.log(0.2 / 0.8)) < 1e-8
    assert L.orc_logit(1e-12) == -30 and L.orc_logit(1 - 1e-12) == 30


# ---- test_internal.cpp:70-80 -----------------------------------------------------------
def test_expit_golden():
    assert L.orc_expit(0.0) == 0.5
    assert L.orc_expit(-30.0) == 0
    assert L.orc_expit(30.0) == 1
    for x in (-2.0, 2.0):
        assert abs(L.orc_expit(x) - 1.0 / (1 + np.exp(-x))) < 1e-8


# ---- test_probability.cpp:38-108 (formula identities) ----------------------------------
def test_gamma_moments_identities():
    for a, b in [(2.0, 3.0), (0.3, 7.0), (150.0, 0.02)]:
        assert L.orc_Egamma(a, b) == a / b
        assert abs(L.orc_ElogGamma(a, b) - (sp.digamma(a) - np.log(b))) < 1e-13
        ent = (1 - a) * sp.digamma(a) + a + sp.gammaln(a) - np.log(b)
        assert abs(L.orc_gamma_entropy1(a, b) - ent) < 1e-12 * max(1, abs(ent))
    rng = np.random.default_rng(0)
    p1 = rng.uniform(0.1, 9, 12)
    p2 = rng.uniform(0.1, 9, 12)
    a1, q1 = arr(p1)
    a2, q2 = arr(p2)
    tot = ((1 - p1) * sp.digamma(p1) + p1 + sp.gammaln(p1) - np.log(p2)).sum()
    assert abs(L.orc_gamma_entropy(q1, q2, 12) - tot) < 1e-11


# ---- test_probability.cpp:110-135 ------------------------------------------------------
def test_bregman_poisson_golden():
    assert L.orc_bregman_poisson(1.0, 1.0) == 0
    assert L.orc_bregman_poisson(10.0, 10.0) == 0
    assert abs(L.orc_bregman_poisson(5.0, 10.0) - (5 * np.log(0.5) - 5 + 10)) < 1e-8
    X = np.array([[2, 4, 3], [10, 7, 3], [9, 3, 11], [3, 9, 3]], float)
    Y = np.array([[7, 11, 1], [2, 10, 4], [3, 2, 6], [7, 6, 2]], float)
    a, pa = arr(X)
    b, pb = arr(Y)
    assert abs(L.orc_bregman_poisson_sum(pa, pb, 12) - 23.57360306) < 1e-8


# ---- test_likelihood.cpp:39-172 (identities) -------------------------------------------
def test_likelihood_identities():
    rng = np.random.default_rng(1)
    n, p = 6, 5
    X = rng.poisson(3.0, (n, p)).astype(float)
    X[0, 0] = 0
    rate = rng.uniform(0.5, 5, (n, p))
    Xf, Xp = po._f(X)
    rf, rp = po._f(rate)
    ref = (X * np.log(rate) - rate - sp.gammaln(X + 1)).sum()
    assert abs(L.orc_poisson_loglike(Xp, rp, n * p) - ref) < 1e-10
    # zero rate -> custom_log gives 0 (likelihood.h:153)
    r0 = rate.copy()
    r0[1, 1] = 0
    r0f, r0p = po._f(r0)
    mask = r0 > 0
    ref0 = (X * np.where(mask, np.log(np.where(mask, r0, 1)), 0) - r0 - sp.gammaln(X + 1)).sum()
    assert abs(L.orc_poisson_loglike(Xp, r0p, n * p) - ref0) < 1e-10
    cm = X.mean(0)
    cmf, cmp_ = arr(cm)
    refv = (X * np.log(cm)[None, :] - cm[None, :] - sp.gammaln(X + 1)).sum()
    assert abs(L.orc_poisson_loglike_vec(Xp, cmp_, n, p) - refv) < 1e-10
    prob = rng.uniform(0.05, 0.95, (n, p))
    pf, pp = po._f(prob)
    refzi = np.where(X > 0, np.log(prob) + X * np.log(rate) - rate - sp.gammaln(X + 1),
                     np.log(1 - prob + prob * np.exp(-rate))).sum()
    assert abs(L.orc_zi_poisson_loglike(Xp, rp, pp, n, p) - refzi) < 1e-10
    # Bernoulli (test_likelihood.cpp:150-172), gate 0<prob<1
    B = (rng.uniform(size=(n, p)) < 0.5).astype(float)
    pr = prob.copy()
    pr[0, 0] = 0.0
    pr[1, 1] = 1.0
    Bf, Bp = po._f(B)
    prf, prp = po._f(pr)
    g = (pr > 0) & (pr < 1)
    refb = (B[g] * np.log(pr[g]) + (1 - B[g]) * np.log(1 - pr[g])).sum()
    assert abs(L.orc_bernoulli_loglike(Bp, prp, n * p) - refb) < 1e-10
    pc = np.array([0.2, 0.0, 1.0, 0.7, 0.5])
    pcf, pcp = arr(pc)
    gc = (pc > 0) & (pc < 1)
    refc = (B[:, gc] * np.log(pc[gc]) + (1 - B[:, gc]) * np.log(1 - pc[gc])).sum()
    assert abs(L.orc_bernoulli_loglike_colvec(Bp, pcp, n, p) - refc) < 1e-10
    # gamma log-likelihood (likelihood.h:85-113)
    a = rng.uniform(0.2, 4, (n, p))
    b = rng.uniform(0.2, 4, (n, p))
    G = rng.gamma(2.0, 1.0, (n, p))
    af, ap = po._f(a)
    bf, bp = po._f(b)
    Gf, Gp = po._f(G)
    lG, lGp = po._f(np.log(G))
    refg = (a * np.log(b) + (a - 1) * np.log(G) - b * G - sp.gammaln(a)).sum()
    assert abs(L.orc_gamma_loglike(Gp, ap, bp, n * p) - refg) < 1e-10
    assert abs(L.orc_gamma_loglike_logX(Gp, lGp, ap, bp, n * p) - refg) < 1e-10


# ---- restated Boost special functions vs scipy -----------------------------------------
@pytest.mark.parametrize("x", [1e-8, 1e-6, 1e-3, 0.1, 0.5, 1.0, 1.4616321449683623, 2.0, 3.7, 9.99, 10.0, 55.5, 1e3, 1e6, 1e12])
def test_digamma_trigamma_vs_scipy(x):
    d, t = L.orc_digamma(x), L.orc_trigamma(x)
    rd, rt = sp.digamma(x), sp.polygamma(1, x)
    assert abs(d - rd) <= 4e-15 * max(1.0, abs(rd))
    assert abs(t - rt) <= 1e-14 * max(1.0, abs(rt))


def test_digamma_inverse_roundtrip():
    for y in (-50.0, -5.0, -2.22, -2.21, 0.0, 1.0, 4.0):
        x = L.orc_digammaInv(y, 6)
        assert x > 0
        assert abs(sp.digamma(x) - y) < 1e-9 * max(1, abs(y))
