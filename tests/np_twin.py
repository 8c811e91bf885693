"""numpy "design twin": the VEM iteration written with the sparse / closed-form identities the GPU path relies on
 instead of the reference's dense loops.  Test infrastructure: it (a) cross-checks the C
oracle with an independent implementation and (b) proves the algebra of the CUDA design (hoisted exponentials,
prob_D recomputed from saved (EU, EVS, prior_D) instead of stored, lagged ELBO pieces) before any kernel exists.
"""
import numpy as np
from scipy import special as sp


def logit(x):
    x = np.asarray(x, dtype=float)
    out = np.log(np.where((x > 1e-12) & (x < 1 - 1e-12), x, 0.5) / (1 - np.where((x > 1e-12) & (x < 1 - 1e-12), x, 0.5)))
    out = np.where(x >= 1 - 1e-12, 30.0, out)
    return np.where(x <= 1e-12, -30.0, out)


def expit(x):
    x = np.asarray(x, dtype=float)
    out = 1.0 / (1 + np.exp(-np.clip(x, -40, 40)))
    out = np.where(x >= 30, 1.0, out)
    return np.where(x <= -30, 0.0, out)


def digamma_inv(y, nit=6):
    y = np.asarray(y, dtype=float)
    x = np.where(y >= -2.22, np.exp(y) + 0.5, -1.0 / (y - sp.digamma(1.0)))
    for _ in range(nit):
        x = x - (sp.digamma(x) - y) / sp.polygamma(1, x)
    return x


class Twin:
    def __init__(self, X, K, zi, sparse, sel_bound=0.5):
        self.X = np.asarray(X, dtype=float)
        self.n, self.p = self.X.shape
        self.K, self.zi, self.sparse, self.sel_bound = K, zi, sparse, sel_bound
        self.nz = self.X != 0
        n, p = self.n, self.p
        self.alpha1, self.alpha2 = np.ones(K), np.ones(K)     # rows identical -> keep one row
        self.beta1, self.beta2 = np.ones(K), np.ones(K)
        self.a1o, self.a2o, self.b1o, self.b2o = np.ones((n, K)), np.ones((n, K)), np.ones((p, K)), np.ones((p, K))
        self.prob_S, self.S, self.prior_S = np.ones((p, K)), np.ones((p, K)), np.ones(p)
        self.first = True   # prob_D still equals (X>0)

    def init(self, a1, a2, b1, b2, prob_S=None, prior_S=None):
        if self.sparse:
            self.prob_S, self.prior_S = prob_S.copy(), prior_S.copy()
            self.S = (self.prob_S >= self.sel_bound).astype(float)
        if self.zi:
            self.prior_D = self.nz.mean(0)
        self.a1, self.a2, self.b1, self.b2 = a1.copy(), a2.copy(), b1.copy(), b2.copy()
        self.stats_U()
        self.stats_V()
        self.hyper()

    def stats_U(self):
        self.EU, self.ElogU = self.a1 / self.a2, sp.digamma(self.a1) - np.log(self.a2)

    def stats_V(self):
        self.EV, self.ElogV = self.b1 / self.b2, sp.digamma(self.b1) - np.log(self.b2)

    def EVS(self):
        return self.EV * self.prob_S if self.sparse else self.EV

    def hyper(self):
        self.alpha1 = digamma_inv(np.log(self.alpha2) + self.ElogU.mean(0))
        self.alpha2 = self.alpha1 / self.EU.mean(0)
        self.beta1 = digamma_inv(np.log(self.beta2) + self.ElogV.mean(0))
        self.beta2 = self.beta1 / self.EV.mean(0)
        if self.sparse:
            self.prior_S = self.prob_S.mean(1)
        if self.zi:
            if self.first:
                self.prior_D = self.nz.mean(0)
            else:
                self.prior_D = self.colsumP / self.n

    def P_formula(self):
        """prob_D on every entry as the saved-state formula; exact entries (nnz -> 1, prior shortcuts) applied on top."""
        if self.first:
            return self.nz.astype(float)
        P = expit(self.cz[None, :] - self.EUz @ self.EVSz.T)
        P = np.where(self.nz, 1.0, P)
        P = np.where(self.pz[None, :] == 1, 1.0, P)
        P = np.where(self.pz[None, :] == 0, 0.0, P)
        return P

    def iterate(self):
        X, nz = self.X, self.nz
        # ---- E-step with hoisted exponentials (valid while the +-100 guards are inactive)
        eu = np.exp(self.ElogU)
        ev = np.exp(self.ElogV) * self.S
        evs = ev * self.prob_S if self.sparse else ev
        T = eu @ ev.T
        Tp = np.where(T > 0, T, 1.0)
        w = X.copy()
        P = self.P_formula() if self.zi else None
        if self.zi:
            w = w * P
        Q = np.where(nz, w / Tp, 0.0)
        EZ_j = eu * (Q @ evs)
        EZ_i = evs * (Q.T @ eu)
        # ---- M-step, U then V (Gauss-Seidel)
        EVS = self.EVS()
        self.a1 = self.alpha1[None, :] + EZ_j
        if self.zi:
            self.a2 = self.alpha2[None, :] + P @ EVS
        else:
            self.a2 = np.repeat(self.alpha2[None, :] + EVS.sum(0)[None, :], self.n, 0)
        self.stats_U()
        self.b1 = self.beta1[None, :] + EZ_i
        tmpMat = P.T @ self.EU if self.zi else np.repeat(self.EU.sum(0)[None, :], self.p, 0)
        self.b2 = self.beta2[None, :] + (self.prob_S * tmpMat if self.sparse else tmpMat)
        self.stats_V()
        # ---- sparse probabilities: second nnz pass with the new stats and the OLD mask
        if self.sparse:
            eu2 = np.exp(self.ElogU)
            ev2 = np.exp(self.ElogV) * self.S
            T2 = eu2 @ ev2.T
            T2 = np.where(T2 > 0, T2, 1.0)
            w2 = X * P if self.zi else X
            Q2 = np.where(nz, w2 / T2, 0.0)
            A1 = Q2.T @ eu2
            A2 = Q2.T @ (eu2 * self.ElogU)
            tmp = -self.EV * tmpMat + ev2 * (A2 + self.ElogV * A1)
            new = expit(logit(self.prior_S)[:, None] + tmp)
            new = np.where(self.prior_S[:, None] == 1, 1.0, new)
            new = np.where(self.prior_S[:, None] == 0, 0.0, new)
            self.prob_S = new
            self.S = (self.prob_S >= self.sel_bound).astype(float)
        # ---- ZI probabilities: only the defining triple is saved, plus column sums
        if self.zi:
            self.EUz, self.EVSz = self.EU.copy(), self.EVS().copy()
            self.pz = self.prior_D.copy()
            self.cz = logit(self.prior_D)
            self.first = False
            self.colsumP = self.P_formula().sum(0)
        self.hyper()

    def gap(self):
        d = ((self.a1 - self.a1o) ** 2).sum() + ((self.a2 - self.a2o) ** 2).sum() + ((self.b1 - self.b1o) ** 2).sum() \
            + ((self.b2 - self.b2o) ** 2).sum()
        nrm = (self.a1o ** 2).sum() + (self.a2o ** 2).sum() + (self.b1o ** 2).sum() + (self.b2o ** 2).sum()
        return np.sqrt(d), np.sqrt(d) / np.sqrt(nrm)

    def next(self):
        self.a1o, self.a2o, self.b1o, self.b2o = self.a1.copy(), self.a2.copy(), self.b1.copy(), self.b2.copy()

    def elbo(self):
        n, p, K = self.n, self.p, self.K
        X, nz = self.X, self.nz
        res = 0.0
        # Gamma parts, closed form over identical alpha rows
        res += (n * (self.alpha1 * np.log(self.alpha2) - sp.gammaln(self.alpha1)) + (self.alpha1 - 1) * self.ElogU.sum(0)
                - self.alpha2 * self.EU.sum(0)).sum()
        res += ((1 - self.a1) * sp.digamma(self.a1) + self.a1 + sp.gammaln(self.a1) - np.log(self.a2)).sum()
        res += (p * (self.beta1 * np.log(self.beta2) - sp.gammaln(self.beta1)) + (self.beta1 - 1) * self.ElogV.sum(0)
                - self.beta2 * self.EV.sum(0)).sum()
        res += ((1 - self.b1) * sp.digamma(self.b1) + self.b1 + sp.gammaln(self.b1) - np.log(self.b2)).sum()
        if self.sparse:
            g = (self.prior_S > 0) & (self.prior_S < 1)
            ps, pr = self.prob_S[g], self.prior_S[g][:, None]
            res += (ps * np.log(pr) + (1 - ps) * np.log(1 - pr)).sum()
            q = self.prob_S[(self.prob_S > 0) & (self.prob_S < 1)]
            res -= (q * np.log(q) + (1 - q) * np.log(1 - q)).sum()
        EVS = self.EVS()
        if self.zi:
            P = self.P_formula()
            c = P.sum(0)
            g = (self.prior_D > 0) & (self.prior_D < 1)
            res += (c[g] * np.log(self.prior_D[g]) + (n - c[g]) * np.log(1 - self.prior_D[g])).sum()
            q = P[(P > 0) & (P < 1)]
            res -= (q * np.log(q) + (1 - q) * np.log(1 - q)).sum()
            res -= (P * (self.EU @ EVS.T)).sum()
        else:
            P = None
            res -= (self.EU.sum(0) * EVS.sum(0)).sum()
        eu = np.exp(self.ElogU)
        ev = np.exp(self.ElogV) * self.S
        T = eu @ ev.T
        Tp = np.where(T > 0, T, 1.0)
        w = X * P if self.zi else X
        if not self.sparse:
            res += (w[nz] * np.log(T[nz])).sum()
        else:
            Q = np.where(nz, w / Tp, 0.0)
            A1 = Q.T @ eu
            A2 = Q.T @ (eu * self.ElogU)
            res += (self.prob_S * ev * (A2 + self.ElogV * A1)).sum()
        res -= ((P[nz] if self.zi else 1.0) * sp.gammaln(X[nz] + 1)).sum()
        return res
