"""ctypes front-end of the CPU oracle (oracle/pcmf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of pcmf_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from pcmf_b200/.
All matrices cross this boundary column-major (Fortran order), like the reference's Eigen/R data.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpcmf_oracle.so")

GAP, ZI, SPARSE, ZI_SPARSE = 0, 1, 2, 3


def kind_of(zero_inflation, sparsity):
    return {(False, False): GAP, (True, False): ZI, (False, True): SPARSE, (True, True): ZI_SPARSE}[
        (bool(zero_inflation), bool(sparsity))]


def build(force=False):
    src = os.path.join(_HERE, "pcmf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        for name in ("orc_digamma", "orc_trigamma", "orc_custom_log", "orc_custom_exp", "orc_logit", "orc_expit"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.orc_digammaInv.restype = C.c_double
        L.orc_digammaInv.argtypes = [C.c_double, C.c_int]
        for name in ("orc_Egamma", "orc_ElogGamma", "orc_gamma_entropy1", "orc_bregman_poisson", "orc_poisson_loglike1"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double, C.c_double]
        L.orc_variance.restype = C.c_double
        L.orc_variance.argtypes = [dp, C.c_int]
        L.orc_bregman_poisson_sum.restype = C.c_double
        L.orc_bregman_poisson_sum.argtypes = [dp, dp, C.c_int]
        L.orc_gamma_entropy.restype = C.c_double
        L.orc_gamma_entropy.argtypes = [dp, dp, C.c_long]
        L.orc_gamma_loglike_logX.restype = C.c_double
        L.orc_gamma_loglike_logX.argtypes = [dp, dp, dp, dp, C.c_long]
        L.orc_gamma_loglike.restype = C.c_double
        L.orc_gamma_loglike.argtypes = [dp, dp, dp, C.c_long]
        L.orc_poisson_loglike.restype = C.c_double
        L.orc_poisson_loglike.argtypes = [dp, dp, C.c_long]
        L.orc_poisson_loglike_vec.restype = C.c_double
        L.orc_poisson_loglike_vec.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_zi_poisson_loglike.restype = C.c_double
        L.orc_zi_poisson_loglike.argtypes = [dp, dp, dp, C.c_int, C.c_int]
        L.orc_bernoulli_loglike.restype = C.c_double
        L.orc_bernoulli_loglike.argtypes = [dp, dp, C.c_long]
        L.orc_bernoulli_loglike_colvec.restype = C.c_double
        L.orc_bernoulli_loglike_colvec.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_bernoulli_loglike_rowvec.restype = C.c_double
        L.orc_bernoulli_loglike_rowvec.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_copy_state.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_copy_state.restype = None
        L.orc_field.restype = dp
        L.orc_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.orc_field_S.restype = ip
        L.orc_field_S.argtypes = [C.c_void_p]
        L.orc_field_factor_order.restype = ip
        L.orc_field_factor_order.argtypes = [C.c_void_p]
        L.orc_init_sparse_param.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.orc_init_sparse_param_random.argtypes = [C.c_void_p, C.c_double]
        L.orc_init_zi_param.argtypes = [C.c_void_p]
        L.orc_init_zi_param_given.argtypes = [C.c_void_p, dp, dp]
        L.orc_init_variational_param.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.orc_init_hyper_param.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.orc_init_all_param.argtypes = [C.c_void_p] + [dp] * 8
        L.orc_init_from_factor.argtypes = [C.c_void_p, dp, dp]
        L.orc_init_random.argtypes = [C.c_void_p]
        L.orc_perturb_param.argtypes = [C.c_void_p, C.c_double]
        L.orc_set_sparse_aware.argtypes = [C.c_void_p, C.c_int]
        for name in ("orc_update_param", "orc_phase_estep", "orc_phase_mstep", "orc_phase_sparse", "orc_phase_poisson",
                     "orc_phase_zi", "orc_phase_hyper", "orc_prepare_next_iterate", "orc_order_factor", "orc_estep_only"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.orc_gap_between_iterates.argtypes = [C.c_void_p, dp, dp]
        for name in ("orc_elbo", "orc_loglikelihood", "orc_deviance", "orc_exp_deviance"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_elbo_terms.restype = dp
        L.orc_partial_deviance.restype = C.c_double
        L.orc_partial_deviance.argtypes = [C.c_void_p, ip, C.c_int]
        L.orc_optimize.restype = C.c_int
        L.orc_optimize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                   ip, ip, dp, dp, dp, dp, dp, dp, dp]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f(a):
    """column-major float64 copy + pointer"""
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    return a, a.ctypes.data_as(dp)


class Model:
    """One reference model instance (gap / zi / sparse / zi_sparse), dense, FP64, column-major."""

    def __init__(self, X, K, zero_inflation=False, sparsity=False):
        self.L = lib()
        X = np.asarray(X, dtype=np.float64)
        self.n, self.p = X.shape
        self.K = int(K)
        self.zi, self.sparse = bool(zero_inflation), bool(sparsity)
        self.kind = kind_of(zero_inflation, sparsity)
        Xf, Xp = _f(X)
        self.h = self.L.orc_create(self.kind, self.n, self.p, self.K, Xp)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- field access -------------------------------------------------------------
    _shapes = {"X": "np", "UVt": "np", "T": "np", "prob_D": "np", "EU": "nK", "ElogU": "nK", "a1": "nK", "a2": "nK",
               "a1old": "nK", "a2old": "nK", "alpha1": "nK", "alpha2": "nK", "EZ_j": "nK", "EV": "pK", "ElogV": "pK",
               "b1": "pK", "b2": "pK", "b1old": "pK", "b2old": "pK", "beta1": "pK", "beta2": "pK", "EZ_i": "pK",
               "prob_S": "pK", "prior_S": "p", "prior_D": "p", "freq_D": "p"}

    def _shape(self, name):
        s = self._shapes[name]
        return {"np": (self.n, self.p), "nK": (self.n, self.K), "pK": (self.p, self.K), "p": (self.p,)}[s]

    def get(self, name):
        if name == "S":
            ptr = self.L.orc_field_S(self.h)
            return np.ctypeslib.as_array(ptr, shape=(self.p * self.K,)).reshape((self.p, self.K), order="F").copy()
        if name == "factor_order":
            return np.ctypeslib.as_array(self.L.orc_field_factor_order(self.h), shape=(self.K,)).copy()
        ln = C.c_long(0)
        ptr = self.L.orc_field(self.h, name.encode(), C.byref(ln))
        if not ptr:
            raise KeyError(name)
        a = np.ctypeslib.as_array(ptr, shape=(ln.value,))
        return a.reshape(self._shape(name), order="F").copy()

    def set(self, name, value):
        ln = C.c_long(0)
        ptr = self.L.orc_field(self.h, name.encode(), C.byref(ln))
        a = np.ctypeslib.as_array(ptr, shape=(ln.value,))
        a[:] = np.asarray(value, dtype=np.float64).reshape(-1, order="F")

    def state(self):
        names = ["a1", "a2", "b1", "b2", "alpha1", "alpha2", "beta1", "beta2", "EU", "EV", "ElogU", "ElogV"]
        if self.sparse:
            names += ["prob_S", "prior_S", "S"]
        if self.zi:
            names += ["prior_D", "prob_D", "freq_D"]
        return {k: self.get(k) for k in names}

    # ---- init -----------------------------------------------------------------------
    def seed(self, s):
        self.L.orc_seed(self.h, int(s) & 0xFFFFFFFF)

    def init_sparse_param(self, sel_bound, prob_S=None, prior_S=None):
        if prob_S is None:
            self.L.orc_init_sparse_param_random(self.h, float(sel_bound))
        else:
            a, ap = _f(prob_S)
            b, bp = _f(prior_S)
            self.L.orc_init_sparse_param(self.h, float(sel_bound), ap, bp)

    def init_zi_param(self, prob_D=None, prior_D=None):
        if prob_D is None:
            self.L.orc_init_zi_param(self.h)
        else:
            a, ap = _f(prob_D)
            b, bp = _f(prior_D)
            self.L.orc_init_zi_param_given(self.h, ap, bp)

    def init_variational_param(self, a1, a2, b1, b2):
        k = [_f(x) for x in (a1, a2, b1, b2)]
        self.L.orc_init_variational_param(self.h, *[x[1] for x in k])

    def init_hyper_param(self, alpha1, alpha2, beta1, beta2):
        k = [_f(x) for x in (alpha1, alpha2, beta1, beta2)]
        self.L.orc_init_hyper_param(self.h, *[x[1] for x in k])

    def init_all_param(self, alpha1, alpha2, beta1, beta2, a1, a2, b1, b2):
        k = [_f(x) for x in (alpha1, alpha2, beta1, beta2, a1, a2, b1, b2)]
        self.L.orc_init_all_param(self.h, *[x[1] for x in k])

    def init_from_factor(self, U, V):
        k = [_f(x) for x in (U, V)]
        self.L.orc_init_from_factor(self.h, *[x[1] for x in k])

    def init_random(self):
        self.L.orc_init_random(self.h)

    def perturb_param(self, noise_level):
        self.L.orc_perturb_param(self.h, float(noise_level))

    def set_sparse_aware(self, flag=True):
        """Same loops, but entries with X = 0 skip the K exponentials their zero weight would be multiplied with (the
        sparse-aware CPU figure of bench.py)."""
        self.L.orc_set_sparse_aware(self.h, int(bool(flag)))

    # ---- iteration ------------------------------------------------------------------
    def update_param(self):
        self.L.orc_update_param(self.h)

    def phase(self, name):
        getattr(self.L, "orc_phase_" + name)(self.h)

    def prepare_next_iterate(self):
        self.L.orc_prepare_next_iterate(self.h)

    def gap_between_iterates(self):
        a, b = C.c_double(0), C.c_double(0)
        self.L.orc_gap_between_iterates(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def elbo(self):
        return self.L.orc_elbo(self.h)

    def elbo_terms(self):
        names = ["gammaU", "entU", "gammaV", "entV", "bernS_prior", "bernS_self", "bernD_prior", "bernD_self",
                 "sumUVt", "data", "lgammaX"]
        t = np.ctypeslib.as_array(self.L.orc_elbo_terms(), shape=(11,)).copy()
        return dict(zip(names, t))

    def loglikelihood(self):
        return self.L.orc_loglikelihood(self.h)

    def deviance(self):
        return self.L.orc_deviance(self.h)

    def exp_deviance(self):
        return self.L.orc_exp_deviance(self.h)

    def copy_from(self, other):
        """my_model = tmp_model (wrapper_gap_factor.h:339-345)."""
        self.L.orc_copy_state(self.h, other.h)

    def order_factor(self):
        self.L.orc_order_factor(self.h)

    def estep_only(self):
        self.L.orc_estep_only(self.h)

    def optimize(self, iter_max, iter_min, epsilon=1e-2, additional_iter=10, conv_mode=1, monitor=True, iter_start=0):
        it = C.c_int(int(iter_start))
        conv = C.c_int(0)
        loss = C.c_double(0)
        v = [np.zeros(max(int(iter_max), 1)) for _ in range(6)]
        ptrs = [x.ctypes.data_as(dp) for x in v]
        nb = self.L.orc_optimize(self.h, int(iter_max), int(iter_min), float(epsilon), int(additional_iter),
                                 int(conv_mode), int(bool(monitor)), C.byref(it), C.byref(conv), C.byref(loss), *ptrs)
        if nb < 0:
            raise ValueError("`conv_mode` parameter should take value in {0,1,2}")
        out = {"nb_iter": nb, "converged": bool(conv.value), "loss": loss.value, "conv_crit": v[0][:nb]}
        if monitor:
            out.update(abs_gap=v[1][:nb], norm_gap=v[2][:nb], loglikelihood=v[3][:nb], optim_criterion=v[4][:nb],
                       deviance=v[5][:nb])
        return out


def run(X, K, zero_inflation=False, sparsity=False, sel_bound=0.5, monitor=True, iter_max=1000, iter_min=500,
        epsilon=1e-2, additional_iter=10, conv_mode=1, reorder_factor=True, seed=None,
        a1=None, a2=None, b1=None, b2=None, prob_S=None, prior_S=None, ninit=1, iter_init=100, prob_D=None, prior_D=None, U=None, V=None,
        alpha1=None, alpha2=None, beta1=None, beta2=None):
    """Restatement of wrapper_gap_factor.h:164-393 / wrapper_sparse_gap_factor.h:176-435 (init order of :389-417,
    multi-init loop of :296-350 / :328-387), returning a dict shaped like the reference's R list."""
    m = Model(X, K, zero_inflation, sparsity)

    def init(mod, multi=False):
        if sparsity:
            mod.init_sparse_param(sel_bound, prob_S, prior_S)
        if zero_inflation:
            mod.init_zi_param(prob_D, prior_D)       # both given or the default (wrapper_gap_factor.h:352-356)
        if U is not None:
            mod.init_from_factor(U, V)             # wrapper_gap_factor.h:353-354
            if multi:                              # every multi-init run perturbs the given factors (:312-314)
                mod.perturb_param(max(np.max(U), np.max(V)) / K)
        elif a1 is not None:
            mod.init_variational_param(a1, a2, b1, b2)
        elif alpha1 is not None:                   # wrapper_gap_factor.h:362-363: init_hyper_param alone (else-if chain: only
            mod.init_hyper_param(alpha1, alpha2, beta1, beta2)   # reached when neither U, V nor a1..b2 are given)
        else:
            mod.init_random()

    head = None
    if ninit > 1:
        tmp = Model(X, K, zero_inflation, sparsity)      # one RNG stream for every run (the wrapper's `rng`)
        if seed is not None:
            tmp.seed(seed)
        best_loss = None
        for nrun in range(ninit):
            init(tmp, multi=True)
            c = tmp.optimize(iter_init, iter_init, 1e-6, additional_iter, conv_mode, monitor, iter_start=0)
            if nrun == 0 or c["loss"] < best_loss:           # sic: keeps the SMALLEST loss (wrapper_gap_factor.h:342)
                best_loss, head = c["loss"], c
                m.copy_from(tmp)
        conv = m.optimize(iter_max, iter_min, epsilon, additional_iter, conv_mode, monitor, iter_start=iter_init)
        for k, v in head.items():                           # entries < iter_init are the kept run's (copied members)
            if isinstance(v, np.ndarray) and k in conv:
                conv[k][:len(v)] = v
    else:
        if seed is not None:
            m.seed(seed)
        init(m)
        conv = m.optimize(iter_max, iter_min, epsilon, additional_iter, conv_mode, monitor)
    if reorder_factor:
        m.order_factor()
    st = m.state()
    res = {
        "factor": {"U": st["EU"], "V": st["EV"]},
        "variational_params": {k: st[k] for k in ("a1", "a2", "b1", "b2")},
        "hyper_params": {k: st[k] for k in ("alpha1", "alpha2", "beta1", "beta2")},
        "stats": {k: st[k] for k in ("EU", "EV", "ElogU", "ElogV")},
        "convergence": {"converged": conv["converged"], "nb_iter": conv["nb_iter"], "conv_crit": conv["conv_crit"],
                        "conv_mode": conv_mode},
        "loss": conv["loss"],
        "exp_dev": m.exp_deviance(),
        "seed": seed,
    }
    if sparsity:
        res["sparse_param"] = {"prob_S": st["prob_S"], "prior_prob_S": st["prior_S"], "S": st["S"]}
    if zero_inflation:
        res["ZI_param"] = {"prob_D": st["prob_D"], "freq_D": st["freq_D"], "prior_prob_D": st["prior_D"]}
    if monitor:
        res["monitor"] = {k: conv[k] for k in ("norm_gap", "abs_gap", "loglikelihood", "deviance", "optim_criterion")}
    res["factor_order"] = m.get("factor_order")
    return res


def prefilter(X, prop=0.05, quant_max=1.0, presel=False, threshold=0.2):
    """pkg/R/prefiltering.R:127-146 restated column by column (the R code is four apply() calls over the dense matrix):
    crit1 :130, crit2 :132, crit3 :134 (quantile type 7), crit4 :137-140.  Checker only."""
    X = np.asarray(X, float)
    n, p = X.shape
    xmax = np.array([X[:, j].max() for j in range(p)])
    nnz = np.array([(X[:, j] != 0).sum() for j in range(p)])
    srt = np.sort(xmax)                                    # quantile(type = 7): h = (p - 1) q, linear between order statistics
    h = (p - 1) * quant_max
    lo = int(np.floor(h))
    q = srt[lo] + (h - lo) * (srt[min(lo + 1, p - 1)] - srt[lo])
    keep = (xmax > 0) & (nnz >= prop * n) & (xmax <= q)
    if presel:
        nzmean = X[X != 0].sum() / (X != 0).sum()
        sd = np.array([np.sqrt(((X[:, j] - X[:, j].mean()) ** 2).sum() / (n - 1)) for j in range(p)])
        keep &= (1 - np.exp(-sd / nzmean)) > threshold
    return keep
