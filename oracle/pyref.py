"""ctypes access to oracle/_ref/libpcmf_ref.so: the reference's own C++ classes (variational_em_algo<model>) compiled
from /root/reference against the stand-in headers of oracle/shim/.  TEST INFRASTRUCTURE ONLY (validates the oracle).

`run()` mirrors oracle.pyoracle.run() / the single-init branch of wrapper_gap_factor.h:352-392."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libpcmf_ref.so")
REF_SRC = "/root/reference/pkg/src"
dp = C.POINTER(C.c_double)
_lib = None


def available():
    return os.path.exists(SO) or os.path.isdir(REF_SRC)


def build(force=False):
    """Compiles the reference sources where they lie (only possible where /root/reference exists)."""
    if not os.path.isdir(REF_SRC):
        return os.path.exists(SO)
    r = subprocess.run(["make", "-C", HERE, "ref"] + (["-B"] if force else []), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference build failed:\n" + r.stdout)
    return True


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int] * 4 + [dp] + [C.c_int] * 3 + [C.c_double, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_last_error.argtypes = [C.c_void_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_init_sparse.argtypes = [C.c_void_p, C.c_double, dp, dp, C.c_int, C.c_int]
        L.ref_init_sparse_random.argtypes = [C.c_void_p, C.c_double]
        L.ref_init_zi.argtypes = [C.c_void_p]
        L.ref_init_zi_given.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int]
        L.ref_init_variational.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int]
        L.ref_init_hyper.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int]
        L.ref_init_factor.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int]
        L.ref_perturb_param.argtypes = [C.c_void_p, C.c_double]
        for nm in ("ref_init_random", "ref_optimize", "ref_order_factor", "ref_output"):
            getattr(L, nm).argtypes = [C.c_void_p]
        L.ref_get.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, dp, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_get.restype = C.c_long
        for nm in ("ref_logit", "ref_expit", "ref_custom_log", "ref_custom_exp", "ref_digamma", "ref_trigamma"):
            getattr(L, nm).argtypes = [C.c_double]
            getattr(L, nm).restype = C.c_double
        L.ref_digammaInv.argtypes = [C.c_double, C.c_int]
        L.ref_digammaInv.restype = C.c_double
        _lib = L
    return _lib


def _f(a):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    return a, a.ctypes.data_as(dp)


class RefModel:
    def __init__(self, X, K, zero_inflation, sparsity, monitor=True, iter_max=1000, iter_min=500, epsilon=1e-2,
                 additional_iter=10, conv_mode=1):
        self.L = lib()
        X = np.asarray(X, dtype=np.float64)
        self.n, self.p = X.shape
        self.K = int(K)
        self.zi, self.sparse = bool(zero_inflation), bool(sparsity)
        kind = (1 if self.zi else 0) + (2 if self.sparse else 0)
        Xf, Xp = _f(X)
        self.h = self.L.ref_create(kind, self.n, self.p, self.K, Xp, int(bool(monitor)), int(iter_max), int(iter_min),
                                   float(epsilon), int(additional_iter), int(conv_mode))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.ref_last_error(self.h).decode())

    def seed(self, s):
        self.L.ref_seed(self.h, int(s))

    def init_sparse_param(self, sel_bound, prob_S=None, prior_S=None):
        if prob_S is None:
            self._ck(self.L.ref_init_sparse_random(self.h, float(sel_bound)))
        else:
            a, ap = _f(prob_S)
            b, bp = _f(prior_S)
            self._ck(self.L.ref_init_sparse(self.h, float(sel_bound), ap, bp, self.p, self.K))

    def init_zi_param(self, prob_D=None, prior_D=None):
        if prob_D is None:
            self._ck(self.L.ref_init_zi(self.h))
        else:
            a, ap = _f(prob_D)
            b, bp = _f(prior_D)
            self._ck(self.L.ref_init_zi_given(self.h, ap, bp, self.n, self.p))

    def init_variational_param(self, a1, a2, b1, b2):
        arrs = [_f(x) for x in (a1, a2, b1, b2)]
        self._ck(self.L.ref_init_variational(self.h, *[a[1] for a in arrs], self.n, self.p, self.K))

    def init_random(self):
        self._ck(self.L.ref_init_random(self.h))

    def init_hyper_param(self, alpha1, alpha2, beta1, beta2):
        arrs = [_f(x) for x in (alpha1, alpha2, beta1, beta2)]
        self._ck(self.L.ref_init_hyper(self.h, *[a[1] for a in arrs], self.n, self.p, self.K))

    def perturb_param(self, noise_level):
        self._ck(self.L.ref_perturb_param(self.h, float(noise_level)))

    def init_from_factor(self, U, V):
        a, ap = _f(U)
        b, bp = _f(V)
        self._ck(self.L.ref_init_factor(self.h, ap, bp, self.n, self.p, self.K))

    def optimize(self):
        self._ck(self.L.ref_optimize(self.h))

    def order_factor(self):
        self._ck(self.L.ref_order_factor(self.h))

    def output(self):
        """algorithm::get_output + model::get_output as a nested dict (the R return list)."""
        self._ck(self.L.ref_output(self.h))
        n, p, K = self.n, self.p, self.K
        buf = np.zeros(max(n * p, n * K, p * K, 1) + 16)
        bp = buf.ctypes.data_as(dp)

        def get(group, name):
            r, c = C.c_int(0), C.c_int(0)
            ln = self.L.ref_get(self.h, group.encode(), name.encode(), bp, buf.size, C.byref(r), C.byref(c))
            if ln < 0:
                return None
            a = buf[:ln].copy()
            if r.value * c.value == ln and ln > 1 and c.value > 1:
                return a.reshape((r.value, c.value), order="F")
            return a if ln != 1 or (r.value, c.value) != (1, 1) else float(a[0])

        out = {}
        groups = {"factor": ("U", "V"), "variational_params": ("a1", "a2", "b1", "b2"),
                  "hyper_params": ("alpha1", "alpha2", "beta1", "beta2"), "stats": ("EU", "EV", "ElogU", "ElogV"),
                  "sparse_param": ("prob_S", "prior_prob_S", "S"), "ZI_param": ("prob_D", "freq_D", "prior_prob_D"),
                  "convergence": ("converged", "nb_iter", "conv_crit", "conv_mode"),
                  "monitor": ("norm_gap", "abs_gap", "loglikelihood", "deviance", "optim_criterion")}
        for g, names in groups.items():
            d = {k: get(g, k) for k in names}
            if any(v is not None for v in d.values()):
                out[g] = {k: v for k, v in d.items() if v is not None}
        out["loss"] = get("", "loss")
        out["exp_dev"] = get("", "exp_dev")
        if "convergence" in out:
            cv = out["convergence"]
            cv["nb_iter"] = int(cv["nb_iter"])
            cv["converged"] = bool(cv["converged"])
            for g in ("convergence", "monitor"):
                for k, v in out.get(g, {}).items():
                    if isinstance(v, float) and k in ("conv_crit", "norm_gap", "abs_gap", "loglikelihood", "deviance", "optim_criterion"):
                        out[g][k] = np.array([v])
        if "sparse_param" in out:
            out["sparse_param"]["S"] = out["sparse_param"]["S"].astype(np.int32)
        return out


def run(X, K, zero_inflation=False, sparsity=False, sel_bound=0.5, monitor=True, iter_max=1000, iter_min=500, epsilon=1e-2,
        additional_iter=10, conv_mode=1, reorder_factor=True, seed=None, a1=None, a2=None, b1=None, b2=None,
        prob_S=None, prior_S=None, prob_D=None, prior_D=None, U=None, V=None, alpha1=None, alpha2=None, beta1=None, beta2=None):
    m = RefModel(X, K, zero_inflation, sparsity, monitor, iter_max, iter_min, epsilon, additional_iter, conv_mode)
    if seed is not None:
        m.seed(seed)
    if sparsity:                                   # wrapper_sparse_gap_factor.h:389-394
        m.init_sparse_param(sel_bound, prob_S, prior_S)
    if zero_inflation:                             # :395-399
        m.init_zi_param(prob_D, prior_D)
    if U is not None:                              # :400-401
        m.init_from_factor(U, V)
    elif a1 is not None:                           # :402-403
        m.init_variational_param(a1, a2, b1, b2)
    elif alpha1 is not None:                       # :404-405 -- init_hyper_param alone; the variational state stays at Ones
        m.init_hyper_param(alpha1, alpha2, beta1, beta2)
    else:                                          # :410-411
        m.init_random()
    m.optimize()                                   # :421
    if reorder_factor:                             # :424-427
        m.order_factor()
    res = m.output()                               # :431-434
    res["seed"] = seed
    return res
