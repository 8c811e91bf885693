"""Host-side mirror of the reference's R interface for the VEM path, on top of the C ABI (include/pcmf_b200.h).

R is not installed in this image, so this Python layer plays the role of pkg/R/pCMF.R + pkg/R/RcppExports.R: same
function names, argument names/order/defaults, error texts and return-object field names, so that tests written
against the reference read the same here.  (INTEGRATION.md shows the Rcpp shim a maintainer drops into the package.)

    pCMF(X, K, zero_inflation=True, sparsity=True, ...)            pkg/R/pCMF.R:197-259
    run_gap_factor / run_zi_gap_factor / run_sparse_gap_factor / run_zi_sparse_gap_factor
                                                                    pkg/R/RcppExports.R:212, 833, 595, 1085
    getU / getV                                                     pkg/R/utils.R:87-101, 171-188

Everything numerical happens in libpcmf_b200.so on the GPU; nothing here computes on the CPU beyond argument
marshalling and the `prior_S` heuristic of pCMF.R:228-229 (an O(n p) preprocessing step of the R wrapper).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Init, Options, PcmfError, Problem, Result

try:  # scipy is optional (CSR inputs)
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None


class PcmfResult(dict):
    """The reference's return list (algorithm.h:355-377 + model get_output) as a dict; `.cls` is the S3 class
    ("pCMF" or "spCMF", run_gap_factor.cpp:280 / run_sparse_gap_factor.cpp:310)."""
    cls = "pCMF"

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class DeviceCSR:
    """CSR count matrix resident in GPU memory (rows = local cells).  Owns its buffers."""

    def __init__(self, device, n, p, nnz, rowptr, colidx, val):
        self.device, self.n, self.p, self.nnz = device, int(n), int(p), int(nnz)
        self.rowptr, self.colidx, self.val = rowptr, colidx, val

    def free(self):
        L = _lib.lib()
        for ptr in (self.rowptr, self.colidx, self.val):
            if ptr:
                L.pcmf_device_free(self.device, ptr)
        self.rowptr = self.colidx = self.val = None

    def to_host(self):
        """-> (rowptr int64, colidx int32, val float32) numpy arrays (bench e2e leg)."""
        L = _lib.lib()
        rp = np.empty(self.n + 1, np.int64)
        ci = np.empty(self.nnz, np.int32)
        vv = np.empty(self.nnz, np.float32)
        for dst, src in ((rp, self.rowptr), (ci, self.colidx), (vv, self.val)):
            rc = L.pcmf_device_to_host(self.device, dst.ctypes.data_as(C.c_void_p), src, dst.nbytes)
            if rc:
                raise PcmfError(rc, "device->host copy failed")
        return rp, ci, vv

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def generate_counts_device(n, p, K_true, U, V, prob1=None, seed=0, device=0, row_offset=0):
    """X = Poisson(U V^T) * Bernoulli(prob1_j) (pkg/R/data_generation.R:393-463) generated on the GPU into CSR."""
    L = _lib.lib()
    U = np.asfortranarray(np.asarray(U, np.float64))
    V = np.asfortranarray(np.asarray(V, np.float64))
    assert U.shape == (n, K_true) and V.shape == (p, K_true)
    pr = None if prob1 is None else np.ascontiguousarray(np.broadcast_to(np.asarray(prob1, np.float64), (p,)))
    rp, ci, vv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz = C.c_int64(0)
    err = C.create_string_buffer(512)
    rc = L.pcmf_generate_counts_device(device, n, p, K_true, U.ctypes.data_as(_lib.dp), V.ctypes.data_as(_lib.dp),
                                       pr.ctypes.data_as(_lib.dp) if pr is not None else None, int(seed), int(row_offset),
                                       C.byref(rp), C.byref(ci), C.byref(vv), C.byref(nnz), err, 512)
    _lib.check(rc, err)
    return DeviceCSR(device, n, p, nnz.value, rp.value, ci.value, vv.value)


def csr_column_stats(X, device=0, with_max=False):
    """Per-gene (sum x, sum x^2, nnz[, max]) over the local cells of a DeviceCSR, a scipy sparse matrix or a host CSR
    tuple (rowptr, colidx, val, p), computed on the GPU."""
    L = _lib.lib()
    if isinstance(X, DeviceCSR):
        n, p, rp, ci, vv, on_dev, keep = X.n, X.p, X.rowptr, X.colidx, X.val, 1, None
        device = X.device
    else:
        if _sp is not None and _sp.issparse(X):
            Xc = X.tocsr()
            X = (Xc.indptr, Xc.indices, Xc.data, Xc.shape[1])
        rp_, ci_, vv_, p = X
        rp_ = np.ascontiguousarray(rp_, np.int64)
        ci_ = np.ascontiguousarray(ci_, np.int32)
        vv_ = np.ascontiguousarray(vv_, np.float32)
        keep = (rp_, ci_, vv_)
        n, on_dev = len(rp_) - 1, 0
        rp, ci, vv = (x.ctypes.data_as(C.c_void_p) for x in keep)
    s1, s2, cn, mx = np.zeros(p), np.zeros(p), np.zeros(p), np.zeros(p)
    err = C.create_string_buffer(512)
    rc = L.pcmf_csr_column_stats(device, n, p, rp, ci, vv, on_dev, s1.ctypes.data_as(_lib.dp), s2.ctypes.data_as(_lib.dp),
                                 cn.ctypes.data_as(_lib.dp), mx.ctypes.data_as(_lib.dp), err, 512)
    _lib.check(rc, err)
    return (s1, s2, cn, mx) if with_max else (s1, s2, cn)


def prefilter(X, prop=0.05, quant_max=1, presel=False, threshold=0.2, device=0):
    """pkg/R/prefiltering.R:110-146: boolean vector over the genes (columns) that are kept.

    crit1 max > 0; crit2 at least prop * n non-null entries; crit3 max <= quantile(max over genes, quant_max) (R's default
    type-7 quantile = numpy's linear interpolation); crit4 (presel) 1 - exp(-sd_j / mean(X[X != 0])) > threshold.
    A dense matrix is reduced with numpy the way the R function uses apply(); sparse / device-resident inputs (which the
    R function cannot take) get their per-gene statistics from one GPU pass over the non-zeros."""
    for name, v in (("prop", prop), ("quant_max", quant_max), ("threshold", threshold)):
        if not np.isscalar(v) or not (0 <= v <= 1):
            raise PcmfError(_lib.EINVAL, "`%s` should be a [0,1]-valued number" % name)
    if isinstance(X, DeviceCSR) or (_sp is not None and _sp.issparse(X)) or isinstance(X, tuple):
        s1, s2, cn, mx = csr_column_stats(X, device=device, with_max=True)
        n = X.n if isinstance(X, DeviceCSR) else (X.shape[0] if not isinstance(X, tuple) else len(X[0]) - 1)
        var = (s2 - s1 ** 2 / n) / (n - 1)
        sd = np.sqrt(np.maximum(var, 0.0))
        nzmean = s1.sum() / cn.sum()
    else:
        A = np.asarray(X)
        if A.ndim != 2:
            raise PcmfError(_lib.EINVAL, "`X` should be a matrix or a data.frame or an array")
        n = A.shape[0]
        mx, cn = A.max(axis=0), (A != 0).sum(axis=0)
        sd = A.std(axis=0, ddof=1) if presel else None
        nzmean = A[A != 0].mean() if presel else None
    kept = (mx > 0) & (cn >= prop * n) & (mx <= np.quantile(mx, quant_max))
    if presel:
        kept &= (1 - np.exp(-sd / nzmean)) > threshold
    return kept


def prior_S_from_stats(colsum, colsumsq, colnnz, n, K):
    """pCMF.R:228-229 from per-gene sums: prior_S = 1 - exp(-sd_j / mean(X[X != 0])), sd with the n-1 estimator."""
    var = (colsumsq - colsum ** 2 / n) / (n - 1)
    sd = np.sqrt(np.maximum(var, 0.0))
    prior_S = 1 - np.exp(-sd / (colsum.sum() / colnnz.sum()))
    return np.repeat(prior_S[:, None], K, axis=1), prior_S


def measure_fp64_tflops(device=0):
    t = C.c_double(0)
    rc = _lib.lib().pcmf_measure_fp64_tflops(device, C.byref(t))
    if rc:
        raise PcmfError(rc, "fp64 peak measurement failed")
    return t.value


def measure_fp64_mma_tflops(device=0):
    """FP64 tensor-pipe (DMMA) throughput in TFLOP/s, measured on the device."""
    v = C.c_double(0)
    rc = _lib.lib().pcmf_measure_fp64_mma_tflops(int(device), C.byref(v))
    if rc != _lib.OK:
        raise PcmfError(rc, "pcmf_measure_fp64_mma_tflops failed")
    return v.value


def measure_gather_gbs(rows, device=0):
    """Gather bandwidth (GB/s) of 160-byte rows out of a `rows`-row table: ceiling of the E-step non-zero passes."""
    v = C.c_double(0)
    rc = _lib.lib().pcmf_measure_gather_gbs(int(device), int(rows), C.byref(v))
    if rc != _lib.OK:
        raise PcmfError(rc, "pcmf_measure_gather_gbs failed")
    return v.value


class _Comm:
    """Sum all-reduce over torch.distributed for the cell-sharded run (one process per GPU).  The device buffer the
    library hands to the callback is wrapped zero-copy through __cuda_array_interface__."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.calls = 0
        self.bytes = 0

        def cb(ctx, devptr, count, stream):
            try:
                t = self._wrap(devptr, count)
                # order the collective on the library's stream: its producers were enqueued there, its consumers follow there
                s = torch.cuda.ExternalStream(int(stream)) if stream else torch.cuda.default_stream()
                with torch.cuda.stream(s):
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
                self.calls += 1
                self.bytes += count * 8
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                import sys
                print("pcmf allreduce callback failed: %r" % (e,), file=sys.stderr)
                return 1

        self.cfn = _lib.ALLREDUCE_FN(cb)

    def allreduce_host(self, arr):
        """Sum of a small host array over the ranks (set-up statistics, e.g. the per-gene sums behind pCMF()'s prior_S)."""
        torch = self.torch
        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.as_tensor(np.ascontiguousarray(arr, np.float64)).to(dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _wrap(self, devptr, count):
        torch = self.torch

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(devptr), False), "version": 2}
        return torch.as_tensor(h, device="cuda")


def _as_problem(X, n_global=None, row_offset=0):
    """-> (Problem, keepalive list, n, p)"""
    keep = []
    pr = Problem()
    if isinstance(X, DeviceCSR):
        n, p = X.n, X.p
        pr.csr_rowptr, pr.csr_colidx, pr.csr_val_f32, pr.csr_on_device = X.rowptr, X.colidx, X.val, 1
        keep.append(X)
    elif _sp is not None and _sp.issparse(X) and X.format == "csc" and X.nnz < 2 ** 31 and X.shape[0] < 2 ** 31 - 1:
        # compressed columns = Matrix::dgCMatrix (slots p, i, x): handed over as it is, CSR is built on the device
        if not X.has_canonical_format:           # a dgCMatrix never holds the same entry twice; scipy may
            X = X.copy()
            X.sum_duplicates()
        n, p = X.shape
        cp = np.ascontiguousarray(X.indptr, np.int32)
        ri = np.ascontiguousarray(X.indices, np.int32)
        xv = np.ascontiguousarray(X.data, np.float64)
        pr.csc_colptr, pr.csc_rowidx, pr.csc_val_f64 = (a.ctypes.data_as(C.c_void_p) for a in (cp, ri, xv))
        keep += [cp, ri, xv]
    elif _sp is not None and _sp.issparse(X):
        Xc = X.tocsr()
        if not Xc.has_canonical_format or (Xc.nnz and (Xc.data == 0).any()):   # one entry per (cell, gene); explicit
            Xc = Xc.copy()                                                      # zeros must not count as X != 0
            Xc.sum_duplicates()
            Xc.eliminate_zeros()
        n, p = Xc.shape
        rp = np.ascontiguousarray(Xc.indptr, np.int64)
        ci = np.ascontiguousarray(Xc.indices, np.int32)
        data = np.asarray(Xc.data)
        f32 = data.astype(np.float32)
        if np.array_equal(f32.astype(np.float64), data.astype(np.float64)):
            vv = np.ascontiguousarray(f32)
            pr.csr_val_f32 = vv.ctypes.data_as(C.c_void_p)
        else:
            vv = np.ascontiguousarray(data, np.float64)
            pr.csr_val_f64 = vv.ctypes.data_as(C.c_void_p)
        pr.csr_rowptr, pr.csr_colidx = rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p)
        keep += [rp, ci, vv]
    elif isinstance(X, tuple) and len(X) == 4:
        # (rowptr, colidx, val, p) host CSR
        rp, ci, vv, p = X
        rp = np.ascontiguousarray(rp, np.int64)
        ci = np.ascontiguousarray(ci, np.int32)
        n = len(rp) - 1
        if vv.dtype == np.float32:
            vv = np.ascontiguousarray(vv)
            pr.csr_val_f32 = vv.ctypes.data_as(C.c_void_p)
        else:
            vv = np.ascontiguousarray(vv, np.float64)
            pr.csr_val_f64 = vv.ctypes.data_as(C.c_void_p)
        pr.csr_rowptr, pr.csr_colidx = rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p)
        keep += [rp, ci, vv]
    else:
        A = np.asarray(X)
        if A.ndim != 2:
            raise PcmfError(_lib.EINVAL, "Input matrix X should be integer or real valued.")
        n, p = A.shape
        if np.issubdtype(A.dtype, np.integer) and A.dtype.itemsize <= 4:
            Af = np.asfortranarray(A, dtype=np.int32)          # INTSXP (wrapper_gap_factor.h:204-206)
            pr.dense_i32 = Af.ctypes.data_as(C.c_void_p)
        elif np.issubdtype(A.dtype, np.integer) or np.issubdtype(A.dtype, np.floating) or A.dtype == bool:
            Af = np.asfortranarray(A, dtype=np.float64)        # REALSXP (:207-208)
            pr.dense_f64 = Af.ctypes.data_as(C.c_void_p)
        else:
            raise PcmfError(_lib.EINVAL, "Input matrix X should be integer or real valued.")
        keep.append(Af)
    pr.n_local, pr.p = n, p
    pr.n_global = n if n_global is None else int(n_global)
    pr.row_offset = int(row_offset)
    return pr, keep, n, p


def _fcol(a, shape, name):
    a = np.asfortranarray(np.asarray(a, np.float64))
    if a.shape != tuple(shape):
        raise PcmfError(_lib.EINVAL, "Wrong dimension for %s input matrices" % name)
    return a


class Solver:
    """Step-wise handle on the GPU engine (pcmf_create / pcmf_iterate / pcmf_optimize / ...)."""

    def __init__(self, X, K, zero_inflation=False, sparsity=False, sel_bound=0.5, iter_max=1000, iter_min=500,
                 epsilon=1e-2, additional_iter=10, conv_mode=1, monitor=True, reorder_factor=False, ninit=1, iter_init=100,
                 seed=None, verbose=False, U=None, V=None, a1=None, a2=None, b1=None, b2=None, alpha1=None, alpha2=None,
                 beta1=None, beta2=None, prob_S=None, prior_S=None, prob_D=None, prior_D=None, device=0, stream=None,
                 comm=None, n_global=None, row_offset=0, progress=None, kernel_variant=0, precision=0, ngpus=0):
        self.L = _lib.lib()
        self.h = None
        pr, self._keep, n, p = _as_problem(X, n_global, row_offset)
        self.n, self.p, self.K = n, p, int(K)
        self.zi, self.sparse = bool(zero_inflation), bool(sparsity)
        self.iter_max = int(iter_max)
        o = Options()
        o.K, o.zero_inflation, o.sparsity, o.sel_bound = int(K), int(self.zi), int(self.sparse), float(sel_bound)
        o.iter_max, o.iter_min, o.epsilon, o.additional_iter = int(iter_max), int(iter_min), float(epsilon), int(additional_iter)
        o.conv_mode, o.monitor, o.reorder_factor, o.ninit, o.iter_init = int(conv_mode), int(monitor), int(bool(reorder_factor)), int(ninit), int(iter_init)
        o.has_seed, o.seed = (0, 0) if seed is None else (1, int(seed) & 0xFFFFFFFF)
        o.verbose, o.device = int(bool(verbose)), int(device)
        o.stream = stream
        o.dev_flags = int(kernel_variant)
        o.precision = {0: 0, 1: 1, "fp64": 0, "fp32": 1}[precision]
        o.ngpus = int(ngpus)
        self.comm = comm
        if comm is not None:
            o.rank, o.world, o.allreduce = comm.rank, comm.world, comm.cfn
        else:
            o.rank, o.world = 0, 1
        if progress is not None:
            self._pcb = _lib.PROGRESS_FN(lambda ctx, it, crit: int(progress(it, crit) or 0))
            o.progress = self._pcb
        o.progress_every = 10
        self.opt = o
        ini = Init()
        k = self._keep

        def put(field, arr, shape, name):
            if arr is None:
                return
            a = _fcol(arr, shape, name)
            k.append(a)
            setattr(ini, field, a.ctypes.data_as(C.c_void_p))

        Kk = self.K
        put("U", U, (n, Kk), "U and/or V")
        put("V", V, (p, Kk), "U and/or V")
        for f, a, r in (("a1", a1, n), ("a2", a2, n), ("b1", b1, p), ("b2", b2, p)):
            put(f, a, (r, Kk), "a1, a2, b1 and/or b2")
        # hyper-parameters: the reference's n x K / p x K matrices as given (gap_factor_model.cpp:126-132), or K-vectors
        hyp = [(f, np.asarray(a, np.float64), r) for f, a, r in (("alpha1", alpha1, n), ("alpha2", alpha2, n),
                                                                  ("beta1", beta1, p), ("beta2", beta2, p)) if a is not None]
        if hyp:
            as_rows = hyp[0][1].ndim == 2
            for f, a, r in hyp:
                if (a.ndim == 2) != as_rows or (as_rows and a.shape != (r, Kk)):
                    raise PcmfError(_lib.EINVAL, "Wrong dimension for alpha1, alpha2, beta1 and/or beta2 input matrices")
                put(f, a, (r, Kk) if as_rows else (Kk,), "alpha1, alpha2, beta1 and/or beta2")
            ini.hyper_rows = int(as_rows)
        if self.sparse and prob_S is not None and prior_S is not None:
            put("prob_S", prob_S, (p, Kk), "prob_S and/or prior_S")
            pv = np.ascontiguousarray(np.asarray(prior_S, np.float64).reshape(-1))
            if pv.shape != (p,):
                raise PcmfError(_lib.EINVAL, "Wrong dimension for prob_S and/or prior_S input matrix/vector")
            k.append(pv)
            ini.prior_S = pv.ctypes.data_as(C.c_void_p)
        if self.zi and prob_D is not None and prior_D is not None:      # wrapper_gap_factor.h:352-356
            put("prob_D", prob_D, (n, p), "prob_D and/or prior_D")
            pd = np.ascontiguousarray(np.asarray(prior_D, np.float64).reshape(-1))
            if pd.shape != (p,):
                raise PcmfError(_lib.EINVAL, "Wrong dimension for prob_D and/or prior_D input matrix/vector")
            k.append(pd)
            ini.prior_D = pd.ctypes.data_as(C.c_void_p)
        self._pr, self._ini = pr, ini
        if o.ngpus > 1:
            return                      # pcmf_fit drives several devices in one call (Solver.fit); no step-wise handle
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.L.pcmf_create(C.byref(pr), C.byref(o), C.byref(ini), C.byref(h), err, 512)
        _lib.check(rc, err)
        self.h = h

    def fit(self, want_prob_D=True):
        """One pcmf_fit call: wrapper_*gap_factor end to end (init -> [multi-init] -> optimize -> order_factor -> output).
        -> (convergence dict, exp_dev, output arrays)"""
        r, bufs = self._result_buffers(want_prob_D, vectors=True)
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_fit(C.byref(self._pr), C.byref(self.opt), C.byref(self._ini), C.byref(r), err, 512), err)
        return self._conv_dict(r, bufs), r.exp_dev, bufs

    def close(self):
        if self.h:
            self.L.pcmf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def iterate(self, niter=1, with_elbo=False):
        """niter x (update_param, check_convergence, prepare_next_iterate); returns conv_crit[, elbo] per iteration."""
        crit = np.zeros(max(niter, 1))
        elbo = np.zeros(max(niter, 1))
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_iterate(self.h, int(niter), crit.ctypes.data_as(_lib.dp),
                                       elbo.ctypes.data_as(_lib.dp) if with_elbo else None, err, 512), err)
        return (crit[:niter], elbo[:niter]) if with_elbo else crit[:niter]

    def elbo(self, terms=False):
        e = C.c_double(0)
        t = np.zeros(11)
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_elbo(self.h, C.byref(e), t.ctypes.data_as(_lib.dp), err, 512), err)
        if terms:
            names = ["gammaU", "entU", "gammaV", "entV", "bernS_prior", "bernS_self", "bernD_prior", "bernD_self",
                     "sumUVt", "data", "lgammaX"]
            return e.value, dict(zip(names, t))
        return e.value

    def set_variational(self, a1, a2, b1, b2):
        arrs = [_fcol(a1, (self.n, self.K), "a1"), _fcol(a2, (self.n, self.K), "a2"),
                _fcol(b1, (self.p, self.K), "b1"), _fcol(b2, (self.p, self.K), "b2")]
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_set_variational(self.h, *[a.ctypes.data_as(_lib.dp) for a in arrs], err, 512), err)

    def _result_buffers(self, want_prob_D=True, vectors=True):
        n, p, K = self.n, self.p, self.K
        bufs = {}
        r = Result()

        def mk(name, shape, dtype=np.float64):
            a = np.zeros(shape, dtype=dtype, order="F")
            bufs[name] = a
            setattr(r, name, a.ctypes.data_as(C.c_void_p))

        for nm in ("EU", "ElogU", "a1", "a2", "alpha1", "alpha2"):
            mk(nm, (n, K))
        for nm in ("EV", "ElogV", "b1", "b2", "beta1", "beta2"):
            mk(nm, (p, K))
        if self.sparse:
            mk("prob_S", (p, K))
            mk("prior_prob_S", (p,))
            mk("S", (p, K), np.int32)
        if self.zi:
            mk("freq_D", (p,))
            mk("prior_prob_D", (p,))
            if want_prob_D:
                mk("prob_D", (n, p))
        mk("factor_order", (K,), np.int32)
        if vectors:
            for nm in ("conv_crit", "norm_gap", "abs_gap", "optim_criterion", "loglikelihood", "deviance"):
                mk(nm, (max(self.iter_max, 1),))
        return r, bufs

    def get_output(self, want_prob_D=True):
        r, bufs = self._result_buffers(want_prob_D, vectors=False)
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_get_output(self.h, C.byref(r), err, 512), err)
        return bufs

    def _conv_dict(self, r, bufs):
        nb = r.nb_iter
        out = {"nb_iter": int(nb), "converged": bool(r.converged), "loss": r.loss, "seconds_iter": r.seconds_iter}
        for nm in ("conv_crit", "norm_gap", "abs_gap", "optim_criterion", "loglikelihood", "deviance"):
            out[nm] = bufs[nm][:nb].copy()
        return out

    def optimize(self, multi_init=False):
        """algorithm::optimize; with multi_init=True the multi-init loop of the wrapper runs first (the per-iteration
        vectors then start with the kept run's first iter_init entries)."""
        r, bufs = self._result_buffers(False, vectors=True)
        err = C.create_string_buffer(512)
        if multi_init:
            _lib.check(self.L.pcmf_multi_init(self.h, C.byref(r), err, 512), err)
        _lib.check(self.L.pcmf_optimize(self.h, C.byref(r), err, 512), err)
        return self._conv_dict(r, bufs)

    def order_factor(self):
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_order_factor(self.h, err, 512), err)

    def monitors(self, loglikelihood=True):
        a, b, c = C.c_double(float("nan")), C.c_double(0), C.c_double(0)
        err = C.create_string_buffer(512)
        _lib.check(self.L.pcmf_monitors(self.h, C.byref(a) if loglikelihood else None, C.byref(b), C.byref(c), err, 512), err)
        return {"loglikelihood": a.value, "deviance": b.value, "exp_dev": c.value}

    def phase_times(self):
        names = (C.c_char_p * 24)()
        ms = np.zeros(24)
        ln = np.zeros(24, np.int64)
        c = self.L.pcmf_phase_times(self.h, names, ms.ctypes.data_as(_lib.dp), ln.ctypes.data_as(_lib.lp), 24)
        return {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(c)}

    def launch_count(self):
        return int(self.L.pcmf_launch_count(self.h))


def _run(X, K, zero_inflation, sparsity, sel_bound=0.5, U=None, V=None, verbose=True, monitor=True, iter_max=1000,
         iter_min=500, init_mode="random", epsilon=1e-2, additional_iter=10, conv_mode=1, ninit=1, iter_init=100,
         ncores=1, reorder_factor=True, seed=None, a1=None, a2=None, b1=None, b2=None, alpha1=None, alpha2=None,
         beta1=None, beta2=None, prob_S=None, prior_S=None, prob_D=None, prior_D=None, return_prob_D=True, **gpu):
    """wrapper_gap_factor<model, variational_em_algo> / wrapper_sparse_gap_factor<...> through pcmf_fit's engine."""
    if init_mode == "nmf":
        raise PcmfError(_lib.EINVAL, "FixMe not implemented yet")               # wrapper_gap_factor.h:370
    if init_mode != "random":
        raise PcmfError(_lib.EINVAL, "`init_mode` input parameter should be 'random' or 'nmf'")   # :372
    del ncores  # kept for signature compatibility (omp_set_num_threads in the reference); the GPU path ignores it
    progress = None
    if verbose:
        progress = lambda it, crit: print("iter %d" % it) or 0   # algorithm.h:170-173
    s = Solver(X, K, zero_inflation, sparsity, sel_bound=sel_bound, iter_max=iter_max, iter_min=iter_min, epsilon=epsilon,
               additional_iter=additional_iter, conv_mode=conv_mode, monitor=monitor, reorder_factor=reorder_factor,
               ninit=ninit, iter_init=iter_init, seed=seed, verbose=verbose, U=U, V=V, a1=a1, a2=a2, b1=b1, b2=b2,
               alpha1=alpha1, alpha2=alpha2, beta1=beta1, beta2=beta2, prob_S=prob_S, prior_S=prior_S, prob_D=prob_D,
               prior_D=prior_D, progress=progress, **gpu)
    try:
        if s.opt.ngpus > 1:
            # the library shards the cells over the devices itself: ONE pcmf_fit call, whole matrices in and out (multi.cu)
            conv, exp_dev, o = s.fit(want_prob_D=return_prob_D)
        else:
            conv = s.optimize(multi_init=bool(ninit and ninit > 1))     # wrapper_gap_factor.h:296-350, :379
            if reorder_factor:
                s.order_factor()                                        # :382-385
            exp_dev = s.monitors(loglikelihood=False)["exp_dev"]        # algorithm.h:367
            o = s.get_output(want_prob_D=return_prob_D)
    finally:
        s.close()
    res = PcmfResult()
    res.cls = "spCMF" if sparsity else "pCMF"
    res["factor"] = {"U": o["EU"], "V": o["EV"]}                                           # gap...cpp:370-372
    res["variational_params"] = {k: o[k] for k in ("a1", "a2", "b1", "b2")}                # :374-377
    res["hyper_params"] = {k: o[k] for k in ("alpha1", "alpha2", "beta1", "beta2")}        # :379-382
    res["stats"] = {k: o[k] for k in ("EU", "EV", "ElogU", "ElogV")}                       # :387-390
    if sparsity:
        res["sparse_param"] = {"prob_S": o["prob_S"], "prior_prob_S": o["prior_prob_S"], "S": o["S"]}   # sparse...cpp:280-283
    if zero_inflation:
        res["ZI_param"] = {"prob_D": o.get("prob_D"), "freq_D": o["freq_D"], "prior_prob_D": o["prior_prob_D"]}   # zi...cpp:250-253
    res["convergence"] = {"converged": conv["converged"], "nb_iter": conv["nb_iter"], "conv_crit": conv["conv_crit"],
                          "conv_mode": int(conv_mode)}                                     # algorithm.h:359-363
    res["loss"] = conv["loss"]
    res["exp_dev"] = exp_dev
    if monitor:
        res["monitor"] = {k: conv[k] for k in ("norm_gap", "abs_gap", "loglikelihood", "deviance",
                                               "optim_criterion")}                         # algorithm.h:369-376
    res["seed"] = seed                                                                     # wrapper_gap_factor.h:391
    res["factor_order"] = o["factor_order"]
    return res


def run_gap_factor(X, K, U=None, V=None, verbose=True, monitor=True, iter_max=1000, iter_min=500, init_mode="random",
                   epsilon=1e-2, additional_iter=10, conv_mode=1, ninit=1, iter_init=100, ncores=1, reorder_factor=True,
                   seed=None, a1=None, a2=None, b1=None, b2=None, alpha1=None, alpha2=None, beta1=None, beta2=None, **gpu):
    """pkg/R/RcppExports.R:212 / pkg/src/run_gap_factor.cpp:252-282"""
    return _run(X, K, False, False, 0.5, U, V, verbose, monitor, iter_max, iter_min, init_mode, epsilon, additional_iter,
                conv_mode, ninit, iter_init, ncores, reorder_factor, seed, a1, a2, b1, b2, alpha1, alpha2, beta1, beta2, **gpu)


def run_zi_gap_factor(X, K, U=None, V=None, verbose=True, monitor=True, iter_max=1000, iter_min=500, init_mode="random",
                      epsilon=1e-2, additional_iter=10, conv_mode=1, ninit=1, iter_init=100, ncores=1, reorder_factor=True,
                      seed=None, a1=None, a2=None, b1=None, b2=None, alpha1=None, alpha2=None, beta1=None, beta2=None,
                      prob_D=None, prior_D=None, **gpu):
    """pkg/R/RcppExports.R:833 / pkg/src/run_zi_gap_factor.cpp:278-310"""
    return _run(X, K, True, False, 0.5, U, V, verbose, monitor, iter_max, iter_min, init_mode, epsilon, additional_iter,
                conv_mode, ninit, iter_init, ncores, reorder_factor, seed, a1, a2, b1, b2, alpha1, alpha2, beta1, beta2,
                prob_D=prob_D, prior_D=prior_D, **gpu)


def run_sparse_gap_factor(X, K, sel_bound=0.5, U=None, V=None, verbose=True, monitor=True, iter_max=1000, iter_min=500,
                          init_mode="random", epsilon=1e-2, additional_iter=10, conv_mode=1, ninit=1, iter_init=100,
                          ncores=1, reorder_factor=True, seed=None, a1=None, a2=None, b1=None, b2=None, alpha1=None,
                          alpha2=None, beta1=None, beta2=None, prob_S=None, prior_S=None, **gpu):
    """pkg/R/RcppExports.R:595 / pkg/src/run_sparse_gap_factor.cpp:279-312"""
    return _run(X, K, False, True, sel_bound, U, V, verbose, monitor, iter_max, iter_min, init_mode, epsilon,
                additional_iter, conv_mode, ninit, iter_init, ncores, reorder_factor, seed, a1, a2, b1, b2, alpha1, alpha2,
                beta1, beta2, prob_S=prob_S, prior_S=prior_S, **gpu)


def run_zi_sparse_gap_factor(X, K, sel_bound=0.5, U=None, V=None, verbose=True, monitor=True, iter_max=1000, iter_min=500,
                             init_mode="random", epsilon=1e-2, additional_iter=10, conv_mode=1, ninit=1, iter_init=100,
                             ncores=1, reorder_factor=True, seed=None, a1=None, a2=None, b1=None, b2=None, alpha1=None,
                             alpha2=None, beta1=None, beta2=None, prob_S=None, prior_S=None, prob_D=None, prior_D=None,
                             **gpu):
    """pkg/R/RcppExports.R:1085 / pkg/src/run_zi_sparse_gap_factor.cpp:292-327"""
    return _run(X, K, True, True, sel_bound, U, V, verbose, monitor, iter_max, iter_min, init_mode, epsilon,
                additional_iter, conv_mode, ninit, iter_init, ncores, reorder_factor, seed, a1, a2, b1, b2, alpha1, alpha2,
                beta1, beta2, prob_S=prob_S, prior_S=prior_S, prob_D=prob_D, prior_D=prior_D, **gpu)


def sparse_prior(X, K, comm=None, device=0, n_global=None):
    """prior_S <- 1-exp(-apply(X,2,sd)/mean(X[X!=0])); prob_S <- matrix(rep(prior_S, K), ncol=K)   (pCMF.R:228-229)"""
    from .datagen import prior_S_heuristic
    sparse_in = isinstance(X, (DeviceCSR, tuple)) or (_sp is not None and _sp.issparse(X))
    if sparse_in or comm is not None:
        # X never exists as a dense matrix, or this process only holds a shard of the cells: the same heuristic from
        # per-gene sums (one GPU pass over the non-zeros), summed over the ranks so that every rank starts from the
        # SAME prob_S / prior_S -- gene-side state is replicated, a per-shard heuristic would silently fork the model
        if sparse_in:
            s1, s2, cn = csr_column_stats(X, device=device)
            n_loc = X.n if isinstance(X, DeviceCSR) else (len(X[0]) - 1 if isinstance(X, tuple) else X.shape[0])
        else:
            A = np.asarray(X, np.float64)
            s1, s2, cn, n_loc = A.sum(0), (A * A).sum(0), (A != 0).sum(0).astype(np.float64), A.shape[0]
        n_tot = n_loc
        if comm is not None:
            red = comm.allreduce_host(np.concatenate([s1, s2, cn, [float(n_loc)]]))
            p_ = len(s1)
            s1, s2, cn, n_tot = red[:p_], red[p_:2 * p_], red[2 * p_:3 * p_], int(round(red[-1]))
            if n_global not in (None, n_tot):
                raise PcmfError(_lib.EINVAL, "n_global does not equal the number of cells summed over the ranks")
        return prior_S_from_stats(s1, s2, cn, n_tot, K)
    return prior_S_heuristic(np.asarray(X), K)


def pCMF(X, K, zero_inflation=True, sparsity=True, sel_bound=0.5, verbose=True, monitor=True, iter_max=1000,
         iter_min=None, ninit=1, iter_init=100, ncores=1, reorder_factor=True, seed=None, **gpu):
    """pkg/R/pCMF.R:197-259, argument for argument.  Extra keyword arguments (device=, comm=, return_prob_D=, ...) are
    GPU plumbing that the R signature does not have; they are all optional."""
    if iter_min is None:
        iter_min = int(iter_max / 2)                                   # pCMF.R:203-205
    common = dict(verbose=verbose, monitor=monitor, iter_max=iter_max, iter_min=iter_min, init_mode="random",
                  epsilon=1e-2, additional_iter=10, conv_mode=1, ninit=ninit, iter_init=iter_init, ncores=ncores,
                  reorder_factor=reorder_factor, seed=seed)
    common.update(gpu)
    if not zero_inflation and not sparsity:
        return run_gap_factor(X, K, **common)
    if zero_inflation and not sparsity:
        return run_zi_gap_factor(X, K, **common)
    prob_S, prior_S = sparse_prior(X, K, comm=gpu.get("comm"), device=gpu.get("device", 0), n_global=gpu.get("n_global"))
    if not zero_inflation:
        return run_sparse_gap_factor(X, K, sel_bound=sel_bound, prob_S=prob_S, prior_S=prior_S, **common)
    return run_zi_sparse_gap_factor(X, K, sel_bound=sel_bound, prob_S=prob_S, prior_S=prior_S, **common)


def getU(model, log_representation=True):
    """pkg/R/utils.R:87-101"""
    EU = np.asarray(model["stats"]["EU"])
    return np.log(EU + 1e-5) if log_representation else EU


def getV(model, log_representation=True):
    """pkg/R/utils.R:171-188"""
    EV = np.asarray(model["stats"]["EV"])
    V = np.log(EV + 1) if log_representation else EV
    if getattr(model, "cls", "pCMF") == "spCMF":
        V = V * model["sparse_param"]["prob_S"]
    return V
