"""ctypes binding of libpcmf_b200.so (include/pcmf_b200.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpcmf_b200.so")

OK, EINVAL, ECUDA, ECOMM, EUNSUPPORTED, EINTERRUPT = 0, 1, 2, 3, 4, 5

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double)

dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class Problem(C.Structure):
    _fields_ = [("n_local", C.c_int64), ("n_global", C.c_int64), ("row_offset", C.c_int64), ("p", C.c_int32),
                ("dense_f64", C.c_void_p), ("dense_i32", C.c_void_p), ("csr_rowptr", C.c_void_p),
                ("csr_colidx", C.c_void_p), ("csr_val_f32", C.c_void_p), ("csr_val_f64", C.c_void_p),
                ("csr_on_device", C.c_int32),
                ("csc_colptr", C.c_void_p), ("csc_rowidx", C.c_void_p), ("csc_val_f64", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("K", C.c_int32), ("zero_inflation", C.c_int32), ("sparsity", C.c_int32), ("sel_bound", C.c_double),
                ("iter_max", C.c_int32), ("iter_min", C.c_int32), ("epsilon", C.c_double), ("additional_iter", C.c_int32),
                ("conv_mode", C.c_int32), ("monitor", C.c_int32), ("reorder_factor", C.c_int32), ("ninit", C.c_int32),
                ("iter_init", C.c_int32), ("has_seed", C.c_int32), ("seed", C.c_uint32), ("verbose", C.c_int32),
                ("device", C.c_int32), ("stream", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32),
                ("allreduce", ALLREDUCE_FN), ("allreduce_ctx", C.c_void_p), ("progress", PROGRESS_FN),
                ("progress_ctx", C.c_void_p), ("progress_every", C.c_int32), ("precision", C.c_int32), ("ngpus", C.c_int32),
                ("dev_flags", C.c_int32)]


class Init(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("U", "V", "a1", "a2", "b1", "b2", "alpha1", "alpha2", "beta1", "beta2",
                                          "prob_S", "prior_S", "prob_D", "prior_D")] + [("hyper_rows", C.c_int32)]


class Result(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("EU", "ElogU", "a1", "a2", "alpha1", "alpha2", "EV", "ElogV", "b1", "b2",
                                          "beta1", "beta2", "prob_S", "prior_prob_S", "S", "prob_D", "freq_D",
                                          "prior_prob_D", "factor_order", "conv_crit", "norm_gap", "abs_gap",
                                          "optim_criterion", "loglikelihood", "deviance")] + \
               [("nb_iter", C.c_int32), ("converged", C.c_int32), ("loss", C.c_double), ("exp_dev", C.c_double),
                ("seconds_setup", C.c_double), ("seconds_iter", C.c_double), ("seconds_total", C.c_double)]


class PcmfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("libpcmf_b200.so is not built (%s).  Run `python -m pcmf_b200.build` or "
                          "__graft_entry__.build(); this package has no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    L.pcmf_version.restype = C.c_char_p
    L.pcmf_device_count.restype = C.c_int
    err = [C.c_char_p, C.c_size_t]
    L.pcmf_fit.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(Init), C.POINTER(Result)] + err
    L.pcmf_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(Init), C.POINTER(C.c_void_p)] + err
    L.pcmf_optimize.argtypes = [C.c_void_p, C.POINTER(Result)] + err
    L.pcmf_iterate.argtypes = [C.c_void_p, C.c_int, dp, dp] + err
    L.pcmf_elbo.argtypes = [C.c_void_p, dp, dp] + err
    L.pcmf_order_factor.argtypes = [C.c_void_p] + err
    L.pcmf_multi_init.argtypes = [C.c_void_p, C.POINTER(Result)] + err
    L.pcmf_monitors.argtypes = [C.c_void_p, dp, dp, dp] + err
    L.pcmf_get_output.argtypes = [C.c_void_p, C.POINTER(Result)] + err
    L.pcmf_set_variational.argtypes = [C.c_void_p, dp, dp, dp, dp] + err
    L.pcmf_phase_times.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), dp, lp, C.c_int]
    L.pcmf_launch_count.argtypes = [C.c_void_p]
    L.pcmf_launch_count.restype = C.c_int64
    L.pcmf_destroy.argtypes = [C.c_void_p]
    L.pcmf_destroy.restype = None
    L.pcmf_generate_counts_device.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_int32, dp, dp, dp, C.c_uint64, C.c_int64,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                              lp] + err
    L.pcmf_csr_column_stats.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, dp, dp, dp, dp] + err
    L.pcmf_measure_fp64_tflops.argtypes = [C.c_int, dp]
    L.pcmf_measure_fp64_mma_tflops.argtypes = [C.c_int, dp]
    L.pcmf_measure_gather_gbs.argtypes = [C.c_int, C.c_int64, dp]
    L.pcmf_device_free.argtypes = [C.c_int, C.c_void_p]
    L.pcmf_device_to_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    for name in ("pcmf_fit", "pcmf_create", "pcmf_optimize", "pcmf_iterate", "pcmf_elbo", "pcmf_order_factor", "pcmf_multi_init",
                 "pcmf_monitors", "pcmf_get_output", "pcmf_set_variational", "pcmf_phase_times",
                 "pcmf_generate_counts_device", "pcmf_device_free", "pcmf_device_to_host", "pcmf_csr_column_stats",
                 "pcmf_measure_fp64_tflops", "pcmf_measure_fp64_mma_tflops", "pcmf_measure_gather_gbs"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, errbuf):
    if rc != OK:
        raise PcmfError(rc, errbuf.value.decode(errors="replace") or ("pcmf error %d" % rc))


# every symbol include/pcmf_b200.h declares (tests check the library exports them all) ...
HEADER_SYMBOLS = ["pcmf_version", "pcmf_device_count", "pcmf_fit", "pcmf_create", "pcmf_optimize", "pcmf_iterate",
                  "pcmf_elbo", "pcmf_order_factor", "pcmf_multi_init", "pcmf_monitors", "pcmf_get_output", "pcmf_set_variational",
                  "pcmf_destroy", "pcmf_csr_column_stats", "pcmf_release_cached_memory"]
# ... and the development / measurement entry points of include/pcmf_b200_dev.h
DEV_HEADER_SYMBOLS = ["pcmf_phase_times", "pcmf_launch_count", "pcmf_generate_counts_device", "pcmf_device_free",
                      "pcmf_device_to_host", "pcmf_measure_fp64_tflops", "pcmf_measure_fp64_mma_tflops",
                      "pcmf_measure_gather_gbs"]
