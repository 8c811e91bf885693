"""pcmf_b200: B200-native (sm_100a CUDA) variational-EM core of pCMF behind the reference's pCMF()/run_* interface.

The numerical path lives in libpcmf_b200.so (C ABI: include/pcmf_b200.h); there is no CPU fallback."""
from .api import (DeviceCSR, PcmfResult, Solver, csr_column_stats, generate_counts_device, getU, getV, pCMF,  # noqa: F401
                  prefilter, run_gap_factor, run_sparse_gap_factor, run_zi_gap_factor, run_zi_sparse_gap_factor)
from ._lib import PcmfError  # noqa: F401

__version__ = "0.1.0"
