"""The CUDA path (through the C ABI) against the committed golden vectors of the reference itself
(tests/golden/ref_*.npz, see tests/golden/make_golden.py): full run_* call with monitors, stopping rule, factor ordering.
Tolerance 1e-6 relative as north_star states for FP64 mode (observed ~1e-12)."""
import pytest

import pcmf_b200
from tests.refcmp import compare_result
from tests.test_golden_vectors import CASES, load_case, run_kwargs

pytestmark = pytest.mark.gpu
RUN = {(False, False): pcmf_b200.run_gap_factor, (True, False): pcmf_b200.run_zi_gap_factor,
       (False, True): pcmf_b200.run_sparse_gap_factor, (True, True): pcmf_b200.run_zi_sparse_gap_factor}


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_matches_golden_reference_vectors(name):
    zi, sparse = CASES[name]
    inp, ref = load_case(name)
    kw = run_kwargs(inp, sparse)
    res = RUN[(zi, sparse)](inp["X"], inp["K"], verbose=False, monitor=True, **kw)
    compare_result(res, ref, sparse, zi, 1e-6, "cuda vs golden " + name)
