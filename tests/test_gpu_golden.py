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


@pytest.mark.parametrize("name", ["seeded_gap", "seeded_zi_sparse", "rv_zi", "factor_sparse"])
def test_cuda_path_matches_second_family_of_golden_vectors(name):
    """Seeded random init (one MT19937 stream, cells then genes), the RV-coefficient stopping rule and init_from_factor,
    each against outputs of the reference build."""
    from tests.test_golden_vectors import EXTRA, compare_extra, load_extra
    zi, sparse = EXTRA[name]
    X, K, kw, ref = load_extra(name)
    res = RUN[(zi, sparse)](X, K, verbose=False, monitor=True, **kw)
    compare_extra(res, ref, 1e-6, "cuda vs golden " + name)
