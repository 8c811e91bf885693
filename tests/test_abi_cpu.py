"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/pcmf_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pcmf_b200
from pcmf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(name="pcmf_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(pcmf_[a-z_0-9]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_library_exports_every_header_symbol():
    L = _lib.lib()
    declared = header_functions()
    assert set(declared) == set(_lib.HEADER_SYMBOLS), (declared, _lib.HEADER_SYMBOLS)
    dev = header_functions("pcmf_b200_dev.h")
    assert set(dev) == set(_lib.DEV_HEADER_SYMBOLS), (dev, _lib.DEV_HEADER_SYMBOLS)
    for name in declared + dev:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.pcmf_version()


def test_public_header_carries_no_bench_or_test_plumbing():
    # the header a pCMF maintainer links against declares the boundary only: measurement hooks, the input
    # generator and the A/B switches live in pcmf_b200_dev.h
    pub = open(os.path.join(ROOT, "include", "pcmf_b200.h")).read()
    for word in ("pcmf_measure", "pcmf_generate", "kernel_variant", "pcmf_phase_times"):
        assert word not in pub, word


def test_struct_layouts_match_header_field_order():
    # field order of the ctypes mirrors == field order in the header (a silent mismatch would corrupt arguments)
    src = open(os.path.join(ROOT, "include", "pcmf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def fields(struct_name):
        end = src.index("} %s;" % struct_name)
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                out.append(re.findall(r"[A-Za-z_0-9]+", part)[-1])
        return out

    assert fields("pcmf_problem") == [f[0] for f in _lib.Problem._fields_]
    assert fields("pcmf_options") == [f[0] for f in _lib.Options._fields_]
    assert fields("pcmf_init") == [f[0] for f in _lib.Init._fields_]
    assert fields("pcmf_result") == [f[0] for f in _lib.Result._fields_]


def test_no_cpu_fallback_without_device():
    L = _lib.lib()
    if L.pcmf_device_count() > 0:
        pytest.skip("a GPU is present")
    X = np.random.default_rng(0).poisson(2.0, (20, 15))
    with pytest.raises(pcmf_b200.PcmfError) as e:
        pcmf_b200.run_gap_factor(X, 2, verbose=False, iter_max=3)
    assert e.value.code == _lib.ECUDA and "no CPU fallback" in str(e.value)


def test_argument_errors_carry_reference_messages():
    # wrapper_gap_factor.h:210-258 texts, raised before any device work
    X = np.random.default_rng(0).poisson(2.0, (20, 15))
    with pytest.raises(pcmf_b200.PcmfError, match="Input paramter K should be lower than min"):
        pcmf_b200.run_gap_factor(X, 16, verbose=False)
    with pytest.raises(pcmf_b200.PcmfError, match="sel_bound should be a real value between 0 and 1"):
        pcmf_b200.run_sparse_gap_factor(X, 2, sel_bound=1.5, verbose=False)
    with pytest.raises(pcmf_b200.PcmfError, match="conv_mode"):
        pcmf_b200.run_gap_factor(X, 2, conv_mode=3, verbose=False)
    with pytest.raises(pcmf_b200.PcmfError, match="'random' or 'nmf'"):
        pcmf_b200.run_gap_factor(X, 2, init_mode="foo", verbose=False)
    with pytest.raises(pcmf_b200.PcmfError, match="Wrong dimension"):
        pcmf_b200.run_gap_factor(X, 2, U=np.ones((3, 2)), V=np.ones((15, 2)), verbose=False)


def test_product_package_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under pcmf_b200/ may import, load or link it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pcmf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "pcmf_oracle" not in txt and "import oracle" not in txt, f


def test_multi_init_longer_than_iter_max_is_rejected():
    # ninit > 1 with iter_init > iter_max would write iter_init entries into iter_max-long caller vectors
    X = np.random.default_rng(0).poisson(2.0, (20, 15))
    with pytest.raises(pcmf_b200.PcmfError, match="iter_init") as e:
        pcmf_b200.run_gap_factor(X, 2, verbose=False, iter_max=50, ninit=2, iter_init=100)
    assert e.value.code == _lib.EINVAL


def test_hyper_parameter_forms_are_checked():
    X = np.random.default_rng(0).poisson(2.0, (20, 15))
    ones_n, ones_p = np.ones((20, 2)), np.ones((15, 2))
    with pytest.raises(pcmf_b200.PcmfError, match="alpha1, alpha2, beta1 and/or beta2"):
        pcmf_b200.run_gap_factor(X, 2, alpha1=ones_n, alpha2=ones_n, beta1=ones_p, beta2=np.ones((14, 2)), verbose=False)
    with pytest.raises(pcmf_b200.PcmfError, match="all alpha1, alpha2, beta1, beta2 or for none"):
        pcmf_b200.run_gap_factor(X, 2, alpha1=ones_n, verbose=False)
