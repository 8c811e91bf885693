"""The boundary seen from plain C: tests/cabi/cabi_client.c includes include/pcmf_b200.h, is compiled with gcc -std=c99 and
linked against libpcmf_b200.so -- no Python, no ctypes mirrors in between.  CPU part: the header is valid C, struct sizes
and offsets equal the ctypes mirrors, and without a device pcmf_fit returns PCMF_ECUDA with a message (no CPU fallback).
GPU part: one pcmf_fit call per model reproduces what the Python host obtains through the same library."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import pcmf_b200
from pcmf_b200 import _lib, datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pcmf_b200")


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    _lib.lib()                                              # makes sure the .so is built
    exe = str(tmp_path_factory.mktemp("cabi") / "cabi_client")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cabi", "cabi_client.c"), "-o", exe, "-L", LIBDIR, "-lpcmf_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def write_case(path, X, K, zi, sparse, iter_max, reorder, seed):
    n, p = X.shape
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=seed)
    prob_S, prior_S = datagen.prior_S_heuristic(X, K)
    with open(path, "wb") as f:
        f.write(struct.pack("7i", n, p, K, int(zi), int(sparse), iter_max, int(reorder)))
        for a in (X.astype(np.float64), a1, a2, b1, b2, prob_S, prior_S):
            f.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
    return dict(a1=a1, a2=a2, b1=b1, b2=b2, prob_S=prob_S, prior_S=prior_S)


def read_result(path, n, p, K, iter_max):
    raw = open(path, "rb").read()
    rc, nb_iter, converged = struct.unpack_from("3i", raw, 0)
    loss, exp_dev = struct.unpack_from("2d", raw, 12)
    off = 28
    out = dict(rc=rc, nb_iter=nb_iter, converged=converged, loss=loss, exp_dev=exp_dev)
    for name, shape in (("EU", (n, K)), ("EV", (p, K)), ("a1", (n, K)), ("a2", (n, K)), ("b1", (p, K)), ("b2", (p, K)),
                        ("prior_prob_S", (p,)), ("prior_prob_D", (p,)), ("optim_criterion", (iter_max,))):
        cnt = int(np.prod(shape))
        out[name] = np.frombuffer(raw, np.float64, cnt, off).reshape(shape, order="F")
        off += 8 * cnt
    out["factor_order"] = np.frombuffer(raw, np.int32, K, off)
    return out


def test_header_is_c99_and_struct_layouts_match_ctypes(client):
    r = subprocess.run([client, "sizes"], stdout=subprocess.PIPE, text=True, check=True)
    got = {}
    for line in r.stdout.splitlines():
        tok = line.split()
        if tok[0].startswith("pcmf_"):
            got[tok[0]] = dict(size=int(tok[1]), **{tok[i]: int(tok[i + 1]) for i in range(2, len(tok), 2)})
    for name, cls in (("pcmf_problem", _lib.Problem), ("pcmf_options", _lib.Options), ("pcmf_init", _lib.Init), ("pcmf_result", _lib.Result)):
        assert got[name].pop("size") == C.sizeof(cls), name
        for field, off in got[name].items():
            assert getattr(cls, field).offset == off, (name, field)
    assert "sm_100a" in r.stdout


def test_c_client_gets_ecuda_without_a_device(client, tmp_path):
    if _lib.lib().pcmf_device_count() > 0:
        pytest.skip("a GPU is present")
    X, _, _ = datagen.simulate(30, 20, 3, signal_U=20.0, signal_V=20.0, prob1=0.6, seed=1)
    write_case(str(tmp_path / "in.bin"), X, 2, False, False, 6, False, 1)
    r = subprocess.run([client, "fit", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 10 + _lib.ECUDA
    assert "no CPU fallback" in r.stderr
    assert read_result(str(tmp_path / "out.bin"), 30, 20, 2, 6)["rc"] == _lib.ECUDA


@pytest.mark.gpu
@pytest.mark.parametrize("zi,sparse", [(False, False), (True, False), (False, True), (True, True)], ids=["gap", "zi", "sparse", "zi_sparse"])
def test_c_client_fit_matches_python_host(client, tmp_path, zi, sparse):
    X, _, _ = datagen.simulate(110, 70, 4, signal_U=20.0, signal_V=20.0, prob1=0.6, seed=5)
    n, p, K, iter_max = 110, 70, 3, 14
    init = write_case(str(tmp_path / "in.bin"), X, K, zi, sparse, iter_max, True, 7)
    r = subprocess.run([client, "fit", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    got = read_result(str(tmp_path / "out.bin"), n, p, K, iter_max)
    run = {(False, False): pcmf_b200.run_gap_factor, (True, False): pcmf_b200.run_zi_gap_factor,
           (False, True): pcmf_b200.run_sparse_gap_factor, (True, True): pcmf_b200.run_zi_sparse_gap_factor}[(zi, sparse)]
    kw = dict(a1=init["a1"], a2=init["a2"], b1=init["b1"], b2=init["b2"])
    if sparse:
        kw.update(prob_S=init["prob_S"], prior_S=init["prior_S"])
    ref = run(X, K, verbose=False, monitor=True, iter_max=iter_max, iter_min=iter_max // 2, reorder_factor=True, **kw)
    assert got["nb_iter"] == ref["convergence"]["nb_iter"]
    assert list(got["factor_order"]) == list(ref["factor_order"])
    for k in ("a1", "a2", "b1", "b2"):
        assert np.max(np.abs(got[k] - ref["variational_params"][k]) / np.abs(ref["variational_params"][k])) < 1e-11, k
    for k in ("EU", "EV"):
        assert np.max(np.abs(got[k] - ref["stats"][k]) / np.abs(ref["stats"][k])) < 1e-11, k
    assert abs(got["loss"] - ref["loss"]) < 1e-11 * abs(ref["loss"])
    assert abs(got["exp_dev"] - ref["exp_dev"]) < 1e-10
    nb = got["nb_iter"]
    assert np.max(np.abs(got["optim_criterion"][:nb] - ref["monitor"]["optim_criterion"]) / np.abs(ref["monitor"]["optim_criterion"])) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("zi,sparse", [(False, False), (True, True)], ids=["gap", "zi_sparse"])
def test_c_client_two_gpus_inside_the_library(client, tmp_path, zi, sparse):
    """pcmf_options.ngpus = 2 from plain C: the library shards the cells over two devices itself (host threads + NCCL
    communicator of its own, multi.cu) -- no launcher, no callback.  Must equal the one-device fit (VERDICT r1 item 5;
    replaces wrapper_gap_factor.h:271-274 + the OpenMP reduction of gap_factor_model.cpp:474-478)."""
    if _lib.lib().pcmf_device_count() < 2:
        pytest.skip("needs two GPUs")
    X, _, _ = datagen.simulate(150, 70, 4, signal_U=20.0, signal_V=20.0, prob1=0.6, seed=6)
    n, p, K, iter_max = 150, 70, 3, 12
    write_case(str(tmp_path / "in.bin"), X, K, zi, sparse, iter_max, True, 7)
    out = {}
    for ng in (1, 2):
        env = dict(os.environ, PCMF_CABI_NGPUS=str(ng))
        r = subprocess.run([client, "fit", str(tmp_path / "in.bin"), str(tmp_path / ("out%d.bin" % ng))], stderr=subprocess.PIPE,
                           text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        out[ng] = read_result(str(tmp_path / ("out%d.bin" % ng)), n, p, K, iter_max)
    a, b = out[1], out[2]
    assert a["nb_iter"] == b["nb_iter"] and list(a["factor_order"]) == list(b["factor_order"])
    for k in ("a1", "a2", "b1", "b2", "EU", "EV"):
        assert np.max(np.abs(a[k] - b[k]) / np.abs(a[k])) < 1e-10, k
    assert abs(a["loss"] - b["loss"]) < 1e-10 * abs(a["loss"])
    nb = a["nb_iter"]
    assert np.max(np.abs(a["optim_criterion"][:nb] - b["optim_criterion"][:nb]) / np.abs(a["optim_criterion"][:nb])) < 1e-10


@pytest.mark.gpu
def test_python_host_ngpus_matches_single_device():
    """run_zi_sparse_gap_factor(..., ngpus = 2): scipy CSR input (nnz-balanced row ranges), seeded random init, monitors."""
    if _lib.lib().pcmf_device_count() < 2:
        pytest.skip("needs two GPUs")
    import scipy.sparse as sp
    X, _, _ = datagen.simulate(400, 120, 4, signal_U=20.0, signal_V=20.0, prob1=0.5, seed=9)
    Xs = sp.csr_matrix(X)
    kw = dict(verbose=False, monitor=True, iter_max=10, iter_min=10, reorder_factor=True, seed=77,
              prob_S=np.full((120, 3), 0.5), prior_S=np.full(120, 0.5))
    one = pcmf_b200.run_zi_sparse_gap_factor(Xs, 3, **kw)
    two = pcmf_b200.run_zi_sparse_gap_factor(Xs, 3, ngpus=2, **kw)
    for k in ("a1", "a2", "b1", "b2"):
        assert np.max(np.abs(one["variational_params"][k] - two["variational_params"][k]) / np.abs(one["variational_params"][k])) < 1e-10, k
    for k in ("optim_criterion", "loglikelihood", "deviance"):
        assert np.max(np.abs(one["monitor"][k] - two["monitor"][k]) / np.abs(one["monitor"][k])) < 1e-10, k
    assert np.abs(one["ZI_param"]["prob_D"] - two["ZI_param"]["prob_D"]).max() < 1e-10
    assert (one["sparse_param"]["S"] == two["sparse_param"]["S"]).all()
