"""N>1 path on CPU: the nnz-balanced cell partitioner and, with a world_size-2 gloo group, the sharding algebra the
engine relies on -- which sums are all-reduced (gene-side E-step sums, prob_D^T E[U], column sums of E[U]/E[log U],
the cell halves of the parameter gap, column sums of prob_D) and which state is replicated -- checked against the
unsharded dense oracle over several iterations of all four models."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcmf_b200 import datagen, sharding
from tests.np_twin import Twin


def test_balanced_row_partition():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, 1000)
    counts[100:120] = 2000     # a few very deep cells
    rowptr = np.concatenate([[0], np.cumsum(counts)])
    for world in (1, 2, 3, 8):
        parts = sharding.balanced_row_partition(rowptr, world)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert all(r1 > r0 for r0, r1 in parts)
        nnz = [rowptr[r1] - rowptr[r0] for r0, r1 in parts]
        assert max(nnz) - min(nnz) <= 2 * 2000 + 50   # within a couple of (deep) rows of perfectly balanced
    with pytest.raises(ValueError):
        sharding.balanced_row_partition(np.arange(3), 4)
    eq = sharding.equal_row_partition(10, 3)
    assert eq == [(0, 4), (4, 7), (7, 10)]
    rp, ci, vv = sharding.shard_csr(rowptr, np.arange(rowptr[-1]), np.ones(rowptr[-1]), 5, 9)
    assert rp[0] == 0 and rp[-1] == counts[5:9].sum() and len(ci) == rp[-1]


class ShardedTwin(Twin):
    """The design twin on one row shard; cell sums go through dist.all_reduce exactly where the engine reduces."""

    def __init__(self, X, K, zi, sparse, n_global):
        super().__init__(X, K, zi, sparse)
        self.n_global = n_global

    @staticmethod
    def ar(x):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)).copy())
        dist.all_reduce(t)
        return t.numpy()

    # column means over ALL cells
    def hyper(self):
        from tests.np_twin import digamma_inv
        n, p = self.n_global, self.p
        self.alpha1 = digamma_inv(np.log(self.alpha2) + self.ar(self.ElogU.sum(0)) / n)
        self.alpha2 = self.alpha1 / (self.ar(self.EU.sum(0)) / n)
        self.beta1 = digamma_inv(np.log(self.beta2) + self.ElogV.mean(0))
        self.beta2 = self.beta1 / self.EV.mean(0)
        if self.sparse:
            self.prior_S = self.prob_S.mean(1)
        if self.zi:
            self.prior_D = self.ar(self.nz.sum(0)) / n if self.first else self.ar(self.colsumP) / n

    def iterate(self):
        from tests.np_twin import expit, logit
        X, nz = self.X, self.nz
        eu = np.exp(self.ElogU)
        ev = np.exp(self.ElogV) * self.S
        evs = ev * self.prob_S if self.sparse else ev
        T = eu @ ev.T
        Tp = np.where(T > 0, T, 1.0)
        P = self.P_formula() if self.zi else None
        w = X * P if self.zi else X
        Q = np.where(nz, w / Tp, 0.0)
        EZ_j = eu * (Q @ evs)
        B1 = self.ar(Q.T @ eu)                       # all-reduce #1: gene-side E-step sums
        EZ_i = evs * B1
        EVS = self.EVS()
        self.a1 = self.alpha1[None, :] + EZ_j
        self.a2 = self.alpha2[None, :] + (P @ EVS if self.zi else np.repeat(EVS.sum(0)[None, :], self.n, 0))
        self.stats_U()
        self.b1 = self.beta1[None, :] + EZ_i
        if self.zi:
            tmpMat = self.ar(P.T @ self.EU)          # all-reduce #1: prob_D^T EU_new
        else:
            tmpMat = np.repeat(self.ar(self.EU.sum(0))[None, :], self.p, 0)   # all-reduce #1: colsum(EU_new)
        self.b2 = self.beta2[None, :] + (self.prob_S * tmpMat if self.sparse else tmpMat)
        self.stats_V()
        if self.sparse:
            eu2 = np.exp(self.ElogU)
            ev2 = np.exp(self.ElogV) * self.S
            T2 = eu2 @ ev2.T
            T2 = np.where(T2 > 0, T2, 1.0)
            Q2 = np.where(nz, (X * P if self.zi else X) / T2, 0.0)
            A1 = self.ar(Q2.T @ eu2)                 # all-reduce #2: sparse-probability sums
            A2 = self.ar(Q2.T @ (eu2 * self.ElogU))
            tmp = -self.EV * tmpMat + ev2 * (A2 + self.ElogV * A1)
            new = expit(logit(self.prior_S)[:, None] + tmp)
            new = np.where(self.prior_S[:, None] == 1, 1.0, new)
            new = np.where(self.prior_S[:, None] == 0, 0.0, new)
            self.prob_S = new
            self.S = (self.prob_S >= self.sel_bound).astype(float)
        if self.zi:
            self.EUz, self.EVSz = self.EU.copy(), self.EVS().copy()
            self.pz = self.prior_D.copy()
            self.cz = logit(self.prior_D)
            self.first = False
            self.colsumP = self.P_formula().sum(0)   # all-reduce #3 happens inside hyper()
        self.hyper()

    def gap(self):
        dU = self.ar([((self.a1 - self.a1o) ** 2).sum() + ((self.a2 - self.a2o) ** 2).sum(),
                      (self.a1o ** 2).sum() + (self.a2o ** 2).sum()])     # rides on all-reduce #1
        d = dU[0] + ((self.b1 - self.b1o) ** 2).sum() + ((self.b2 - self.b2o) ** 2).sum()
        nrm = dU[1] + (self.b1o ** 2).sum() + (self.b2o ** 2).sum()
        return np.sqrt(d), np.sqrt(d) / np.sqrt(nrm)


def _worker(rank, world, port, X, K, zi, sparse, inits, ref_states, ref_gaps, errors):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        rowptr = np.concatenate([[0], np.cumsum((X != 0).sum(1))])
        r0, r1 = sharding.balanced_row_partition(rowptr, world)[rank]
        a1, a2, b1, b2, prob_S, prior_S = inits
        t = ShardedTwin(X[r0:r1], K, zi, sparse, X.shape[0])
        t.init(a1[r0:r1], a2[r0:r1], b1, b2, prob_S, prior_S)
        for it, (ref, rg) in enumerate(zip(ref_states, ref_gaps)):
            t.iterate()
            g = t.gap()
            for name in ("a1", "a2"):
                e = np.max(np.abs(getattr(t, name) - ref[name][r0:r1]) / np.abs(ref[name][r0:r1]))
                assert e < 1e-9, (it, name, e)
            for name in ("b1", "b2"):      # replicated state: identical on every rank, equal to the unsharded run
                e = np.max(np.abs(getattr(t, name) - ref[name]) / np.abs(ref[name]))
                assert e < 1e-9, (it, name, e)
            assert abs(g[1] - rg) < 1e-9 * rg
            if sparse:
                assert (t.S.astype(int) == ref["S"]).all()
            if zi:
                assert np.max(np.abs(t.prior_D - ref["prior_D"])) < 1e-10
            t.next()
        dist.destroy_process_group()
    except Exception as e:   # pragma: no cover
        errors.put((rank, repr(e)))
        raise


@pytest.mark.parametrize("zi,sparse", [(False, False), (True, False), (False, True), (True, True)])
def test_world2_gloo_sharding_matches_unsharded_oracle(zi, sparse):
    from oracle import pyoracle as po
    X, _, _ = datagen.simulate(41, 30, 4, signal_U=20.0, signal_V=20.0, prob1=0.6, seed=3)
    X = X.astype(float)
    K = 3
    a1, a2, b1, b2 = datagen.random_init(X, K, seed=11)
    prob_S, prior_S = np.full((30, K), 0.5), np.full(30, 0.5)
    m = po.Model(X, K, zi, sparse)
    if sparse:
        m.init_sparse_param(0.5, prob_S, prior_S)
    if zi:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    states, gaps = [], []
    for _ in range(4):
        m.update_param()
        states.append(m.state())
        gaps.append(m.gap_between_iterates()[1])
        m.prepare_next_iterate()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, X, K, zi, sparse, (a1, a2, b1, b2, prob_S, prior_S), states, gaps, errors))
             for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
    assert errors.empty(), errors.get()
    assert all(pr.exitcode == 0 for pr in procs)


def _prior_worker(rank, world, port, X, K, expect, errors):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pcmf_b200 import api
        r0, r1 = sharding.equal_row_partition(X.shape[0], world)[rank]
        prob_S, prior_S = api.sparse_prior(X[r0:r1], K, comm=api._Comm(), n_global=X.shape[0])
        assert np.max(np.abs(prior_S - expect)) < 1e-12, np.max(np.abs(prior_S - expect))
        assert prob_S.shape == (X.shape[1], K) and np.array_equal(prob_S[:, 0], prior_S)
        dist.destroy_process_group()
    except Exception as e:   # pragma: no cover
        errors.put((rank, repr(e)))
        raise


def test_world2_pcmf_prior_S_heuristic_is_global():
    """pCMF(X_shard, comm=...) must derive prob_S / prior_S (pCMF.R:228-229) from the statistics of ALL cells --
    gene-side state is replicated, a per-shard heuristic would give every rank a different model."""
    X, _, _ = datagen.simulate(60, 25, 4, signal_U=20.0, signal_V=20.0, prob1=0.6, seed=8)
    X = X.astype(float)
    _, expect = datagen.prior_S_heuristic(X, 3)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    procs = [ctx.Process(target=_prior_worker, args=(r, 2, port, X, 3, expect, errors)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
    assert errors.empty(), errors.get()
    assert all(pr.exitcode == 0 for pr in procs)
