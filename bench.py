#!/usr/bin/env python
"""bench.py -- VEM iterations/sec of the pCMF hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W           (N>1 under torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W     (the CPU path of the reference, see below)

A "step" is one VEM iteration of the zero-inflated + sparse-V model: E-step, M-step, spike-and-slab update, ZI update,
hyper-parameter update, normalised-gap check and ELBO, on the configuration BASELINE.json quotes the
metric on: 1M cells x 20k genes, ~5% non-zeros, K=20, synthetic counts from the package's own generative model
(pkg/R/data_generation.R), generated on the GPU.  For N>1 the cells are sharded over the ranks (fixed total problem:
strong scaling) and the p x K gene-side statistics are all-reduced over NCCL every iteration.

The reference arm: the reference's own implementation cannot be built here (needs R + Rcpp + RcppEigen + BH, none of
which exist in the image, no network), so `--impl reference` times the CPU oracle -- the dense, OpenMP, loop-for-loop
restatement of the reference's C++ (oracle/pcmf_oracle.c) -- on all host cores on a bounded row sample of the same
workload and extrapolates linearly in n (the reference's cost is linear in n at fixed p, K).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # Synthetic inputs follow the package's generative model (data_generation.R): X = Poisson(U V^T) * Bernoulli(prob1_j).
    # `mean_rate` is the mean Poisson rate before drop-out; density ~= prob1 * P(Poisson > 0).  The sparse-V runs pass
    # prob_S = prior_S = 0.5 explicitly, as the reference's own tests do (test-gap_factor_model.R:641-690): with the
    # pCMF.R:228 heuristic ~99% of the genes have prior_S < sel_bound at 5% density and are masked out from iteration 0,
    # and with mean counts ~1 the reference's spike-and-slab update zeroes every loading, which would leave the E-step
    # with nothing to do.
    "c4": dict(n=1_000_000, p=20_000, K_true=20, K=20, prob1=0.05, mean_rate=20.0, zi=True, sparse=True,
               label="C4: 1M cells x 20k genes, ~5% nnz, ZI + sparse-V, K=20"),
    "c3": dict(n=50_000, p=5_000, K_true=20, K=20, prob1=0.3, mean_rate=20.0, zi=True, sparse=True,
               label="C3: 50k x 5k, ~30% nnz, ZI + sparse-V, K=20"),
    "c2": dict(n=10_000, p=2_000, K_true=10, K=10, prob1=None, mean_rate=20.0, zi=False, sparse=False,
               label="C2: 10k x 2k dense Gamma-Poisson, K=10"),
    "c4gap": dict(n=1_000_000, p=20_000, K_true=20, K=20, prob1=0.05, mean_rate=20.0, zi=False, sparse=False,
                  label="C4 data, plain Gamma-Poisson (no ZI, no sparsity), K=20"),
    **{"c5k%d" % k: dict(n=200_000, p=10_000, K_true=20, K=k, prob1=0.05, mean_rate=20.0, zi=True, sparse=False,
                         label="C5: 200k x 10k, ~5%% nnz, ZI, K=%d (latent-dimension sweep)" % k) for k in (2, 10, 30, 64)},
    "c4s8": dict(n=125_000, p=20_000, K_true=20, K=20, prob1=0.05, mean_rate=20.0, zi=True, sparse=True,
                 label="one 8-GPU shard of C4 on one GPU: 125k cells x 20k genes (kernel tuning only)"),
    "tiny": dict(n=20_000, p=2_000, K_true=20, K=20, prob1=0.05, mean_rate=20.0, zi=True, sparse=True,
                 label="tiny: 20k x 2k, ~5% nnz, ZI + sparse-V, K=20 (smoke only)"),
}


def factor_matrices(cfg, seed):
    """U (n x K_true, 3 cell groups), V (p x K_true, 2 gene groups, 60% noise genes): SURVEY 8d / README.md:119-137.
    U is rescaled so that the mean Poisson rate mean_ij (U V^T)_ij equals cfg['mean_rate']."""
    from pcmf_b200 import datagen
    rng = np.random.Generator(np.random.Philox(seed))
    n, p, Kt = cfg["n"], cfg["p"], cfg["K_true"]
    U = datagen.generate_factor_matrix(n, Kt, ngroup=3, average_signal=60.0, group_separation=0.8,
                                       distribution="exponential", shuffle_feature=True, rng=rng)["factor_matrix"]
    V = datagen.generate_factor_matrix(p, Kt, ngroup=2, average_signal=60.0, group_separation=0.8,
                                       distribution="exponential", shuffle_feature=True, prop_noise_feature=0.6,
                                       rng=rng)["factor_matrix"]
    mean_lam = float((U.mean(0) * V.mean(0)).sum())
    U = U * (cfg["mean_rate"] / mean_lam)
    return U, V


def shard_rows(n, rank, world):
    """contiguous, equal-count row ranges (the synthetic cells are exchangeable, so equal rows ~ equal nnz)"""
    base, rem = divmod(n, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would cripple a CPU baseline)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(cfg, U, V, seed, max_rows):
    """First `max_rows` cells of the workload as a dense host matrix (the reference's input format), plus the same seeded
    initial values the GPU arm uses."""
    from pcmf_b200 import datagen
    n, p, K = cfg["n"], cfg["p"], cfg["K"]
    ns = min(max_rows, n)
    rng = np.random.Generator(np.random.Philox(seed + 77))
    lam = U[:ns] @ V.T
    X = rng.poisson(lam).astype(np.float64)
    if cfg["prob1"] is not None:
        X *= rng.uniform(size=X.shape) < cfg["prob1"]
    X[X.sum(1) == 0, 0] = 1
    X[0, X.sum(0) == 0] = 1
    return X, datagen.random_init(X, K, seed=seed + 1000)


def cpu_baseline(cfg, U, V, seed, max_rows, iters, threads=None, warmup=0, sparse_aware=False, sample=None):
    """The oracle (dense OpenMP restatement of the reference) on the first `max_rows` cells of the same workload."""
    from oracle import pyoracle as po
    n, p, K = cfg["n"], cfg["p"], cfg["K"]
    X, (a1, a2, b1, b2) = sample if sample is not None else cpu_sample(cfg, U, V, seed, max_rows)
    ns = X.shape[0]
    po.lib().orc_set_num_threads(int(threads or host_threads()))
    cores = po.lib().orc_num_threads()
    m = po.Model(X, K, cfg["zi"], cfg["sparse"])
    if cfg["sparse"]:
        m.init_sparse_param(0.5, np.full((p, K), 0.5), np.full(p, 0.5))
    if cfg["zi"]:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    def timed(model):
        times = []
        for it in range(warmup + iters):
            t0 = time.perf_counter()
            model.update_param()
            model.elbo()
            model.gap_between_iterates()
            model.prepare_next_iterate()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        return float(np.mean(times))

    t_iter = timed(m)
    full = (1.0 / t_iter) * (ns / n)
    measured = ns == n
    out = {"value": full, "unit": "iter/s", "cores": int(cores), "kind": "port",
           "sample": "oracle (dense OpenMP restatement of the reference C++) on %d of %d cells, same p=%d K=%d, %d timed "
                     "iterations of update_param+elbo+gap, %.3f s/iter on the sample%s"
                     % (ns, n, p, K, iters, t_iter, "; the whole matrix, nothing extrapolated" if measured else
                        "; extrapolated linearly in n (x %d/%d)" % (ns, n)),
           "sample_rows": int(ns), "extrapolated": not measured, "sample_seconds_per_iter": t_iter}
    # Second figure, for a fairer comparison: the same loops made sparse-aware -- an entry with X = 0, whose
    # multinomial weights are multiplied by 0 wherever they are used, skips the K exponentials (orc_set_sparse_aware; the
    # dense zero-inflation products stay).  Continues from the state the first model reached.
    if sparse_aware:
        m.set_sparse_aware(True)
        t_sa = timed(m)
        out["sparse_aware"] = {"value": (1.0 / t_sa) * (ns / n), "unit": "iter/s", "cores": int(cores), "sample_seconds_per_iter": t_sa,
                               "note": "same oracle, zero entries skip the K exponentials of the multinomial split; same sample, "
                                       "same extrapolation"}
    return out


def ref_build_baseline(cfg, sample, iters):
    """The reference's OWN model and algorithm sources (oracle/_ref: compiled unmodified against stand-in Rcpp / Eigen / Boost
    headers, oracle/README_ref.md) on the same sample: algorithm::optimize for `iters` iterations with monitor = FALSE, i.e.
    update_param + check_convergence per iteration -- no ELBO, a lighter iteration than the one the other arms time."""
    from oracle import pyref
    if not pyref.available():
        return None
    n, p, K = cfg["n"], cfg["p"], cfg["K"]
    X, (a1, a2, b1, b2) = sample
    ns = X.shape[0]
    m = pyref.RefModel(X, K, cfg["zi"], cfg["sparse"], monitor=False, iter_max=iters, iter_min=iters, epsilon=1e-12)
    if cfg["sparse"]:
        m.init_sparse_param(0.5, np.full((p, K), 0.5), np.full(p, 0.5))
    if cfg["zi"]:
        m.init_zi_param()
    m.init_variational_param(a1, a2, b1, b2)
    t0 = time.perf_counter()
    m.optimize()
    t_iter = (time.perf_counter() - t0) / iters
    return {"value": (1.0 / t_iter) * (ns / n), "unit": "iter/s", "cores": int(os.environ.get("OMP_NUM_THREADS", "1")),
            "kind": "reference", "sample_rows": int(ns), "extrapolated": ns != n, "sample_seconds_per_iter": t_iter,
            "sample": "the reference's own gap_factor_model / algorithm sources (oracle/_ref build, stand-in Eigen underneath) on "
                      "%d of %d cells, %d iterations of optimize() with monitor = FALSE (no ELBO), %.3f s/iter" % (ns, n, iters, t_iter)}


def workload_config(cfg, precision="fp64", nnz=None):
    """The `config` object both arms print: same keys, so that `same_config` is decidable."""
    c = {"workload": cfg["label"], "precision": precision, "n": cfg["n"], "p": cfg["p"], "K": cfg["K"],
         "model": ("zi_" if cfg["zi"] else "") + ("sparse_" if cfg["sparse"] else "") + "gap"}
    if nnz is not None:
        c["nnz"] = nnz
        c["density"] = nnz / (cfg["n"] * cfg["p"])
    return c


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())      # before any OpenMP runtime is loaded (torchrun sets it to 1)
    U, V = factor_matrices(cfg, args.seed)
    # K timed + W warm-up steps, each on the bounded cell sample; rows are reduced if the run would exceed a few minutes
    iters, warm = max(1, args.steps), max(0, args.warmup)
    rows = args.cpu_rows if args.cpu_rows > 0 else cfg["n"]
    if (iters + warm) > 20:
        rows = max(200, int(rows * 20 / (iters + warm)))
    sample = cpu_sample(cfg, U, V, args.seed, rows)
    cb = cpu_baseline(cfg, U, V, args.seed, rows, iters, warmup=warm, sample=sample)
    try:
        rb = ref_build_baseline(cfg, sample, max(2, min(iters, 3)))
    except Exception as e:   # the checker's build is optional on a box without oracle/_ref
        rb = {"unavailable": repr(e)}
    config = workload_config(cfg)
    config.update({"sample_rows": cb["sample_rows"], "extrapolated": cb["extrapolated"]})
    line = {"impl": "reference", "metric": "VEM iterations/sec", "value": cb["value"], "unit": "iter/s", "n_gpus": args.gpus,
            "steps": iters, "warmup": warm, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": cb,
            "reference_build": rb,
            "e2e": {"value": cb["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's R package cannot be built here (no R/Rcpp/Eigen/Boost). Its own sources do compile against "
                    "stand-in headers (oracle/_ref, used to pin parity at 1e-10); that build is timed beside the port "
                    "(`reference_build`), but its linear algebra is not Eigen's, so the C port of the same loops -- the faster "
                    "of the two -- is the line's value (the conservative baseline); all host cores, dense n x p loops exactly "
                    "as the reference runs them"}
    emit(line)


def ncu_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of this
    workload (profiles/r02_ncu_traffic.json, written by profiles/ncu_traffic.py); {} when there is none."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(config, {})
    except Exception:
        return {}


_JSON_FD = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries print there too (NCCL's "NCCL version ..." banner at
    communicator creation): everything written to fd 1 from now on goes to stderr, the JSON line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=1500, help="cells of the CPU sample (0: the whole matrix -- c2 / c3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffers) leg")
    ap.add_argument("--variant", type=int, default=0, help="experimental kernel variant bits")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp64 (default, the headline: results within 1e-6 of the reference) or the optional FP32 mode")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    import pcmf_b200
    from pcmf_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the pcmf_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = api._Comm()
    args.warmup = max(args.warmup, 3)
    n, p, K = cfg["n"], cfg["p"], cfg["K"]
    r0, r1 = shard_rows(n, rank, world)
    nl = r1 - r0

    # ---- synthetic input, generated on the device straight into CSR (package's generative model)
    t_gen = time.perf_counter()
    U, V = factor_matrices(cfg, args.seed)
    csr = pcmf_b200.generate_counts_device(nl, p, cfg["K_true"], U[r0:r1], V, prob1=cfg["prob1"], seed=args.seed,
                                           device=local, row_offset=r0)
    t_gen = time.perf_counter() - t_gen
    nnz_local = csr.nnz
    nnz_t = torch.tensor([float(nnz_local)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())
    # prior_S heuristic of the R wrapper (pCMF.R:228-229) from per-gene sums
    prob_S = prior_S = None
    if cfg["sparse"]:
        prob_S, prior_S = np.full((p, K), 0.5), np.full(p, 0.5)   # test-gap_factor_model.R:641-690

    stream = torch.cuda.current_stream()
    solver = pcmf_b200.Solver(csr, K, cfg["zi"], cfg["sparse"], seed=args.seed + 1000, prob_S=prob_S, prior_S=prior_S,
                              device=local, stream=stream.cuda_stream, comm=comm, n_global=n, row_offset=r0,
                              iter_max=10 ** 6, kernel_variant=args.variant, precision=args.precision)
    # ---- warm-up, then exactly `steps` timed iterations (ELBO evaluated every iteration, as monitor=TRUE does)
    solver.iterate(args.warmup, with_elbo=True)
    ph0, l0 = solver.phase_times(), solver.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    crit, elbo = solver.iterate(args.steps, with_elbo=True)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clk = clocks.stop() if rank == 0 else None
    ph1, l1 = solver.phase_times(), solver.launch_count()
    launches = l1 - l0
    value = args.steps / (ms * 1e-3)

    # ---- roofline of the E-step kernels, from the per-phase CUDA-event times taken inside the timed region
    hbm_peak, peak_src = measured_peaks()
    stored_prob_D_bytes = int(ph1.pop("stored_prob_D_bytes", (0.0, 0))[1])
    ph0.pop("stored_prob_D_bytes", None)
    cell_blocks = int(ph1.pop("gene_pass_cell_blocks", (0.0, 1))[1])
    ph0.pop("gene_pass_cell_blocks", None)
    tiled = bool(int(ph1.pop("tiled_nonzero_passes", (0.0, 0))[1]))
    ph0.pop("tiled_nonzero_passes", None)
    per = {k: ((ph1[k][0] - ph0[k][0]) / max(1, args.steps), (ph1[k][1] - ph0[k][1]) // max(1, args.steps)) for k in ph1}
    w = 8
    zi, sparse = cfg["zi"], cfg["sparse"]
    # algorithmic bytes per launch, this rank
    b_cells = 8 * nnz_local + 8 * (nl + 1) + w * K * nl * (1 + 2 + 5 + (1 if zi else 0)) + w * K * p * (2 if (zi or sparse) else 1)
    b_genes = 8 * nnz_local + 8 * (p + 1) + w * K * nl * (2 if sparse else 1) + w * K * p * (1 + (2 if sparse else 1))
    # The non-zero passes read one or two 8K-byte factor rows per non-zero through L2 (gene table p x K for the cell
    # pass, cell table n x K for the gene passes); that gather, not the HBM stream of X, is what bounds them, so the
    # measured gather bandwidth of this device is reported next to the HBM figure.
    gather_gene_tab = api.measure_gather_gbs(p, local)
    # the gene passes walk the cells in blocks whose slice of the cell table stays in L2 (28 MB of eu + ElogU rows, solver.cu):
    # their ceiling is the gather rate out of a table of that size, or of the whole table when it is smaller than a block
    KPAD = next(w for w in (4, 8, 12, 16, 20, 32, 64) if K <= w)
    blk_rows = max(4096, (28 << 20) // (16 * KPAD))
    gather_cell_tab = api.measure_gather_gbs(min(nl, 2 * blk_rows), local)
    traffic = ncu_traffic(args.config)
    n_sms = torch.cuda.get_device_properties(local).multi_processor_count
    sm_mhz_max = float((clk or {}).get("sm_max_mhz") or 1965.0)
    kernels = []
    # gathered bytes per non-zero: the cell pass of the ZI / sparse models reads one K-vector (evw) plus the stored q of the
    # entry (a 32-byte sector) and its 4-byte CSR->CSC position; the plain model reads ev and recomputes T; the gene passes
    # read eu (+ ElogU when the spike-and-slab sums are wanted)
    stored_q = zi or sparse
    nz_rows = {"estep_cells_mstep_u": (b_cells, (8.0 * K + 36.0) if stored_q else 8.0 * K, gather_gene_tab),
               "estep_genes": (b_genes, None if sparse else 8.0 * K, gather_cell_tab)}
    if sparse:
        nz_rows["sparse_gene_pass"] = (b_genes, 16.0 * K, gather_cell_tab)
    for name, (byts, per_nnz, gpk) in nz_rows.items():
        t = per[name][0]
        row = {"kernel": name, "ms": t, "share": t * args.steps / ms if ms else None, "bound": "hbm",
               "algorithmic_bytes": byts, "achieved": byts / (t * 1e-3) / 1e9 if t > 0 else None,
               "peak": hbm_peak, "unit": "GB/s", "frac": (byts / (t * 1e-3) / 1e9 / hbm_peak) if t > 0 else None,
               "traffic": traffic.get(name),
               "note": "steady state: genes whose mask row did not change copy the sums of the previous spike-and-slab "
                       "pass; the CSC segments of the others are walked" if name == "estep_genes" and sparse else None}
        if tiled and not (name == "estep_genes" and sparse):
            # 2-D tiled copy of X (xtile.cuh): the factor rows of both sides of a tile sit in shared memory, so what bounds
            # the pass is the shared-memory pipe -- one 8K-byte row per non-zero (two in the spike-and-slab pass: eu and
            # eu o ElogU), read as 64-byte pieces by groups of 4 lanes.  Peak: 128 B / clk / SM.
            rows_per_nnz = 2 if name == "sparse_gene_pass" else 1
            sb = 8.0 * KPAD * rows_per_nnz * nnz_local
            speak = 128.0 * n_sms * sm_mhz_max * 1e6 / 1e9
            row.update({"layout": "tiled", "shared_bytes": sb, "shared_achieved": sb / (t * 1e-3) / 1e9 if t > 0 else None,
                        "shared_peak": speak, "shared_frac": (sb / (t * 1e-3) / 1e9 / speak) if t > 0 else None})
        else:
            gb = per_nnz * nnz_local if per_nnz else None
            row.update({"layout": "csr/csc", "gathered_bytes": gb, "gather_achieved": gb / (t * 1e-3) / 1e9 if (gb and t > 0) else None,
                        "gather_peak_measured": gpk, "gather_frac": (gb / (t * 1e-3) / 1e9 / gpk) if (gb and t > 0) else None})
        kernels.append(row)
    fp64_peak = api.measure_fp64_tflops(local)
    dmma_peak = api.measure_fp64_mma_tflops(local)
    if zi:
        # dense ZI passes: 2 contractions of length K per entry (4K flop) + expit, on the FP64 tensor pipe (DMMA)
        # When it fits in HBM the cell pass keeps prob_D (4 bytes per entry) and the gene pass streams it back: that pass is
        # then the second contraction only (2K flop per entry) over n p 4 bytes of HBM reads.
        stored_bytes = stored_prob_D_bytes
        # FP64 mode, K <= 32 (zi_i8.cuh): the second contraction of both passes runs as EXACT integer arithmetic on tcgen05
        # (u8 x u8 -> s32; prob_D as its stored 32-bit fixed-point word, the factor matrix in six base-256 digits).  The
        # cell-side pass keeps lambda = EU . EVS on the FP64 tensor pipe (DMMA) and the expit on the FP64 pipe, so its roofline
        # stays the DMMA peak over the 4 K flop per entry of the two contractions; the gene-side pass on the stored prob_D
        # has no FP64 work left per entry and is the HBM stream of prob_D (4 bytes per entry) plus the digit records.
        int_passes = args.precision == "fp64" and KPAD <= 32 and not (args.variant & (4 | 8 | 8192))
        for name in ("zi_pass_cells", "zi_pass_genes"):
            t = per[name][0]
            stored = stored_bytes > 0 and name == "zi_pass_genes"
            fl = (2.0 if stored else 4.0) * K * nl * p
            if stored and int_passes:
                digit_bytes = ((nl + 7) // 8) * ((6 * ((KPAD + 7) // 8 * 8) + 15) // 16 * 16) * 32
                byts = stored_bytes + digit_bytes + 8.0 * KPAD * nl + 8.0 * KPAD * p
                row = {"kernel": name, "ms": t, "share": t * args.steps / ms, "bound": "hbm", "algorithmic_bytes": byts,
                       "algorithmic_flops": fl, "achieved": byts / (t * 1e-3) / 1e9 if t > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                       "frac": (byts / (t * 1e-3) / 1e9 / hbm_peak) if t > 0 else None, "traffic": traffic.get(name),
                       "note": "integer tcgen05 product on the stored prob_D: bulk copies HBM -> shared memory -> UTCIMMA; the phase "
                               "time includes the digit packing of E[U]; MEASURED_PEAKS' HBM figure is a copy (read + write), "
                               "this pass only reads"}
            else:
                row = {"kernel": name, "ms": t, "share": t * args.steps / ms, "bound": "tensor", "algorithmic_flops": fl,
                       "achieved": fl / (t * 1e-3) / 1e12 if t > 0 else None, "peak": dmma_peak, "unit": "TFLOP/s",
                       "frac": (fl / (t * 1e-3) / 1e12 / dmma_peak) if t > 0 else None, "traffic": traffic.get(name),
                       "peak_source": "FP64 DMMA m8n8k4 microbenchmark run by this bench (MEASURED_PEAKS.json holds no FP64 "
                                      "figure; FP64 has no tcgen05 kind)"}
                if int_passes and name == "zi_pass_cells":
                    row["note"] = ("first contraction (2 K flop per entry) as DMMA + FP64 expit; second contraction as exact u8 MMAs on "
                                   "tcgen05 (UTCIMMA), not on the FP64 pipe; frac = all 4 K flop per entry over the DMMA peak")
            if stored_bytes > 0:
                row["stored_prob_D_bytes"] = stored_bytes
                row["hbm_stream_gbs"] = stored_bytes / (t * 1e-3) / 1e9 if t > 0 else None   # written by the cell pass, read by the gene pass
                row["hbm_peak"] = hbm_peak
            kernels.append(row)
    dom = max(kernels, key=lambda k: k["ms"])
    roofline = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"],
                "traffic": dom.get("traffic"), "kernel": dom["kernel"], "ms_per_launch": dom["ms"],
                "peak_source": dom.get("peak_source", peak_src)}
    phases = {k: {"ms_per_step": v[0], "launches_per_step": v[1]} for k, v in per.items()}
    if not args.no_e2e:
        solver.get_output(want_prob_D=False)     # untimed: creates the process-wide pinned staging buffers of the result download
    solver.close()

    # ---- end-to-end leg: the reference-facing call (run_zi_sparse_gap_factor -> pcmf C ABI) with HOST buffers:
    # H2D of the CSR matrix, preprocessing, `steps` iterations with the ELBO, D2H of the factors, all inside the timed region
    e2e = e2e_defaults = None
    if not args.no_e2e:
        rp, ci, vv = csr.to_host()
        csr.free()
        hrp = torch.from_numpy(rp).pin_memory()
        hci = torch.from_numpy(ci).pin_memory()
        hvv = torch.from_numpy(vv).pin_memory()
        fn = {(False, False): pcmf_b200.run_gap_factor, (True, False): pcmf_b200.run_zi_gap_factor,
              (False, True): pcmf_b200.run_sparse_gap_factor, (True, True): pcmf_b200.run_zi_sparse_gap_factor}[(zi, sparse)]
        h2d = hrp.numel() * 8 + hci.numel() * 4 + hvv.numel() * 4 + (p * K + p) * 8 * (1 if sparse else 0)
        d2h = 8 * (6 * nl * K + 6 * p * K) + (8 * (p * K + p) + 4 * p * K if sparse else 0) + (16 * p if zi else 0)

        def whole_call(Xh, monitor, reorder):
            kw = dict(verbose=False, monitor=monitor, iter_max=args.steps, iter_min=args.steps, epsilon=0.0, reorder_factor=reorder,
                      seed=args.seed + 1000, device=local, comm=comm, n_global=n, row_offset=r0, return_prob_D=False,
                      precision=args.precision)
            if sparse:
                kw.update(prob_S=prob_S, prior_S=prior_S)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = fn(Xh, K, **kw)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), res

        t_e2e, res = whole_call((hrp.numpy(), hci.numpy(), hvv.numpy(), p), 2, False)
        e2e = {"value": args.steps / t_e2e, "unit": "iter/s", "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "seconds_total": t_e2e, "nb_iter": res["convergence"]["nb_iter"],
               "note": "whole run_* call from pinned host CSR: H2D + CSC/bitmap/tile preprocessing + init + %d iterations with ELBO "
                       "+ D2H of the return object; copies and preprocessing are amortised over the %d steps" % (args.steps, args.steps)}
        # the same call with the arguments pCMF() passes by default -- monitor = TRUE (ELBO, log-likelihood and deviance
        # every iteration), reorder_factor = TRUE -- from ordinary (pageable) host arrays
        del hrp, hci, hvv
        t_def, res = whole_call((rp, ci, vv, p), 1, True)
        e2e_defaults = {"value": args.steps / t_def, "unit": "iter/s", "seconds_total": t_def, "nb_iter": res["convergence"]["nb_iter"],
                        "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": (d2h + 8 * 3 * args.steps) / args.steps,
                        "note": "monitor = TRUE, reorder_factor = TRUE (the defaults of pCMF(), pCMF.R:197-201), pageable host CSR"}
        del rp, ci, vv
    else:
        csr.free()

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:      # the CPU baseline is reported at N = 1 only
        cb = cpu_baseline(cfg, U, V, args.seed, args.cpu_rows if args.cpu_rows > 0 else n, 2, sparse_aware=True)

    if rank == 0:
        line = {"metric": "VEM iterations/sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
                "config": {**workload_config(cfg, args.precision, nnz),
                           "parallelism": "cells sharded over %d GPU(s), all-reduce of p x K gene statistics" % world,
                           "l2": "inputs larger than L2 (CSR+CSC %.1f GB per rank); no explicit flush" % (16e-9 * nnz_local),
                           "generator_seconds": t_gen},
                "roofline": roofline, "kernels": kernels, "phases": phases, "cpu_baseline": cb, "e2e": e2e, "e2e_defaults": e2e_defaults,
                "gpu_launches": int(launches), "clocks": clk, "fp64_peak_tflops_measured": fp64_peak, "fp64_dmma_peak_tflops_measured": dmma_peak,
                "elbo_last": float(elbo[-1]), "conv_crit_last": float(crit[-1])}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
