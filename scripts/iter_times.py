"""Per-iteration wall time of the first iterations after a seeded random init (C4): where optimize(5) differs from 5 x steady state."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcmf_b200
import bench

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
n, p, K = cfg["n"], cfg["p"], cfg["K"]
U, V = bench.factor_matrices(cfg, 3)
csr = pcmf_b200.generate_counts_device(n, p, cfg["K_true"], U, V, prob1=cfg["prob1"], seed=3, device=0)
prob_S, prior_S = np.full((p, K), 0.5), np.full(p, 0.5)
s = pcmf_b200.Solver(csr, K, True, True, seed=1003, prob_S=prob_S, prior_S=prior_S, iter_max=20, iter_min=20, epsilon=0.0, monitor=2)
prev = {k: v[0] for k, v in s.phase_times().items()}
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.iterate(1, with_elbo=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ph = s.phase_times()
    d = {k: ph[k][0] - prev.get(k, 0.0) for k in ph}
    prev = {k: v[0] for k, v in ph.items()}
    fb = {k: ph[k][1] for k in ph if k.startswith("exact_fallback")}
    print("iter %d: %.1f ms | " % (it, 1e3 * (t1 - t0)) + ", ".join("%s %.1f" % (k, v) for k, v in d.items() if abs(v) > 0.5), "| fallback entries (cumulative)", fb)
torch.cuda.synchronize(); t0 = time.perf_counter()
e = s.elbo()
torch.cuda.synchronize(); print("final elbo pass %.1f ms" % (1e3 * (time.perf_counter() - t0)))
s.close()
