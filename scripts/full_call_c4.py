"""The full default call shape (monitor = TRUE, reorder_factor = TRUE) at C4 scale: timing of the parts around the iterations."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcmf_b200
import bench

cfg = bench.CONFIGS["c4"]
n, p, K = cfg["n"], cfg["p"], cfg["K"]
U, V = bench.factor_matrices(cfg, 3)
csr = pcmf_b200.generate_counts_device(n, p, cfg["K_true"], U, V, prob1=cfg["prob1"], seed=3, device=0)
prob_S, prior_S = np.full((p, K), 0.5), np.full(p, 0.5)
s = pcmf_b200.Solver(csr, K, True, True, seed=1003, prob_S=prob_S, prior_S=prior_S, iter_max=4, iter_min=4, epsilon=0.0, monitor=1)
torch.cuda.synchronize(); t0 = time.perf_counter()
conv = s.optimize()
torch.cuda.synchronize(); t1 = time.perf_counter()
print("optimize: 4 iterations with ELBO + log-likelihood + deviance every iteration: %.3f s (%.1f ms/iter)" % (t1 - t0, 250 * (t1 - t0)))
print("  loglikelihood", conv["loglikelihood"], "\n  deviance", conv["deviance"], "\n  elbo", conv["optim_criterion"])
s.order_factor()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("order_factor (K = 20 greedy rounds + reorder): %.3f s" % (t2 - t1))
m = s.monitors()
torch.cuda.synchronize(); t3 = time.perf_counter()
print("monitors: %.3f s" % (t3 - t2), m)
o = s.get_output(want_prob_D=False)
print("factor order", list(o["factor_order"]), "exact rounds", s.phase_times().get("order_factor_exact_rounds"))
s.close(); csr.free()
