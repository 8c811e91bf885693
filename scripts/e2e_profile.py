"""cProfile of the reference-facing call (run_zi_sparse_gap_factor from pinned host CSR, C4): where the time outside the
library goes."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcmf_b200
import bench

cfg = bench.CONFIGS["c4"]
n, p, K = cfg["n"], cfg["p"], cfg["K"]
U, V = bench.factor_matrices(cfg, 3)
csr = pcmf_b200.generate_counts_device(n, p, cfg["K_true"], U, V, prob1=cfg["prob1"], seed=3, device=0)
rp, ci, vv = csr.to_host(); csr.free()
Xh = tuple(torch.from_numpy(a).pin_memory().numpy() for a in (rp, ci, vv)) + (p,)
prob_S, prior_S = np.full((p, K), 0.5), np.full(p, 0.5)
kw = dict(verbose=False, monitor=2, iter_max=5, iter_min=5, epsilon=0.0, reorder_factor=False, seed=1003, return_prob_D=False,
          prob_S=prob_S, prior_S=prior_S)
for rep in range(2):
    t0 = time.perf_counter(); res = pcmf_b200.run_zi_sparse_gap_factor(Xh, K, **kw); torch.cuda.synchronize()
    print("call %d: %.3f s" % (rep, time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
res = pcmf_b200.run_zi_sparse_gap_factor(Xh, K, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
