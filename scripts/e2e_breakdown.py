"""Where the end-to-end call spends its time outside the iterations (C4, host CSR in, host factors out)."""
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
rp, ci, vv = csr.to_host(); csr.free()
hrp, hci, hvv = [torch.from_numpy(a).pin_memory() for a in (rp, ci, vv)]
Xh = (hrp.numpy(), hci.numpy(), hvv.numpy(), p)
prob_S, prior_S = np.full((p, K), 0.5), np.full(p, 0.5)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = pcmf_b200.Solver(Xh, K, True, True, seed=1003, prob_S=prob_S, prior_S=prior_S, iter_max=5, iter_min=5, epsilon=0.0, monitor=2)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    conv = s.optimize()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ed = s.monitors(loglikelihood=False)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    o = s.get_output(want_prob_D=False)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    s.close()
    t5 = time.perf_counter()
    print("rep %d: create %.3f s | optimize(5) %.3f s | exp_dev %.3f s | get_output %.3f s | close %.3f s | total %.3f s"
          % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0))
