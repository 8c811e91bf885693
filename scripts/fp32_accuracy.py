"""FP32 mode against FP64 mode on the same problem: relative error of the variational parameters after 1..N iterations
(where the FP32-mode error comes from and how it grows).  Usage: python scripts/fp32_accuracy.py [n p K iters]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import pcmf_b200
from pcmf_b200 import datagen

n, p, K, iters = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (3000, 1500, 20, 5)))
X, _, _ = datagen.simulate(n, p, 20, signal_U=20.0, signal_V=20.0, prob1=0.3, seed=5)
a1, a2, b1, b2 = datagen.random_init(X, K, seed=11)
for zi, sparse in ((True, False), (True, True)):
    kw = dict(a1=a1, a2=a2, b1=b1, b2=b2)
    if sparse:
        kw.update(prob_S=np.full((p, K), 0.5), prior_S=np.full(p, 0.5))
    s64 = pcmf_b200.Solver(X, K, zi, sparse, precision=0, **kw)
    s32 = pcmf_b200.Solver(X, K, zi, sparse, precision=1, **kw)
    print("model zi=%d sparse=%d" % (zi, sparse))
    for it in range(iters):
        s64.iterate(1)
        s32.iterate(1)
        o64, o32 = s64.get_output(want_prob_D=False), s32.get_output(want_prob_D=False)
        row = []
        for k in ("a1", "a2", "b1", "b2", "prior_prob_D"):
            d = np.abs(o32[k] - o64[k]) / np.maximum(np.abs(o64[k]), 1e-300)
            sgn = np.mean(np.sign(o32[k] - o64[k]))
            row.append("%s max %.2e mean %.2e sign %+.2f" % (k, d.max(), d.mean(), sgn))
        print("  iter %d: %s" % (it, " | ".join(row)))
    s64.close(); s32.close()
