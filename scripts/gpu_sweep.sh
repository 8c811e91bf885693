#!/bin/bash
# The smaller BASELINE configurations on one GPU: C5 (200k x 10k, ZI) at K = 2, 10, 30, 64, and C3 / C2.
cd "$(dirname "$0")/.."
for c in c5k2 c5k10 c5k30 c5k64 c3 c2; do
    timeout 600 python bench.py --config $c --no-cpu > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err || tail -3 gpurun_out/bench_$c.err
done
python - <<'PY'
import json
out = []
for k in (2, 10, 30, 64):
    d = json.load(open("gpurun_out/bench_c5k%d.json" % k))
    out.append({"K": k, "iter_per_s": d["value"], "ms_per_step": d["ms_per_step"], "e2e_iter_per_s": d["e2e"]["value"],
                "phases_ms": {p: v["ms_per_step"] for p, v in d["phases"].items() if v["ms_per_step"] > 0.05},
                "kernels": [{x: kk.get(x) for x in ("kernel", "ms", "bound", "achieved", "peak", "unit", "frac", "gather_frac", "stored_prob_D_bytes")} for kk in d["kernels"]]})
    print(k, d["value"], d["ms_per_step"])
json.dump(out, open("gpurun_out/bench_c5_ksweep.json", "w"), indent=1)
for c in ("c3", "c2"):
    d = json.load(open("gpurun_out/bench_%s.json" % c)); print(c, d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
