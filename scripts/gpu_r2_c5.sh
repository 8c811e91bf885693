#!/bin/bash
# Round 2: C5 latent-dimension sweep (200k x 10k ZI, K = 2 / 10 / 30 / 64)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 2 10 30 64; do
  timeout 600 python bench.py --config c5k$k --steps 10 --no-cpu --no-e2e $C5_ARGS > gpurun_out/r2_bench_c5k$k.json 2> gpurun_out/r2_bench_c5k$k.err || tail -c 600 gpurun_out/r2_bench_c5k$k.err
done
python - <<'PY'
import json
for k in (2,10,30,64):
    try:
        d=json.load(open("gpurun_out/r2_bench_c5k%d.json"%k))
        print("K=%d value %.2f it/s  %.3f ms"%(k,d["value"],d["ms_per_step"]), {n:round(v["ms_per_step"],2) for n,v in d["phases"].items() if v["ms_per_step"]>0.05})
    except Exception as e: print(k,"failed",e)
PY
