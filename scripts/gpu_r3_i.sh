#!/bin/bash
# Round 2, session 3, GPU call I: C5 latent-dimension sweep and C2 / C3 lines with the integer ZI passes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 2 10 30 64; do
  timeout 600 python bench.py --config c5k$k --steps 10 --no-cpu --no-e2e > gpurun_out/r3_bench_c5k$k.json 2> gpurun_out/r3_bench_c5k$k.err || tail -c 600 gpurun_out/r3_bench_c5k$k.err
done
timeout 900 python bench.py --config c3 --no-cpu --steps 10 > gpurun_out/r3_bench_c3.json 2> gpurun_out/r3_bench_c3.err || tail -c 600 gpurun_out/r3_bench_c3.err
python - <<'PY'
import json
for k in ("c5k2","c5k10","c5k30","c5k64","c3"):
    try:
        d=json.load(open("gpurun_out/r3_bench_%s.json"%k))
        print("%s value %.2f it/s  %.3f ms e2e %s"%(k,d["value"],d["ms_per_step"],(d.get("e2e") or {}).get("value")), {n:round(v["ms_per_step"],2) for n,v in d["phases"].items() if v["ms_per_step"]>0.05})
    except Exception as e: print(k,"failed",e)
PY
