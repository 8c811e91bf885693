#!/bin/bash
# Round 2, session 3, GPU call G: whole GPU suite, default C4 bench (with CPU arm and e2e), FP32 line, launch list, ncu capture of the two integer-tcgen05 ZI passes
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r3g_pytest.log 2>&1
tail -16 gpurun_out/r3g_pytest.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r3g_bench_c4.json 2> gpurun_out/r3g_bench_c4.err || tail -c 800 gpurun_out/r3g_bench_c4.err
python - <<'PY'
import json
for f in ("c4",):
    try:
        d=json.load(open("gpurun_out/r3g_bench_%s.json"%f))
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", d["roofline"]["frac"])
        for k,v in (d.get("phases") or {}).items(): print("    ", k, round(v["ms_per_step"],3))
    except Exception as e: print(f, "failed", e)
PY
