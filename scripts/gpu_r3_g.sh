#!/bin/bash
# Round 2, session 3, GPU call G: whole GPU suite, default C4 bench (with CPU arm and e2e legs), FP32 line, reference arm
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r3g_pytest.log 2>&1
tail -14 gpurun_out/r3g_pytest.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r3g_bench_c4.json 2> gpurun_out/r3g_bench_c4.err || tail -c 800 gpurun_out/r3g_bench_c4.err
timeout 600 python bench.py --precision fp32 --no-cpu --no-e2e > gpurun_out/r3g_bench_c4_fp32.json 2> gpurun_out/r3g_bench_c4_fp32.err || tail -c 800 gpurun_out/r3g_bench_c4_fp32.err
timeout 600 python bench.py --impl reference > gpurun_out/r3g_bench_reference.json 2> gpurun_out/r3g_bench_reference.err || tail -c 800 gpurun_out/r3g_bench_reference.err
python - <<'PY'
import json
for f in ("c4","c4_fp32","reference"):
    try:
        d=json.load(open("gpurun_out/r3g_bench_%s.json"%f))
        print(f, "value", d["value"], "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "clocks", d.get("clocks"))
        for k,v in (d.get("phases") or {}).items(): print("    ", k, round(v["ms_per_step"],3))
    except Exception as e: print(f, "failed", e)
PY
