#!/bin/bash
# Round 2, session 3, GPU call A: parity of the integer-tcgen05 cell-side ZI pass (smoke, parity suites), then a short C4 bench.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/r3a_smoke.log 2>&1; tail -5 gpurun_out/r3a_smoke.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_monitors.py -m gpu -q -x --durations=5 ) > gpurun_out/r3a_pytest.log 2>&1
tail -25 gpurun_out/r3a_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/r3a_bench_c4.json 2> gpurun_out/r3a_bench_c4.err || tail -c 1500 gpurun_out/r3a_bench_c4.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r3a_bench_c4.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
    for k,v in (d.get("phases") or {}).items(): print("    ", k, round(v["ms_per_step"],3))
except Exception as e: print("failed", e)
PY
