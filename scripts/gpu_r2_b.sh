#!/bin/bash
# Round 2, GPU session B: all GPU tests without -x (every failure visible), then the C4 bench in both precisions.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 -x --deselect tests/test_gpu_fp32.py ) > gpurun_out/r2b_pytest.log 2>&1
tail -22 gpurun_out/r2b_pytest.log
( time timeout 600 python -m pytest tests/test_gpu_fp32.py -m gpu -q -x ) > gpurun_out/r2b_pytest_fp32.log 2>&1
tail -30 gpurun_out/r2b_pytest_fp32.log
timeout 600 python bench.py --no-cpu --no-e2e --precision fp32 > gpurun_out/r2b_bench_c4_fp32.json 2> gpurun_out/r2b_bench_c4_fp32.err
tail -c 600 gpurun_out/r2b_bench_c4_fp32.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_c4_fp32.json"))
    print("fp32 value", d["value"], "ms", d["ms_per_step"])
    for k,v in d["phases"].items(): print(" ", k, round(v["ms_per_step"],3))
except Exception as e: print("no fp32 bench", e)
PY
