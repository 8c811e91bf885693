#!/bin/bash
# Round 2, GPU session C: FP32 mode tests, FP32-mode C4 bench, ncu of the two tcgen05 passes.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fp32.py -m gpu -q ) > gpurun_out/r2c_pytest_fp32.log 2>&1
tail -30 gpurun_out/r2c_pytest_fp32.log | cut -c1-300
CMD="python bench.py --steps 3 --no-cpu --no-e2e --precision fp32"
timeout 600 $CMD > gpurun_out/r2c_bench_c4_fp32.json 2> gpurun_out/r2c_bench_c4_fp32.err || tail -c 600 gpurun_out/r2c_bench_c4_fp32.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c_bench_c4_fp32.json"))
    print("fp32 value", d["value"], "ms", d["ms_per_step"])
    for k,v in d["phases"].items(): print(" ", k, round(v["ms_per_step"],3))
except Exception as e: print("no fp32 bench", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zi_tc -s 6 -c 2 -o gpurun_out/r2c_prof_zi_tc $CMD > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
