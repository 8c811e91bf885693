#!/bin/bash
# quick check: tile tests + C4 bench phases
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tiles.py -m gpu -q -x ) > gpurun_out/r2g_pytest.log 2>&1
tail -3 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 5 --no-cpu --no-e2e $BENCH_ARGS > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err || tail -c 600 gpurun_out/r2g_bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g_bench_c4.json"))
print("value", d["value"], "ms", d["ms_per_step"])
for k,v in d["phases"].items(): print(" ", k, round(v["ms_per_step"],3))
PY
