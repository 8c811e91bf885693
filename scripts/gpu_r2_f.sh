#!/bin/bash
# Round 2, GPU session F: tiled non-zero passes -- tests, C4 bench, ncu full capture of the three tile-kernel launches of one step.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_tiles.py -m gpu -q -x ) > gpurun_out/r2f_pytest.log 2>&1
tail -5 gpurun_out/r2f_pytest.log
CMD="python bench.py --steps 5 --no-cpu --no-e2e"
timeout 600 $CMD > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err || tail -c 600 gpurun_out/r2f_bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench_c4.json"))
print("value", d["value"], "ms", d["ms_per_step"])
for k,v in d["phases"].items(): print(" ", k, round(v["ms_per_step"],3))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 9 -c 3 -o gpurun_out/r2f_prof_tile $CMD > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
