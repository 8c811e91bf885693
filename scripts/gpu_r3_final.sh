#!/bin/bash
# Round 2, session 3: final check of the tree -- smoke, whole GPU suite, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r3final_pytest.log 2>&1
tail -4 gpurun_out/r3final_pytest.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r3final_bench_c4.json 2> gpurun_out/r3final_bench_c4.err || tail -c 800 gpurun_out/r3final_bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3final_bench_c4.json"))
print("c4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"), "frac", d["roofline"]["frac"], "launches", d.get("gpu_launches"))
PY
