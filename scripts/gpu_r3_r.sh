#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_scale_parity.py -m gpu -q -x -k "dense_mid_size or c3_shape" ) > gpurun_out/r3r_pytest.log 2>&1
grep -E "^E  |passed|failed|Error" gpurun_out/r3r_pytest.log | cut -c1-260 | head
timeout 900 python bench.py --config c3 --cpu-rows 0 --steps 10 > gpurun_out/r3_bench_c3.json 2> gpurun_out/r3_bench_c3.err || tail -c 600 gpurun_out/r3_bench_c3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3_bench_c3.json"))
print("c3 value %.2f ms %.3f e2e %s cpu %s"%(d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value")), {k:round(v["ms_per_step"],2) for k,v in d["phases"].items() if v["ms_per_step"]>0.05})
PY
