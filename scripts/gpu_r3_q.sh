#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 2048; do
  timeout 600 python bench.py --config c3 --variant $v --no-cpu --no-e2e --steps 10 > gpurun_out/r3q_c3_v$v.json 2> gpurun_out/r3q_c3_v$v.err || tail -c 600 gpurun_out/r3q_c3_v$v.err
  python - $v <<'PY'
import json,sys
d=json.load(open("gpurun_out/r3q_c3_v%s.json"%sys.argv[1]))
print("c3 variant", sys.argv[1], "value %.2f ms %.3f"%(d["value"], d["ms_per_step"]), {k:round(v["ms_per_step"],2) for k,v in d["phases"].items() if v["ms_per_step"]>0.05})
PY
done
