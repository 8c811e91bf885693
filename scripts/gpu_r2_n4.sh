#!/bin/bash
# Round 2: C4 bench at N = $1 GPUs (default 4) with the set-up stage timers of rank 0
N=${1:-4}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PCMF_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err
grep "pcmf timing" gpurun_out/r2_bench_c4_n$N.err | tail -60 | cut -c1-110
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_c4_n$N.json"))
print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("seconds_total"), "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"))
for k,v in d["phases"].items(): print("   ", k, round(v["ms_per_step"],3))
PY
