#!/bin/bash
# Round 2 session 3: C4 bench at N GPUs (N = first argument)
N=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/r3_bench_c4_n$N.json 2> gpurun_out/r3_bench_c4_n$N.err
tail -c 400 gpurun_out/r3_bench_c4_n$N.err
python - $N <<'PY'
import json,sys
d=json.load(open("gpurun_out/r3_bench_c4_n%s.json"%sys.argv[1]))
print("N=%s value"%sys.argv[1], d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"))
for k,v in d["phases"].items(): print("   ", k, round(v["ms_per_step"],3))
PY
