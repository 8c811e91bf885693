#!/bin/bash
# Round 2 session 3, two GPUs: the multi-device tests (in-library ngpus driver, sharded emulation) and the C4 bench at N = 2.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cabi_client.py tests/test_gpu_sharded.py -m gpu -q ) > gpurun_out/r3n2_pytest.log 2>&1
tail -6 gpurun_out/r3n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r3_bench_c4_n2.json 2> gpurun_out/r3_bench_c4_n2.err
tail -c 600 gpurun_out/r3_bench_c4_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3_bench_c4_n2.json"))
print("N=2 value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"))
for k,v in d["phases"].items(): print("   ", k, round(v["ms_per_step"],3))
PY
