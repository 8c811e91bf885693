#!/bin/bash
# Round 2, session 3, GPU call C: A/B of kernel variants through environment switches (bench only, phases printed).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/r3c_$tag.json 2> gpurun_out/r3c_$tag.err || tail -c 1500 gpurun_out/r3c_$tag.err
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/r3c_%s.json"%sys.argv[1]))
    print(sys.argv[1], "value %.3f ms %.2f"%(d["value"], d["ms_per_step"]), " ".join("%s=%.2f"%(k,v["ms_per_step"]) for k,v in d["phases"].items() if v["ms_per_step"]>1))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  run $tag $(echo $envs | tr ',' ' ')
done
