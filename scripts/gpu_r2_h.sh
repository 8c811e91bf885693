#!/bin/bash
# Round 2, GPU session H: whole GPU suite, default bench (C4, both precisions), C2 / C3 with the CPU arm at full size, launch list.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2h_pytest.log 2>&1
tail -16 gpurun_out/r2h_pytest.log
timeout 900 python bench.py > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err || tail -c 800 gpurun_out/r2h_bench_c4.err
timeout 600 python bench.py --precision fp32 --no-cpu --no-e2e > gpurun_out/r2h_bench_c4_fp32.json 2> gpurun_out/r2h_bench_c4_fp32.err || tail -c 800 gpurun_out/r2h_bench_c4_fp32.err
timeout 900 python bench.py --config c2 --cpu-rows 0 --steps 20 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err || tail -c 800 gpurun_out/r2h_bench_c2.err
timeout 1500 python bench.py --config c3 --cpu-rows 0 --steps 10 > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err || tail -c 800 gpurun_out/r2h_bench_c3.err
timeout 600 python bench.py --impl reference > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err || tail -c 800 gpurun_out/r2h_bench_reference.err
python - <<'PY'
import json
for f in ("c4","c4_fp32","c2","c3","reference"):
    try:
        d=json.load(open("gpurun_out/r2h_bench_%s.json"%f))
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "e2e_defaults", (d.get("e2e_defaults") or {}).get("value"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("extrapolated"), "refbuild", (d.get("reference_build") or {}).get("value"))
        for k,v in (d.get("phases") or {}).items(): print("    ", k, round(v["ms_per_step"],3))
    except Exception as e: print(f, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2h_ncu_launches.log 2>&1
tail -2 gpurun_out/r2h_ncu_launches.log | cut -c1-200
