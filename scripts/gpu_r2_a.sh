#!/bin/bash
# Round 2, GPU session A: all GPU tests (incl. the new scale parity tests), the default C4 bench, the launch list and one
# `ncu --set full` capture of pass A with the FP64-pipe / DMMA / bank-conflict metrics added.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; free -g | head -2
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2a_pytest.log 2>&1
tail -25 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
tail -c 400 gpurun_out/r2a_bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2a_bench_c4.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k,v in d["phases"].items(): print(" ", k, round(v["ms_per_step"],3))
PY
CMD="python bench.py --steps 1 --no-cpu --no-e2e"
M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"
timeout 600 $CMD > gpurun_out/r2a_plain.log 2>&1 &&
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_zi_cells_mma -s 3 -c 1 -o gpurun_out/r2a_prof_zi_cells $CMD > gpurun_out/r2a_ncu_a.log 2>&1
tail -3 gpurun_out/r2a_ncu_a.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_c4.csv python bench.py --steps 2 --no-cpu --no-e2e > gpurun_out/r2a_ncu_launches.log 2>&1
tail -2 gpurun_out/r2a_ncu_launches.log
