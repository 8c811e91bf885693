#!/bin/bash
# Round 2, session 3, final captures: ncu full of the two integer ZI passes (final code) + launch list of two bench steps
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --no-cpu --no-e2e"
M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_zi_cells_i8 -s 2 -c 1 -o gpurun_out/r3t_prof_zi_cells_i8 $CMD > gpurun_out/r3t_ncu_a.log 2>&1
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_zi_genes_i8 -s 1 -c 1 -o gpurun_out/r3t_prof_zi_genes_i8 $CMD > gpurun_out/r3t_ncu_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3t_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r3t_ncu_launches.log 2>&1
ls -la gpurun_out/r3t*
