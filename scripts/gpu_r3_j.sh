#!/bin/bash
# Round 2, session 3, GPU call J: ncu full capture of the cell-side integer-tcgen05 ZI pass (final form) at C4
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --no-cpu --no-e2e"
M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active"
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_zi_cells_i8 -s 2 -c 1 -o gpurun_out/r3j_prof_zi_cells_i8 $CMD > gpurun_out/r3j_ncu.log 2>&1
tail -3 gpurun_out/r3j_ncu.log
