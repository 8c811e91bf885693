#!/bin/bash
# `ncu --set full` of the hot kernels of the timed C4 step, one kernel family per capture.
set -x
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 1 --no-cpu --no-e2e"
timeout 600 $CMD > gpurun_out/plain_short1.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zi_cells_mma -s 4 -c 1 -o gpurun_out/prof_c4_zi_cells $CMD > gpurun_out/ncu_full_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zi_genes_mma -s 3 -c 1 -o gpurun_out/prof_c4_zi_genes $CMD > gpurun_out/ncu_full_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gene_pass -s 7 -c 2 -o gpurun_out/prof_c4_gene_pass $CMD > gpurun_out/ncu_full_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_estep_rows -s 3 -c 1 -o gpurun_out/prof_c4_rows $CMD > gpurun_out/ncu_full_d.log 2>&1
tail -1 gpurun_out/ncu_full_d.log
