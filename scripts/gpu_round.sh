#!/bin/bash
# One GPU session: parity tests, the default bench (with cpu_baseline + e2e), the reference arm, the ncu launch list of a
# short bench run and `ncu --set full` captures of the hot kernels of one timed C4 step.  Results land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 300 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
CMD2="python bench.py --steps 2 --no-cpu --no-e2e"
timeout 600 $CMD2 > gpurun_out/plain_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $CMD2 > gpurun_out/ncu_launches.log 2>&1
CMD="python bench.py --steps 1 --no-cpu --no-e2e"
timeout 600 $CMD > gpurun_out/plain_short1.log 2>&1 || exit 1
# launches matching each filter before the timed step: 3 warm-up iterations (+ init for the gene side, + the warm-up's final data-term pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zi_cells_mma -s 3 -c 1 -o gpurun_out/prof_c4_zi_cells $CMD > gpurun_out/ncu_full_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zi_genes_ -s 3 -c 1 -o gpurun_out/prof_c4_zi_genes $CMD > gpurun_out/ncu_full_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gene_pass -s 7 -c 2 -o gpurun_out/prof_c4_gene_pass $CMD > gpurun_out/ncu_full_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_estep_rows -s 3 -c 1 -o gpurun_out/prof_c4_rows $CMD > gpurun_out/ncu_full_d.log 2>&1
tail -1 gpurun_out/ncu_full_d.log
