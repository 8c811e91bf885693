#!/bin/bash
# One GPU session: parity tests, the default bench (with cpu_baseline + e2e), the ncu launch list of a short bench run and
# one `ncu --set full` capture of the five hot kernels of one C4 step.  Results land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 400 gpurun_out/bench_default.err
timeout 600 python bench.py --steps 2 --no-cpu --no-e2e > gpurun_out/plain_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --steps 2 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py --steps 1 --no-cpu --no-e2e > gpurun_out/plain_short1.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'k_estep_rows|k_gene_pass|k_zi_cells_mma|k_zi_genes_mma' -s 16 -c 5 -o gpurun_out/prof_c4_step \
    python bench.py --steps 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
