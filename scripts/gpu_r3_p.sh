#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gene_order -s 3 -c 1 -o gpurun_out/r3p_prof_order python scripts/full_call_c4.py > gpurun_out/r3p_ncu.log 2>&1
tail -3 gpurun_out/r3p_ncu.log | cut -c1-200
