#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_formats.py -m gpu -q ) > gpurun_out/r3h_pytest.log 2>&1
grep -E "^E  |passed|failed|Error" gpurun_out/r3h_pytest.log | cut -c1-260 | head -40
bash scripts/gpu_r3_c.sh nw4:PCMF_I8_NW=4
