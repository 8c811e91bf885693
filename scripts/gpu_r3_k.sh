#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_monitors.py tests/test_gpu_golden.py -m gpu -q ) > gpurun_out/r3k_pytest.log 2>&1
grep -E "^E  |passed|failed|Error" gpurun_out/r3k_pytest.log | cut -c1-260 | head -30
timeout 600 python scripts/full_call_c4.py 2>&1 | grep -v "^  " | cut -c1-300
