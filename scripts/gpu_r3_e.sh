#!/bin/bash
# parity suites + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_monitors.py tests/test_gpu_scale_parity.py -m gpu -q -x ) > gpurun_out/r3e_pytest.log 2>&1
tail -15 gpurun_out/r3e_pytest.log | cut -c1-300
bash scripts/gpu_r3_c.sh nw8:PCMF_I8_NW=8
