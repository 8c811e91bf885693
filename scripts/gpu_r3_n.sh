#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( PCMF_I8_TW=1 timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r3n_pytest.log 2>&1
grep -E "^E  |passed|failed|Error" gpurun_out/r3n_pytest.log | cut -c1-260 | head -10
bash scripts/gpu_r3_c.sh tw0:PCMF_I8_TW=0 tw1:PCMF_I8_TW=1
