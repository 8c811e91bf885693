#!/bin/bash
# stored-path parity tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -m gpu -q -x -k "stored or c3 or c4 or scale or blocked" ) > gpurun_out/r3f_pytest.log 2>&1
tail -4 gpurun_out/r3f_pytest.log | cut -c1-300
bash scripts/gpu_r3_c.sh nw8:PCMF_I8_NW=8 nw4:PCMF_I8_NW=4
