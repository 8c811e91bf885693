#!/bin/bash
cd "$(dirname "$0")/.."
cp pcmf_b200/libpcmf_b200.so /tmp/lib_u1.so
bash scripts/gpu_r3_c.sh u1:X=1
for u in 2 4; do cp gpurun_lib_u$u.so pcmf_b200/libpcmf_b200.so; bash scripts/gpu_r3_c.sh u$u:X=1; done
cp /tmp/lib_u1.so pcmf_b200/libpcmf_b200.so
