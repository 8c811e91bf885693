#!/bin/bash
# A/B of prebuilt library variants gpurun_lib_<tag>.so against the in-tree build (bench only)
cd "$(dirname "$0")/.."
cp pcmf_b200/libpcmf_b200.so /tmp/lib_base.so
bash scripts/gpu_r3_c.sh base:X=1
for f in gpurun_lib_*.so; do t=${f#gpurun_lib_}; t=${t%.so}; cp $f pcmf_b200/libpcmf_b200.so; bash scripts/gpu_r3_c.sh $t:X=1; done
cp /tmp/lib_base.so pcmf_b200/libpcmf_b200.so
