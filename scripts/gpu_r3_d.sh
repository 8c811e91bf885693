#!/bin/bash
# (needs the library built with PCMF_EXTRA_NVCC=-DPCMF_I8_ABLATE: PCMF_I8_ABL bits switch parts of k_zi_cells_i8 off; timing only, results are wrong)
cd "$(dirname "$0")/.."
bash scripts/gpu_r3_c.sh base:PCMF_I8_ABL=0 nosttm:PCMF_I8_ABL=65 nostg:PCMF_I8_ABL=2 nocs:PCMF_I8_ABL=4 noent:PCMF_I8_ABL=8 nodmma:PCMF_I8_ABL=16 noexpit:PCMF_I8_ABL=32 nommaissue:PCMF_I8_ABL=64 nothing:PCMF_I8_ABL=127
