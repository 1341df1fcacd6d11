#!/bin/bash
mkdir -p gpurun_out
export OWC_QKV_ATTN=1
OWC_WIDE_EXP=4 timeout 600 python tools/wide_probe.py 300000 8 2>&1 | grep -v "^{" | head -3 | tee gpurun_out/r2n_probe.log
