#!/bin/bash
mkdir -p gpurun_out
for g in 1 0; do echo "=== OWC_ATTN_GRP=$g"; OWC_ATTN_GRP=$g timeout 600 python tools/fused_check.py --rows 100000 --no-oracle 2>&1 | grep -v "^    [0-9 ]* *[0-9]* *S \|^    [0-9 ]* *[0-9]* *  M " ; done > gpurun_out/r2j_fused.log 2>&1
head -120 gpurun_out/r2j_fused.log
