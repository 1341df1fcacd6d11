#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_path.py -m gpu -x -q 2>&1 | tail -4
for fs in 1 0 1 0; do OWC_WIDE_FFN2_SPLIT=$fs timeout 600 python tools/wide_probe.py 300000 8 2>&1 | head -1 | sed "s/^/ffn2split=$fs /"; done | tee gpurun_out/r2k_probe.log
for w in 2 4 16; do OWC_WIDE_WAVES=$w timeout 600 python tools/wide_probe.py 300000 $w 2>&1 | head -1 | sed "s/^/waves=$w /"; done | tee -a gpurun_out/r2k_probe.log
