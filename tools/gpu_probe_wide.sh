#!/bin/bash
# Wide-path parity tests + the per-shape probe with the in-kernel counters (tools/wide_probe.py). Usage (under gpurun):
#   tools/gpu_probe_wide.sh [rows] [waves]     OWC_QKV_ATTN=0 / OWC_WIDE_EXP=2|4 select the A/B and ablation runs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_path.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/wide_probe.py ${1:-300000} ${2:-8} 2>&1 | tee gpurun_out/wide_probe.log
