#!/bin/bash
mkdir -p gpurun_out
export OWC_QKV_ATTN=1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_path.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/wide_probe.py 300000 8 2>&1 | tee gpurun_out/r2n_probe.log
