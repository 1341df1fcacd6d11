#!/bin/bash
mkdir -p gpurun_out
export OWC_QKV_ATTN=1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_path.py -m gpu -x -q 2>&1 | tail -12
for qa in 1 0 1 0; do OWC_QKV_ATTN=$qa timeout 600 python tools/wide_probe.py 300000 8 2>&1 | head -1 | sed "s/^/qkv_attn=$qa /"; done | tee gpurun_out/r2m_probe.log
