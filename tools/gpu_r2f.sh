#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_path.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2f_tests.log; echo "tests rc=$?"; tail -5 gpurun_out/r2f_tests.log
for rot in 1 0 1 0; do OWC_WIDE_ROT=$rot timeout 600 python tools/wide_probe.py 100000 8 2>&1 | head -1 | sed "s/^/rot=$rot /"; done | tee gpurun_out/r2f_probe.log
