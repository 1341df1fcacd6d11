#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_tests.log; echo "tests rc=$?"; tail -15 gpurun_out/r2b_tests.log
timeout 900 python tools/wide_probe.py 100000 1,2,4,8,16 > gpurun_out/r2b_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2b_probe.log
