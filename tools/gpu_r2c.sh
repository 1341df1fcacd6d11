#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_tests.log; echo "tests rc=$?"; tail -15 gpurun_out/r2c_tests.log
OWC_WIDE_EXP=1 timeout 600 python tools/wide_probe.py 100000 8 > gpurun_out/r2c_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2c_probe.log
