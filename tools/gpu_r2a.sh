#!/bin/bash
# round 2, first GPU pass: parity tests, then c3 with the wide path on/off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log; echo "tests rc=$?"; tail -25 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --rows 200000 --sub '' --no-cpu-baseline --no-sim-leg > gpurun_out/r2a_c3_wide.json 2> gpurun_out/r2a_c3_wide.err; echo "bench wide rc=$?"; tail -c 2500 gpurun_out/r2a_c3_wide.json; tail -5 gpurun_out/r2a_c3_wide.err
OWC_WIDE=0 timeout 600 python bench.py --steps 3 --warmup 3 --rows 200000 --sub '' --no-cpu-baseline --no-sim-leg > gpurun_out/r2a_c3_old.json 2> gpurun_out/r2a_c3_old.err; echo "bench old rc=$?"; tail -c 1500 gpurun_out/r2a_c3_old.json; tail -5 gpurun_out/r2a_c3_old.err
