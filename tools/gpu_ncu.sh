#!/bin/bash
# ncu full capture of selected kernels from a short bench run. Usage: tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-p}; KRE=${2:-k_gemm_tc}; SKIP=${3:-60}; CNT=${4:-4}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f \
    python bench.py --rows 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_$TAG.log
