#!/bin/bash
# ncu launch list (per-launch gpu time; cold-cache, serialised -> compare SHARES). Usage: tools/gpu_launchlist.sh <tag> [mode]
TAG=${1:-l}; MODE=${2:-2}
mkdir -p gpurun_out
OWC_GEMM_MODE=$MODE timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --rows 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches_$TAG.csv
