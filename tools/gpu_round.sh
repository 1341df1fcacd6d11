#!/bin/bash
# Round evidence for the fused-encoder path: bench, reference arm, ncu launch list, ncu full capture of k_encoder_fused.
# Usage: tools/gpu_round.sh <tag>
TAG=${1:-r01f}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
tail -c 700 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --rows 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt; cat gpurun_out/launch_summary_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encoder_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f \
    python bench.py --rows 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_fused_$TAG.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_fused_$TAG.log
