#!/bin/bash
# Round-2 ncu evidence for the wide (mpnet-base, c3) path: launch list of one short step + full capture of the GEMMs / attention.
TAG=${1:-r02}
mkdir -p gpurun_out
ARGS="--rows 60000 --steps 1 --warmup 1 --sub  --no-cpu-baseline --no-sim-leg"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --rows 60000 --steps 1 --warmup 3 --sub '' --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt; cat gpurun_out/launch_summary_$TAG.txt
# class-name encode = 62 launches first; skip into the prediction chunks, second layer: 4 GEMM shapes x 2 layers
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_wide -s 56 -c 8 -o gpurun_out/prof_wide_$TAG -f \
    python bench.py --rows 60000 --steps 1 --warmup 3 --sub '' --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_wide_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_attn_grp|k_embed_ln|k_pool_ln_l2" -s 16 -c 4 -o gpurun_out/prof_attn_$TAG -f \
    python bench.py --rows 60000 --steps 1 --warmup 3 --sub '' --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_attn_$TAG.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
