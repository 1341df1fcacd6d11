#!/bin/bash
# ncu evidence for the wide (mpnet-base, c3) path: launch list of one short step + full captures of k_gemm_wide, k_qkv_attn
# and the HBM-side kernels. Usage: tools/gpu_ncu_wide.sh [tag]   (under gpurun, one GPU)
TAG=${1:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --rows 60000 --steps 1 --warmup 3 --sub '' --no-cpu-baseline --no-sim-leg"
[ -n "$SKIP_LIST" ] || timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv \
    $BENCH > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt; cat gpurun_out/launch_summary_$TAG.txt
# class-name encode first; skip into the prediction chunks: the 3 k_gemm_wide shapes of one layer, one k_qkv_attn (gpurun brings back at most 64 MiB)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_wide -s 42 -c 3 -o gpurun_out/prof_wide_$TAG -f \
    $BENCH > gpurun_out/ncu_wide_$TAG.log 2>&1; echo "ncu k_gemm_wide rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_qkv_attn -s 14 -c 1 -o gpurun_out/prof_qkvattn_$TAG -f \
    $BENCH > gpurun_out/ncu_qkvattn_$TAG.log 2>&1; echo "ncu k_qkv_attn rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_embed_ln|k_pool_ln_l2" -s 4 -c 2 -o gpurun_out/prof_hbm_$TAG -f \
    $BENCH > gpurun_out/ncu_hbm_$TAG.log 2>&1; echo "ncu hbm kernels rc=$?"
ls -la gpurun_out/*.ncu-rep
