#!/bin/bash
# ncu capture of k_qkv_attn (per-line stall reasons): second layer of a prediction chunk
mkdir -p gpurun_out
export OWC_QKV_ATTN=1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_qkv_attn -s 14 -c 1 -o gpurun_out/prof_qkvattn_r02 -f \
    python bench.py --rows 60000 --steps 1 --warmup 3 --sub '' --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_qkvattn_r02.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_qkvattn_r02.log
ls -la gpurun_out/prof_qkvattn_r02.ncu-rep
