#!/bin/bash
# Run on the GPU box via gpurun: smoke, bench, ncu launch list, ncu full capture of the top kernels.
# Usage: tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_ref_$TAG.json
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --rows 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
# full capture of the encoder GEMM and the similarity kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 60 -c 4 -o gpurun_out/prof_gemm_$TAG -f \
    python bench.py --rows 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-sim-leg > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
