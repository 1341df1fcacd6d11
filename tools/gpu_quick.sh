#!/bin/bash
# tests + bench only (no ncu). Usage: tools/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/tests_$TAG.log; tail -8 gpurun_out/tests_$TAG.log
for MODE in ${MODES:-2}; do
OWC_GEMM_MODE=$MODE timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_m$MODE.json 2> gpurun_out/bench_${TAG}_m$MODE.err; echo "bench mode $MODE rc=$?"
tail -c 400 gpurun_out/bench_${TAG}_m$MODE.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_m$MODE.json"))
print("mode $MODE: value %.0f pred/s  ms/step %.2f  e2e %.0f  launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]))
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["per_kernel_ms"])
print("sim", d.get("sim_gemm_topk"))
print("clocks", d["clocks"])
PY
done
