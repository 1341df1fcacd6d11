#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2i_tests.log; echo "tests rc=$?"; tail -6 gpurun_out/r2i_tests.log
for g in 1 0 1 0; do
OWC_ATTN_GRP=$g timeout 600 python bench.py --workload c2 --sub '' --no-cpu-baseline --no-sim-leg --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('grp=$g', round(b['value']), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']), 'strings', round(b['e2e_strings']['value']), b['roofline']['achieved'], b['roofline']['per_kernel_ms'], b['clocks']['sm_mhz'])
"
done | tee gpurun_out/r2i_c2.log
