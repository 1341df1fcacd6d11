#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 600 gpurun_out/r2g_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
def brief(r):
    out={k:r.get(k) for k in ('value','ms_per_step','gpu_launches')}
    out['e2e']=r['e2e']['value']; 
    if 'e2e_strings' in r: out['e2e_strings']=r['e2e_strings']['value']
    if 'roofline' in r: out['roof']={k:r['roofline'][k] for k in ('achieved','frac','launches_per_step','share_of_step','whole_step_tflops','per_kernel_ms')}
    if 'sim_gemm_topk_in_step' in r: out['sim_in_step']=r['sim_gemm_topk_in_step']
    return out
print(json.dumps(brief(b),indent=1))
for k,v in b.get('sub',{}).items():
    print(k, json.dumps(brief(v) if 'error' not in v else v,indent=1))
print({k:b.get(k) for k in ('sim_gemm_topk','sim_gemm_topk_wide','cpu_baseline','clocks')})
PY
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 --ref-budget-s 120 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err ) 2>&1 | grep real; tail -c 900 gpurun_out/r2g_ref.json
