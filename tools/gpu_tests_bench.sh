#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( time timeout 1800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err ) 2>&1 | grep real; tail -c 400 gpurun_out/r2_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
def brief(r):
    out={k:r.get(k) for k in ('value','ms_per_step','gpu_launches')}
    out['e2e']=r['e2e']['value']
    if 'e2e_strings' in r: out['e2e_strings']=r['e2e_strings']['value']
    if 'roofline' in r: out['roof']={k:r['roofline'][k] for k in ('achieved','frac','launches_per_step','share_of_step','whole_step_tflops','traffic','per_kernel_ms')}
    return out
print(json.dumps(brief(b)))
for k,v in b.get('sub',{}).items():
    print(k, json.dumps(brief(v) if 'error' not in v else v))
print({k:b.get(k) for k in ('sim_gemm_topk','cpu_baseline','clocks')})
PY
( time timeout 1800 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_ref.json 2> gpurun_out/r2_ref.err ) 2>&1 | grep real; tail -c 700 gpurun_out/r2_ref.json
