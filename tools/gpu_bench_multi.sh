#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_$N.json 2> gpurun_out/r2_bench_$N.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
b=json.loads(open(f'gpurun_out/r2_bench_{n}.json').read().strip().splitlines()[-1])
def brief(r):
    out={k:r.get(k) for k in ('value','ms_per_step','gpu_launches')}
    out['e2e']=r['e2e']['value']; out['detail']=r.get('detail')
    if 'roofline' in r: out['roof']={k:r['roofline'][k] for k in ('achieved','frac','share_of_step','per_kernel_ms','class_names')}
    return out
print(json.dumps(brief(b)))
for k,v in b.get('sub',{}).items():
    print(k, json.dumps(brief(v) if 'error' not in v else v))
PY
