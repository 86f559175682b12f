#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "== bench2 rc=$?"; tail -3 gpurun_out/bench_2gpu.log | cut -c1-1500
tail -1 gpurun_out/bench_2gpu.log > gpurun_out/bench_2gpu_line.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_2gpu_line.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'])
    for k in ('secondary','secondary_t2i512','secondary_train'):
        v=d.get(k) or {}
        print(k, v.get('value'), v.get('unit'), v.get('ms_per_step', v.get('ms_per_decode_step')), (v.get('roofline') or {}).get('frac'))
except Exception as e:
    print('parse failed', e)
PY
