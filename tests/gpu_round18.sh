#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -p no:cacheprovider > gpurun_out/test_gpu_train.log 2>&1; echo "== train rc=$?"; tail -8 gpurun_out/test_gpu_train.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mmu or embeddings" -p no:cacheprovider > gpurun_out/mmu_tests.log 2>&1; echo "== mmu rc=$?"; tail -4 gpurun_out/mmu_tests.log
timeout 600 python tests/train_probe.py 2 2>&1 | tail -1
SHOWO_BENCH_SKIP_CPU=1 timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log > gpurun_out/bench_line.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_line.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
for k in ('secondary','secondary_t2i512','secondary_train'):
    v=d.get(k) or {}
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_step', v.get('ms_per_decode_step')), (v.get('roofline') or {}).get('frac'), v.get('losses'))
PY
