#!/bin/bash
# round-2 session 3: mm_projector training path + mixed input, editing flows, bench with the config-5 training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -x -p no:cacheprovider -k "mm_projector or mixed or clip_vit" > gpurun_out/train_new.log 2>&1; echo "== train new rc=$?"; grep -v Warn gpurun_out/train_new.log | tail -12
timeout 900 python -m pytest tests/test_gpu_editing.py -m gpu -q -s -p no:cacheprovider > gpurun_out/editing.log 2>&1; echo "== editing rc=$?"; grep -v Warn gpurun_out/editing.log | tail -8
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "not (mm_projector or mixed or clip_vit)" > gpurun_out/train_old.log 2>&1; echo "== train old rc=$?"; tail -3 gpurun_out/train_old.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "mm_projector or embeddings" > gpurun_out/mmp.log 2>&1; echo "== mmp fwd rc=$?"; tail -2 gpurun_out/mmp.log
SHOWO_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log > gpurun_out/bench_line.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_line.json'))
    print('t2i', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
    for k in ('secondary','secondary_t2i512','secondary_train'):
        v=d.get(k) or {}
        print(k, v.get('value'), v.get('unit'), v.get('ms_per_step', v.get('ms_per_decode_step')), (v.get('roofline') or {}).get('frac'), v.get('losses'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.log').read()[-3000:])
PY
