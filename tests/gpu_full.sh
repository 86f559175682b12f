#!/bin/bash
# Full status session: every -m gpu test file in its own process, then the bench.
mkdir -p gpurun_out
for f in tests/test_gpu_train.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_train_inputs.py tests/test_gpu_editing.py tests/test_gpu_clip.py tests/test_gpu_variants.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s --maxfail=30 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "== $n rc=$?"; tail -3 gpurun_out/$n.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log > gpurun_out/bench_line.json; cut -c1-2500 gpurun_out/bench_line.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_line.json'))
for k in ('secondary','secondary_t2i512','secondary_train'):
    v=d.get(k) or {}
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_step', v.get('ms_per_decode_step')), (v.get('roofline') or {}).get('frac'))
PY
