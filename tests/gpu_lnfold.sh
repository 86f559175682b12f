#!/bin/bash
# A/B of the LayerNorm-folding switch: parity files under SHOWO_LN_FOLD=1, then the headline bench with and without it.
mkdir -p gpurun_out
for f in tests/test_gpu_parity.py tests/test_gpu_full_size.py; do
  n=$(basename $f .py)
  SHOWO_LN_FOLD=1 timeout 1200 python -m pytest $f -m gpu -q -s --maxfail=30 -p no:cacheprovider > gpurun_out/lnfold_$n.log 2>&1
  echo "== fold $n rc=$?"; tail -4 gpurun_out/lnfold_$n.log
done
for v in 1 0; do
  SHOWO_LN_FOLD=$v SHOWO_BENCH_SKIP_CPU=1 SHOWO_BENCH_SKIP_TRAIN=1 timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/lnfold_bench_$v.log 2>&1
  echo "== bench fold=$v rc=$?"
  tail -1 gpurun_out/lnfold_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], d['ms_per_step'], (d.get('secondary') or {}).get('value'), (d.get('secondary') or {}).get('ms_per_decode_step'), (d.get('secondary_t2i512') or {}).get('value'))"
done
