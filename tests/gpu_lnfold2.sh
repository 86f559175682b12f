#!/bin/bash
mkdir -p gpurun_out
SHOWO_LN_FOLD=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "forward or t2i or mmu or prefix" > gpurun_out/lnfold_parity.log 2>&1; echo "== fold parity rc=$?"; tail -2 gpurun_out/lnfold_parity.log
for v in 0 1; do
  SHOWO_LN_FOLD=$v timeout 600 python tests/t2i_trace.py > gpurun_out/t2i_trace_$v.log 2>&1; echo "== trace fold=$v rc=$?"; head -8 gpurun_out/t2i_trace_$v.log | cut -c1-150
done
for v in 1 0; do
  SHOWO_LN_FOLD=$v SHOWO_BENCH_SKIP_CPU=1 SHOWO_BENCH_SKIP_TRAIN=1 timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/lnfold_bench_$v.log 2>&1
  echo "== bench fold=$v rc=$?"
  tail -1 gpurun_out/lnfold_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], d['ms_per_step'], (d.get('secondary') or {}).get('value'), (d.get('secondary') or {}).get('ms_per_decode_step'), (d.get('secondary_t2i512') or {}).get('value'))"
done
