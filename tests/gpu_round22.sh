#!/bin/bash
# A/B of the persistent GEMM's tile order (SHOWO_GEMM_ORDER=0: sweep M first always; 1: sweep N first when A is the larger operand)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "gemm or forward or prefix" > gpurun_out/order_parity.log 2>&1; echo "== parity rc=$?"; tail -2 gpurun_out/order_parity.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > gpurun_out/order_train.log 2>&1; echo "== train rc=$?"; tail -2 gpurun_out/order_train.log
for v in 1 0; do
  SHOWO_GEMM_ORDER=$v timeout 600 python tests/train_trace.py 2>&1 | grep -v Warn | head -9 > gpurun_out/order_train_trace_$v.txt; echo "== train trace order=$v"; cat gpurun_out/order_train_trace_$v.txt | cut -c1-150
  SHOWO_GEMM_ORDER=$v timeout 600 python tests/t2i_trace.py > gpurun_out/order_t2i_trace_$v.txt 2>&1; echo "== t2i trace order=$v"; head -6 gpurun_out/order_t2i_trace_$v.txt | cut -c1-150
done
for v in 1 0; do
  SHOWO_GEMM_ORDER=$v SHOWO_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/order_bench_$v.log 2>&1
  echo "== bench order=$v rc=$?"
  tail -1 gpurun_out/order_bench_$v.log > gpurun_out/order_bench_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/order_bench_$v.json'))
print('t2i', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'gemm', d['roofline']['per_shape'])
for k in ('secondary','secondary_t2i512','secondary_train'):
    v=d.get(k) or {}
    print(k, v.get('value'), v.get('ms_per_step', v.get('ms_per_decode_step')), (v.get('roofline') or {}).get('frac'))
PY
done
