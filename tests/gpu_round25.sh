#!/bin/bash
# A/B of the L2 eviction hints on the GEMM operand loads (SHOWO_GEMM_HINT=1 default / 0), parity first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "gemm or forward or prefix or conv" > gpurun_out/hint_parity.log 2>&1; echo "== parity rc=$?"; tail -2 gpurun_out/hint_parity.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider -k "golden or adamw" > gpurun_out/hint_train.log 2>&1; echo "== train rc=$?"; tail -2 gpurun_out/hint_train.log
for v in 1 0; do
  SHOWO_GEMM_HINT=$v timeout 300 python tests/gemm_order_probe.py 2>&1 | grep -v Warn | sed "s/^/hint=$v /"
  SHOWO_GEMM_HINT=$v timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_hint_ncu_$v.csv python tests/gemm_order_probe.py > gpurun_out/gemm_hint_ncu_$v.log 2>&1; echo "== ncu hint=$v rc=$?"
done
python - <<'PY'
import csv, collections
for v in (1, 0):
    rows = [r for r in csv.reader(open(f'gpurun_out/gemm_hint_ncu_{v}.csv')) if len(r) > 10]
    hdr = rows[0]; iN = hdr.index('Metric Name'); iV = hdr.index('Metric Value'); iI = hdr.index('ID')
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[iI], {})[r[iN]] = float(r[iV].replace(',', ''))
    ids = list(per)
    for s in range(len(ids) // 13):
        d = per[ids[s * 13 + 12]]
        print(f"hint={v} shape {s}: dram {(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6:.1f} MB  time {d.get('gpu__time_duration.sum', 0) / 1e3:.1f} us  L2 hit {d.get('lts__t_sector_hit_rate.pct', 0):.1f} %")
PY
for v in 1 0; do
  SHOWO_GEMM_HINT=$v SHOWO_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/hint_bench_$v.log 2>&1
  echo "== bench hint=$v rc=$?"
  tail -1 gpurun_out/hint_bench_$v.log > gpurun_out/hint_bench_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/hint_bench_$v.json'))
print('t2i', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'gemm', d['roofline']['per_shape'])
for k in ('secondary','secondary_t2i512','secondary_train'):
    v=d.get(k) or {}
    print(k, v.get('value'), v.get('ms_per_step', v.get('ms_per_decode_step')), (v.get('roofline') or {}).get('frac'))
PY
done
