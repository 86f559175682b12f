#!/bin/bash
# evidence: DRAM traffic of the GEMM per tile order (ncu), then the launch list of one bench step
mkdir -p gpurun_out
for v in 1 0; do
  SHOWO_GEMM_ORDER=$v timeout 300 python tests/gemm_order_probe.py 2>&1 | grep -v Warn > gpurun_out/gemm_order_time_$v.txt; cat gpurun_out/gemm_order_time_$v.txt
  SHOWO_GEMM_ORDER=$v timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_order_ncu_$v.csv python tests/gemm_order_probe.py > gpurun_out/gemm_order_ncu_$v.log 2>&1; echo "== ncu order=$v rc=$?"
done
python - <<'PY'
import csv, collections
for v in (1, 0):
    rows = [r for r in csv.reader(open(f'gpurun_out/gemm_order_ncu_{v}.csv')) if len(r) > 10]
    hdr = rows[0]; iN = hdr.index('Metric Name'); iV = hdr.index('Metric Value'); iI = hdr.index('ID')
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[iI], {})[r[iN]] = float(r[iV].replace(',', ''))
    ids = list(per)
    # 13 launches per shape (3 warm-up + 10 timed): report the last one of each shape
    for s in range(len(ids) // 13):
        d = per[ids[s * 13 + 12]]
        print(f"order={v} shape {s}: dram {(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6:.1f} MB  time {d.get('gpu__time_duration.sum', 0) / 1e3:.1f} us  L2 hit {d.get('lts__t_sector_hit_rate.pct', 0):.1f} %")
PY
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15000 --launch-count 2700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "== ncu launches rc=$?"
python profiles/summarize.py launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches_by_kernel.txt 2>&1; head -12 gpurun_out/r2_launches_by_kernel.txt
gzip -f gpurun_out/r2_launches.csv
