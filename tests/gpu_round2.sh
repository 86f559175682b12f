#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "attention or forward or t2i or mask" --maxfail=30 -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "== attn tests rc=$?"; tail -15 gpurun_out/attn_tests.log
rm -f gpurun_out/attn_probe.jsonl
SHOWO_ATTN_TC=0 timeout 300 python tests/attn_probe.py 2>&1 | tail -5
SHOWO_ATTN_TC=1 timeout 300 python tests/attn_probe.py 2>&1 | tail -5
timeout 600 python tests/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "== e2e probe rc=$?"; tail -12 gpurun_out/e2e_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches.csv python tests/train_probe.py 0 > gpurun_out/train_ncu.log 2>&1; echo "== ncu train rc=$?"; tail -2 gpurun_out/train_ncu.log
python profiles/summarize.py launches gpurun_out/train_launches.csv > gpurun_out/train_launches_by_kernel.txt 2>&1; head -30 gpurun_out/train_launches_by_kernel.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
