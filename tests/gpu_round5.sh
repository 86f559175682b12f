#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "attention or forward or t2i or mask or prefix" --maxfail=30 -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "== attn tests rc=$?"; tail -8 gpurun_out/attn_tests.log
rm -f gpurun_out/attn_probe.jsonl
timeout 300 python tests/attn_probe.py 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:omni_attention --csv --log-file gpurun_out/attn_launches.csv python tests/attn_probe.py > /dev/null 2>&1; echo "== ncu list rc=$?"
python profiles/summarize.py launches gpurun_out/attn_launches.csv | head -8
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > gpurun_out/full_train.log 2>&1; echo "== full+train rc=$?"; tail -5 gpurun_out/full_train.log
SHOWO_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-700
