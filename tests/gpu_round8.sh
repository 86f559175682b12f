#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --maxfail=30 -p no:cacheprovider > gpurun_out/parity_tests.log 2>&1; echo "== parity tests rc=$?"; tail -6 gpurun_out/parity_tests.log
rm -f gpurun_out/attn_probe.jsonl gpurun_out/decode_probe.jsonl
timeout 300 python tests/attn_probe.py 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:omni_attention --csv --log-file gpurun_out/attn_launches.csv python tests/attn_probe.py > /dev/null 2>&1; echo "== ncu list rc=$?"
SHOWO_L2_PREFETCH=0 timeout 300 python tests/decode_probe.py 2>&1 | tail -1
timeout 300 python tests/decode_probe.py 2>&1 | tail -1
timeout 600 python tests/decode_trace.py 12 > gpurun_out/decode_trace.txt 2>&1; echo "== decode trace rc=$?"; cat gpurun_out/decode_trace.txt | tail -9
