#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "attention or forward or t2i or mask or prefix" --maxfail=30 -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "== attn tests rc=$?"; tail -4 gpurun_out/attn_tests.log
SHOWO_TC_PROF=1 timeout 300 python tests/tc_prof_probe.py > gpurun_out/tc_prof.txt 2>&1; echo "== prof rc=$?"; cat gpurun_out/tc_prof.txt | head -90
rm -f gpurun_out/attn_probe.jsonl
timeout 300 python tests/attn_probe.py 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:omni_attention --csv --log-file gpurun_out/attn_launches.csv python tests/attn_probe.py > /dev/null 2>&1; echo "== ncu list rc=$?"
