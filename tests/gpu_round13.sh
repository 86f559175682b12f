#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "attention or forward or t2i or mask or prefix" --maxfail=30 -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "== attn tests rc=$?"; tail -4 gpurun_out/attn_tests.log
cd tests; timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4; SHOWO_ATTN_SKIP_TAIL=1 timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4; cd ..
SHOWO_TC_PROF=1 timeout 300 python tests/tc_prof_probe.py > gpurun_out/tc_prof.txt 2>&1; echo "== prof rc=$?"; cat gpurun_out/tc_prof.txt | head -24
rm -f gpurun_out/decode_probe.jsonl
timeout 300 python tests/decode_probe.py 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_ATTN_TC=0 timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1
