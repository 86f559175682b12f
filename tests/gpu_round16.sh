#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "skinny or mmu or decode or megakernel or step_and" --maxfail=30 -p no:cacheprovider > gpurun_out/decode_tests.log 2>&1; echo "== decode tests rc=$?"; tail -4 gpurun_out/decode_tests.log
rm -f gpurun_out/decode_probe.jsonl
timeout 300 python tests/decode_probe.py 2>&1 | tail -1
SHOWO_DECODE_LN_FUSED=0 timeout 300 python tests/decode_probe.py 2>&1 | tail -1
timeout 600 python tests/decode_trace.py 12 > gpurun_out/decode_trace.txt 2>&1; echo "== decode trace rc=$?"; cat gpurun_out/decode_trace.txt | tail -9
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q -k config2 -p no:cacheprovider 2>&1 | tail -2
