#!/bin/bash
mkdir -p gpurun_out
cd tests
SHOWO_ATTN_TC=0 timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4
timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4
SHOWO_TC_SLEEP=0 timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4
SHOWO_TC_SLEEP=100 timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4
cd ..
SHOWO_TC_PROF=1 timeout 300 python tests/tc_prof_probe.py > gpurun_out/tc_prof.txt 2>&1; echo "== prof rc=$?"; cat gpurun_out/tc_prof.txt | head -52
