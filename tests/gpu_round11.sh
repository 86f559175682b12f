#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "attention or forward or t2i or mask or prefix" --maxfail=30 -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "== attn tests rc=$?"; tail -4 gpurun_out/attn_tests.log
cd tests; timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4; SHOWO_ATTN_SKIP_TAIL=1 timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4; cd ..
SHOWO_TC_PROF=1 timeout 300 python tests/tc_prof_probe.py > gpurun_out/tc_prof.txt 2>&1; echo "== prof rc=$?"; cat gpurun_out/tc_prof.txt | head -24
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/train_tests.log 2>&1; echo "== train+full rc=$?"; tail -3 gpurun_out/train_tests.log
SHOWO_BENCH_SKIP_CPU=1 SHOWO_BENCH_SKIP_TRAIN=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-900
