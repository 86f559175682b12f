#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --maxfail=30 -p no:cacheprovider > gpurun_out/parity_tests.log 2>&1; echo "== parity tests rc=$?"; tail -5 gpurun_out/parity_tests.log
cd tests; timeout 300 python attn_trace.py 2>&1 | grep -v Warn | tail -4; cd ..
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_full_size.py tests/test_gpu_train_inputs.py -m gpu -q -p no:cacheprovider > gpurun_out/train_tests.log 2>&1; echo "== train+full rc=$?"; tail -3 gpurun_out/train_tests.log
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_ATTN_TC=0 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1
