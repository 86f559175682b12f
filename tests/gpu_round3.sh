#!/bin/bash
# Full status session: every -m gpu test file in its own process, then the probes and the bench.
mkdir -p gpurun_out
for f in tests/test_gpu_train.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_variants.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s --maxfail=30 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "== $n rc=$?"; tail -4 gpurun_out/$n.log
done
rm -f gpurun_out/attn_probe.jsonl
SHOWO_ATTN_TC=0 timeout 300 python tests/attn_probe.py 2>&1 | tail -5
SHOWO_ATTN_TC=1 timeout 300 python tests/attn_probe.py 2>&1 | tail -5
timeout 600 python tests/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "== e2e probe rc=$?"; tail -12 gpurun_out/e2e_probe.log
timeout 600 python tests/train_probe.py 2 > gpurun_out/train_probe.log 2>&1; echo "== train_probe rc=$?"; tail -2 gpurun_out/train_probe.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-3000
