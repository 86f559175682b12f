#!/bin/bash
# One GPU-box session: every -m gpu test file in its own process (a faulting kernel cannot poison the others), then the probes.
mkdir -p gpurun_out
for f in tests/test_gpu_train.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_variants.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s --maxfail=30 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "== $n rc=$?"; tail -4 gpurun_out/$n.log
done
timeout 600 python tests/train_probe.py 2 > gpurun_out/train_probe.log 2>&1; echo "== train_probe rc=$?"; tail -2 gpurun_out/train_probe.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-1500
