#!/bin/bash
# fp32 verification paths (MAGVIT-v2 and the backbone)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "fp32" > gpurun_out/fp32.log 2>&1; echo "== fp32 rc=$?"; grep -v Warn gpurun_out/fp32.log | tail -25
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "magvit or adamw or forward" > gpurun_out/fp32_side.log 2>&1; echo "== side rc=$?"; tail -2 gpurun_out/fp32_side.log
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "adamw or projector" > gpurun_out/fp32_train.log 2>&1; echo "== train rc=$?"; tail -2 gpurun_out/fp32_train.log
