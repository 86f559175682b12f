#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "token_major" > gpurun_out/tn.log 2>&1; echo "== tn gemm rc=$?"; tail -12 gpurun_out/tn.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > gpurun_out/train_mn.log 2>&1; echo "== train (mn wgrad) rc=$?"; tail -5 gpurun_out/train_mn.log
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -p no:cacheprovider -k "wgrad" > gpurun_out/variants_wgrad.log 2>&1; echo "== variants rc=$?"; tail -3 gpurun_out/variants_wgrad.log
timeout 600 python tests/train_trace.py 2>&1 | grep -v Warn | head -24
