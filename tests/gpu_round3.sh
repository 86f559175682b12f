#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -p no:cacheprovider -k "ln_fused or ln_fold" > gpurun_out/variants_ln.log 2>&1; echo "== variants rc=$?"; tail -3 gpurun_out/variants_ln.log
timeout 600 python tests/train_trace.py 2>&1 | grep -v Warn | head -30
timeout 600 python tests/decode_trace.py 12 2>&1 | grep -v Warn | head -14
