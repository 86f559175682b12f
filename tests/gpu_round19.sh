#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_train.log 2>&1; echo "== train rc=$?"; tail -4 gpurun_out/test_gpu_train.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mm_projector or embeddings" -p no:cacheprovider > gpurun_out/mmp_tests.log 2>&1; echo "== mmp rc=$?"; tail -6 gpurun_out/mmp_tests.log
