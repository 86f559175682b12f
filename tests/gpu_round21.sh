#!/bin/bash
# round-2 session 3: native CLIP ViT tower, editing flows on arbitrary extrapolation grids
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clip.py -m gpu -q -s -p no:cacheprovider > gpurun_out/clip.log 2>&1; echo "== clip rc=$?"; grep -v Warn gpurun_out/clip.log | tail -14
timeout 900 python -m pytest tests/test_gpu_editing.py -m gpu -q -s -p no:cacheprovider > gpurun_out/editing.log 2>&1; echo "== editing rc=$?"; grep -v Warn gpurun_out/editing.log | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "magvit or conv" > gpurun_out/magvit.log 2>&1; echo "== magvit rc=$?"; tail -2 gpurun_out/magvit.log
timeout 300 python -m pytest tests/test_host_logic.py -q -p no:cacheprovider -k "clip" > gpurun_out/clip_host.log 2>&1; echo "== clip host rc=$?"; tail -2 gpurun_out/clip_host.log
