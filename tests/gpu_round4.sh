#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_inputs.py -m gpu -q -s -p no:cacheprovider > gpurun_out/test_gpu_train_inputs.log 2>&1; echo "== train_inputs rc=$?"; tail -5 gpurun_out/test_gpu_train_inputs.log
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:omni_attention_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/attn_tc_v2 -f python tests/attn_probe.py > gpurun_out/ncu_attn.log 2>&1; echo "== ncu attn rc=$?"; tail -3 gpurun_out/ncu_attn.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:omni_attention --csv --log-file gpurun_out/attn_launches.csv python tests/attn_probe.py > /dev/null 2>&1; echo "== ncu list rc=$?"
for s in 4 5 6 8; do SHOWO_SKINNY_STAGES=$s timeout 300 python tests/decode_probe.py 2>&1 | tail -1; done
timeout 300 python tests/decode_probe.py 2>&1 | tail -1
