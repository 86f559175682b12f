#!/bin/bash
# evidence session: launch list of one bench step, ncu --set full of the attention and GEMM kernels, training-step kernel shares
mkdir -p gpurun_out
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15000 --launch-count 2700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "== ncu launches rc=$?"
python profiles/summarize.py launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches_by_kernel.txt 2>&1; head -16 gpurun_out/r2_launches_by_kernel.txt
cd tests
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:omni_attention_tc_kernel --launch-skip 3 --launch-count 1 -o ../gpurun_out/r2_attn_tc -f python attn_probe.py > ../gpurun_out/ncu_attn.log 2>&1; echo "== ncu attn rc=$?"
cd ..
python profiles/summarize.py full gpurun_out/r2_attn_tc.ncu-rep > gpurun_out/r2_attention_tc_full.txt 2>&1; head -30 gpurun_out/r2_attention_tc_full.txt
timeout 600 python tests/train_trace.py 2>&1 | grep -v Warn | head -24
