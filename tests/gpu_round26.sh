#!/bin/bash
# evidence: ncu --set full of the layer's second GEMM (dense|fc2, 4128 x 2048 x 10240) with the final tile order
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:gemm_tcgen05 --launch-skip 12 --launch-count 1 -o gpurun_out/r2_gemm2_order1 -f python tests/gemm_order_probe.py > gpurun_out/ncu_gemm2.log 2>&1; echo "== ncu gemm2 rc=$?"
python profiles/summarize.py full gpurun_out/r2_gemm2_order1.ncu-rep > gpurun_out/r2_gemm_dense_fc2_full.txt 2>&1; head -40 gpurun_out/r2_gemm_dense_fc2_full.txt
