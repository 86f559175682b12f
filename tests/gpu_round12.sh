#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/decode_probe.jsonl
timeout 300 python tests/decode_probe.py 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_ATTN_TC=0 timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_ATTN_SKIP_TAIL=1 timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1
timeout 600 python tests/e2e_probe.py 2>&1 | tail -9
timeout 300 python tests/decode_probe.py 2>&1 | tail -1
