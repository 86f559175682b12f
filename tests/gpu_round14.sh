#!/bin/bash
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_PDL=0 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_TC_SLEEP=0 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_ATTN_SKIP_TAIL=1 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1
