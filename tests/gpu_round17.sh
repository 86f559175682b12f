#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "gemm or forward or prefix or teacher or step_and" --maxfail=30 -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1; echo "== gemm tests rc=$?"; tail -5 gpurun_out/gemm_tests.log
SHOWO_BENCH_HEADLINE_ONLY=1 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-400
SHOWO_BENCH_HEADLINE_ONLY=1 SHOWO_GEMM_STREAMK=0 timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-400
python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from showo_b200 import _lib
lib = _lib.require_gpu(); dev = torch.device('cuda', 0)
m_, n_, k_ = 16*258, 2048, 10240
A = (torch.randn(m_, k_, device=dev)*0.5).bfloat16(); B = (torch.randn(n_, k_, device=dev)*0.02).bfloat16()
o = torch.zeros(m_, n_, device=dev)
S = _lib.current_stream_ptr
for _ in range(3): lib.showo_gemm_bf16(_lib.ptr(A), k_, _lib.ptr(B), k_, m_, n_, k_, _lib.ptr(o), n_, None, _lib.ptr(o), n_, n_, 1, 0, S())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): lib.showo_gemm_bf16(_lib.ptr(A), k_, _lib.ptr(B), k_, m_, n_, k_, _lib.ptr(o), n_, None, _lib.ptr(o), n_, n_, 1, 0, S())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/20
print('dense_fc2 %s: %.4f ms  %.1f TFLOP/s' % (os.environ.get('SHOWO_GEMM_STREAMK','on'), ms, 2.0*m_*n_*k_/ms/1e9))
PY
