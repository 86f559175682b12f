"""DRAM traffic of the persistent GEMM per tile order (not a test): the shapes whose A operand is the larger one, launched through
showo_gemm_bf16 -- run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` with
SHOWO_GEMM_ORDER=1 (default: N first when M > N) and =0 (M first, the first version).   python tests/gemm_order_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from showo_b200 import _lib  # noqa: E402

SHAPES = [("dense_fc2_t2i256", 4128, 2048, 10240, 1), ("dense_fc2_train", 9240, 2048, 10240, 1), ("dgrad_w1_train", 9240, 2048, 14336, 2),
          ("wgrad_w1_train", 14336, 2048, 9240, 3)]


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.require_gpu()
    S = _lib.current_stream_ptr
    for name, m_, n_, k_, epi in SHAPES:
        if epi == 3:            # token-major operands: A = [K, M], B = [K, N]
            A = (torch.randn(k_, m_, device=dev) * 0.5).bfloat16()
            Bw = (torch.randn(k_, n_, device=dev) * 0.02).bfloat16()
            lda, ldb = m_, n_
        else:
            A = (torch.randn(m_, k_, device=dev) * 0.5).bfloat16()
            Bw = (torch.randn(n_, k_, device=dev) * 0.02).bfloat16()
            lda, ldb = k_, k_
        o = torch.zeros(m_, n_, device=dev)
        r = o if epi == 1 else None
        for _ in range(3):
            _lib.check(lib.showo_gemm_bf16(_lib.ptr(A), lda, _lib.ptr(Bw), ldb, m_, n_, k_, _lib.ptr(o), n_, None, _lib.ptr(r), n_, n_, epi, 0, S()), name)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.showo_gemm_bf16(_lib.ptr(A), lda, _lib.ptr(Bw), ldb, m_, n_, k_, _lib.ptr(o), n_, None, _lib.ptr(r), n_, n_, epi, 0, S())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        alg = (m_ * k_ + n_ * k_) * 2 + m_ * n_ * 4 * (2 if epi == 1 else 1)
        print(f"{name:20s} order={os.environ.get('SHOWO_GEMM_ORDER', '1')} M={m_} N={n_} K={k_}: {ms * 1e3:8.1f} us  {2.0 * m_ * n_ * k_ / ms / 1e9:7.1f} TFLOP/s  algorithmic {alg / 1e6:.1f} MB")
        del A, Bw, o


if __name__ == "__main__":
    main()
