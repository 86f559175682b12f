"""In-kernel timeline of the tcgen05 attention kernel (CTA 0): SHOWO_TC_PROF=1 python tests/tc_prof_probe.py  (not a test)"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from showo_b200 import _lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.require_gpu()
    H, D = 32, 2048
    n_seq, rows, pos0, n_keys = 16, 258, 129, 387
    Lmax = 448
    q = torch.randn(n_seq * rows, D, device=dev).bfloat16()
    o = torch.empty_like(q)
    kc = torch.randn(n_seq, H, Lmax, 64, device=dev).bfloat16()
    vt = torch.randn(n_seq, H, 64, Lmax, device=dev).bfloat16()
    md = torch.tensor([(60 + 4 * i if i < 8 else 126, 129, 387, 0, 0) for i in range(n_seq)], dtype=torch.int32, device=dev)
    for _ in range(3):
        _lib.check(lib.showo_attention_run(_lib.ptr(q), D, n_seq, rows, pos0, H, _lib.ptr(kc), _lib.ptr(vt), Lmax, n_keys, _lib.ptr(md),
                                           _lib.ptr(o), D, _lib.current_stream_ptr()), "attention_run")
    torch.cuda.synchronize()
    n = 3 * 64 * 6
    buf = (C.c_ulonglong * n)()
    lib.showo_debug_tc_prof.restype = C.c_int
    lib.showo_debug_tc_prof.argtypes = [C.c_void_p, C.c_int]
    assert lib.showo_debug_tc_prof(buf, n) == 0
    v = list(buf)
    t0 = min(x for x in v if x > 0)
    names = {0: "softmax(w2)  [wait_s, got_s, loaded+masked, exp_done, p_free, published]",
             1: "mma          [qk_enter, k_full, qk_issued, pv_enter, v+p_ready, pv_issued]",
             2: "tma          [enter, k_empty, issued]"}
    out = {}
    for role in range(3):
        print(names[role])
        rows_ = []
        for g in range(26):
            rec = v[(role * 64 + g) * 6:(role * 64 + g) * 6 + 6]
            if not any(rec):
                continue
            rel = [int(x - t0) if x else -1 for x in rec]
            rows_.append(rel)
            print(f"  block {g:2d}: " + " ".join(f"{x:7d}" for x in rel))
        out[role] = rows_
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tc_prof.json"), "w"))


if __name__ == "__main__":
    main()
