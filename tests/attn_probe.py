"""Timing probe of the prefill / denoise-step attention kernels alone (not a test): CUDA events over 20 launches per geometry.
    SHOWO_ATTN_TC=0|1 python tests/attn_probe.py     -> JSON lines, appended to gpurun_out/attn_probe.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from showo_b200 import _lib  # noqa: E402

GEOMS = [  # name, n_seq, rows, pos0, n_keys, descriptor
    ("t2i256_step", 16, 258, 129, 387, (60, 129, 387, 0, 0)),
    ("t2i512_step", 16, 1026, 129, 1155, (60, 129, 1155, 0, 0)),
    ("train_1155", 8, 1155, 0, 1155, (60, 129, 1155, 0, 0)),
    ("mmu_prefill_276", 16, 276, 0, 276, (0, 0, 0, 0, 259)),
]


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.require_gpu()
    H, D = 32, 2048
    out = []
    for name, n_seq, rows, pos0, n_keys, desc in GEOMS:
        Lmax = (n_keys + 63) // 64 * 64
        q = torch.randn(n_seq * rows, D, device=dev).bfloat16()
        o = torch.empty_like(q)
        kc = torch.randn(n_seq, H, Lmax, 64, device=dev).bfloat16()
        vt = torch.randn(n_seq, H, 64, Lmax, device=dev).bfloat16()
        md = torch.tensor([desc] * n_seq, dtype=torch.int32, device=dev)

        def run():
            _lib.check(lib.showo_attention_run(_lib.ptr(q), D, n_seq, rows, pos0, H, _lib.ptr(kc), _lib.ptr(vt), Lmax, n_keys, _lib.ptr(md),
                                               _lib.ptr(o), D, _lib.current_stream_ptr()), "attention_run")
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        flops = 4.0 * n_seq * H * rows * n_keys * 64
        rec = {"geom": name, "attn_tc": os.environ.get("SHOWO_ATTN_TC", "default"), "us": round(us, 1), "dense_tflops": round(flops / us / 1e6, 1),
               "finite": bool(torch.isfinite(o.float()).all())}
        print(json.dumps(rec))
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "attn_probe.jsonl"), "a") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
