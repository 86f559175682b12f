"""Real (concurrent, PDL) per-kernel durations of the attention launches through CUPTI / torch.profiler (not a test):
    [SHOWO_ATTN_TC=0] python tests/attn_trace.py   -> per geometry: kernel name, mean duration, mean start-to-start period"""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from showo_b200 import _lib  # noqa: E402
from attn_probe import GEOMS  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.require_gpu()
    H, D = 32, 2048
    res = {}
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
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                run()
            torch.cuda.synchronize()
        ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
        dur, cnt = collections.Counter(), collections.Counter()
        for e in ev:
            n = e.name.split("(")[0].replace("void ", "").replace("showo::", "")[:40]
            dur[n] += e.time_range.end - e.time_range.start
            cnt[n] += 1
        total = (ev[-1].time_range.end - ev[0].time_range.start) / 10
        res[name] = {"per_call_us": round(total, 1), "kernels": {n: round(dur[n] / cnt[n], 1) for n in dur}}
        print(name, json.dumps(res[name]))
    tag = os.environ.get("SHOWO_ATTN_TC", "d") + "_" + os.environ.get("SHOWO_TC_SLEEP", "d")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"attn_trace_{tag}.json"), "w"))


if __name__ == "__main__":
    main()
