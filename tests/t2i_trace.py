"""Kernel timeline of one t2i denoise-step forward (16 rows x 387, full-size model) on the live PDL stream, from CUPTI through
torch.profiler (not a test): per kernel type the count, mean duration and share.   [SHOWO_LN_FOLD=1] python tests/t2i_trace.py"""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import showo_b200  # noqa: E402
from showo_b200 import _lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.require_gpu()
    model = showo_b200.Showo(False, bench.V, 50295, materialize=False)
    model._make_engine(dev)
    for name, t in bench.gpu_random_weights(torch, dev, seed=0):
        _lib.check(lib.showo_load_weight(model._engine, name.encode(), _lib.ptr(t), t.numel(), 1), f"load {name}")
    _lib.check(lib.showo_weights_complete(model._engine), "weights_complete")
    B, L = 16, 387
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 50257, (B, L), generator=g).to(dev)
    masks = _lib.masks_array([(100, 129, 387, 0, 0)] * B)
    logits = torch.empty(B * L * bench.V, device=dev, dtype=torch.float32)

    def run():
        _lib.check(lib.showo_forward(model._engine, _lib.ptr(ids), None, B, L, masks, _lib.ptr(logits), _lib.current_stream_ptr()), "forward")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    dur, cnt = collections.Counter(), collections.Counter()
    for e in ev:
        n = e.name.replace("void ", "").replace("showo::", "")[:70]
        dur[n] += e.time_range.end - e.time_range.start
        cnt[n] += 1
    total = (ev[-1].time_range.end - ev[0].time_range.start) / 3
    print(f"forward: {total:.1f} us")
    out = {"forward_us": total, "kernels": {}}
    for n in sorted(dur, key=lambda k: -dur[k]):
        out["kernels"][n] = {"n": cnt[n] // 3, "mean_dur_us": dur[n] / cnt[n]}
        print(f"{n:72s} n={cnt[n] // 3:4d} dur {dur[n] / cnt[n]:8.2f} us  share {100 * dur[n] / 3 / total:5.1f}%")
    tag = os.environ.get("SHOWO_LN_FOLD", "0")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"t2i_trace_fold{tag}.json"), "w"))


if __name__ == "__main__":
    main()
