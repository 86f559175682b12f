"""Kernel timeline of the MMU decode path on the live (concurrent, PDL) stream, from CUPTI through torch.profiler (not a test):
per kernel type the mean duration and the mean gap to the previous kernel's end.   python tests/decode_trace.py [n_new]"""
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
    n_new = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.require_gpu()
    model = showo_b200.Showo(False, bench.V, 50295, materialize=False)
    model._make_engine(dev)
    for name, t in bench.gpu_random_weights(torch, dev, seed=0):
        _lib.check(lib.showo_load_weight(model._engine, name.encode(), _lib.ptr(t), t.numel(), 1), f"load {name}")
    _lib.check(lib.showo_weights_complete(model._engine), "weights_complete")
    model._streamed = True
    B, L0 = 16, 276
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 50257, (B, L0), generator=g).to(dev)
    descs = [(0, 0, 0, 0, 259)] * B
    for _ in range(2):
        model.mmu_generate_batched(ids, attention_mask=descs, max_new_tokens=n_new, top_k=1)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.mmu_generate_batched(ids, attention_mask=descs, max_new_tokens=n_new, top_k=1)
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    recs = [(e.name.split("(")[0].replace("void ", "").replace("showo::", "")[:48], e.time_range.start, e.time_range.end) for e in ev]
    # keep the decode part: everything after the first mmu_finish_token kernel
    first = next(i for i, r in enumerate(recs) if "mmu_finish_token" in r[0])
    recs = recs[first:]
    dur, gap, cnt = collections.Counter(), collections.Counter(), collections.Counter()
    for i, (n, s, e) in enumerate(recs):
        dur[n] += e - s
        cnt[n] += 1
        if i:
            gap[n] += s - recs[i - 1][2]
    total = recs[-1][2] - recs[0][1]
    steps = sum(1 for r in recs if "mmu_finish_token" in r[0]) - 1
    print(f"decode part: {total:.1f} us over {steps} steps = {total / max(steps, 1):.1f} us per step")
    out = {}
    for n in sorted(dur, key=lambda k: -dur[k]):
        out[n] = {"n": cnt[n], "mean_dur_us": dur[n] / cnt[n], "mean_gap_before_us": gap[n] / cnt[n]}
        print(f"{n:50s} n={cnt[n]:4d} dur {dur[n] / cnt[n]:7.2f} us   gap before {gap[n] / cnt[n]:7.2f} us   share {100 * dur[n] / total:5.1f}%")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"total_us": total, "steps": steps, "kernels": out, "first_200": recs[:200]}, open(os.path.join(ROOT, "gpurun_out", "decode_trace.json"), "w"))


if __name__ == "__main__":
    main()
