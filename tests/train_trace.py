"""Per-kernel shares of one training step (forward + backward, configs[4] geometry) on the live stream through CUPTI (not a test):
    python tests/train_trace.py  -> gpurun_out/train_trace.txt"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import showo_b200  # noqa: E402
from showo_b200 import _lib  # noqa: E402
from train_probe import make_batch  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.require_gpu()
    model = showo_b200.Showo(False, bench.V, 50295, materialize=False)
    model._make_engine(dev)
    for name, t in bench.gpu_random_weights(torch, dev, seed=0):
        _lib.check(lib.showo_load_weight(model._engine, name.encode(), _lib.ptr(t), t.numel(), 1), f"load {name}")
    _lib.check(lib.showo_weights_complete(model._engine), "weights_complete")
    model._streamed = True
    ids, labels, descs, sizes = make_batch(dev)
    B, L = ids.shape
    terms = model._loss_terms(B, L, *sizes, 128)
    for _ in range(2):
        model.train_forward(ids, None, descs, labels, terms, want_logits=False)
        model.backward((1.0, 0.1, 1.0))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.train_forward(ids, None, descs, labels, terms, want_logits=False)
        model.backward((1.0, 0.1, 1.0))
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    dur, cnt = collections.Counter(), collections.Counter()
    for e in ev:
        n = e.name.split("(")[0].replace("void ", "").replace("showo::", "")[:60]
        dur[n] += e.time_range.end - e.time_range.start
        cnt[n] += 1
    total = ev[-1].time_range.end - ev[0].time_range.start
    lines = [f"# one training step (8 x 1155 rows, forward + backward), CUPTI kernel durations; wall {total / 1e3:.2f} ms, sum of kernels {sum(dur.values()) / 1e3:.2f} ms"]
    for n in sorted(dur, key=lambda k: -dur[k]):
        lines.append(f"{n:62s} n={cnt[n]:4d} total {dur[n] / 1e3:8.3f} ms  mean {dur[n] / cnt[n]:8.1f} us  share {100 * dur[n] / sum(dur.values()):5.1f}%")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "train_trace.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:22]))


if __name__ == "__main__":
    main()
