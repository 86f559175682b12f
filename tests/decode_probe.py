"""Timing probe of the MMU decode path alone (not a test): bench.mmu_decode_bench under the current SHOWO_* environment.
    SHOWO_SKINNY_STAGES=5 python tests/decode_probe.py   -> one JSON line, appended to gpurun_out/decode_probe.jsonl"""
import json
import os
import sys

import torch

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
    model._streamed = True
    vq = showo_b200.MAGVITv2(materialize=False)
    vq.load_weights(bench.gpu_random_magvit_weights(torch, dev, seed=1), device=dev)
    r = bench.mmu_decode_bench(torch, model, vq, dev, bench.measured_peaks())
    rec = {"env": {k: v for k, v in os.environ.items() if k.startswith("SHOWO_")}, "tokens_per_s": r["value"],
           "ms_per_decode_step": r["ms_per_decode_step"], "prefill_ms": r["prefill_ms"], "hbm_frac": r["roofline"]["frac"]}
    print(json.dumps(rec))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "decode_probe.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
