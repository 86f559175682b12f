"""Timing probe of the training step at BASELINE.json configs[4] geometry (not a test): full-size random-init model,
per-GPU micro-batch 8 = 3 t2i + 1 lm + 4 mmu-vit rows of L = 1155, forward and backward timed with CUDA
events.    python tests/train_probe.py [iters]    -> one JSON line (also written to gpurun_out/train_probe.json)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import showo_b200  # noqa: E402
from showo_b200 import _lib  # noqa: E402


def make_batch(dev, B_t2i=3, B_lm=1, B_mmu=4, L=1155, seed=0):
    g = torch.Generator().manual_seed(seed)
    V, Vt, P = bench.V, 50305, 129
    ids = torch.randint(0, 50257, (B_t2i + B_lm + B_mmu, L), generator=g)
    labels = torch.full_like(ids, -100)
    descs = []
    for b in range(B_t2i):
        n = int(torch.randint(8, 65, (1,), generator=g))
        ids[b, :P - n - 3] = 50295
        ids[b, P - n - 3] = 50300
        ids[b, P] = 50296
        codes = torch.randint(Vt, Vt + 8192, (L - P - 2,), generator=g)
        masked = torch.rand(L - P - 2, generator=g) < 0.6
        ids[b, P + 1:L - 1] = torch.where(masked, torch.full_like(codes, V - 1), codes)
        labels[b, P + 1:L - 1] = torch.where(masked, codes, torch.full_like(codes, -100))
        ids[b, L - 1] = 50297
        descs.append((P - n - 3, P, L, 0, 0))
    for b in range(B_t2i, B_t2i + B_lm):
        labels[b] = ids[b]
        descs.append((0, 0, 0, 0, 0))
    for b in range(B_t2i + B_lm, B_t2i + B_lm + B_mmu):
        labels[b, L - 548:] = ids[b, L - 548:]
        descs.append((0, 0, 0, 30, 606))
    return ids.to(dev), labels.to(dev), descs, (B_t2i, B_lm, B_mmu)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
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

    def ev():
        return torch.cuda.Event(enable_timing=True)
    fw, bw = [], []
    for it in range(iters + 1):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        _, losses = model.train_forward(ids, None, descs, labels, terms, want_logits=False)
        e1.record()
        model.backward((1.0, 0.1, 1.0))
        e2.record()
        torch.cuda.synchronize()
        if it > 0:
            fw.append(e0.elapsed_time(e1)); bw.append(e1.elapsed_time(e2))
    gnorm = float(model.read_grad("showo.model.layers.12.mlp.fc1.weight", shape=(8192, 2048)).norm())
    f_step = 3 * B * L * (bench.G_TOK + bench.A_PAIR * L + 2 * bench.D * bench.V)
    ms = min(fw) + min(bw)
    peaks = bench.measured_peaks()
    out = {"workload": "train step, 8 x 1155 rows (3 t2i + 1 lm + 4 mmu-vit), full-size model", "forward_ms": round(min(fw), 2),
           "backward_ms": round(min(bw), 2), "step_ms": round(ms, 2), "tokens_per_s": round(B * L / ms * 1e3, 1),
           "algorithmic_tflop_per_step": round(f_step / 1e12, 2), "achieved_tflops": round(f_step / ms / 1e9, 1),
           "frac_of_sustained_peak": round(f_step / ms / 1e9 / peaks["bf16_sustained"], 4), "losses": losses[:, 0].tolist(),
           "grad_norm_layer12_fc1": gnorm, "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_probe.json"), "w"))


if __name__ == "__main__":
    main()
