"""Small driver for profiling the MMU decode loop (full-size random weights, B=16, L0=276): used under ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import showo_b200
from showo_b200 import _lib
import bench

dev = torch.device("cuda", 0)
lib = _lib.require_gpu()
model = showo_b200.Showo(False, bench.V, 50295, materialize=False)
model._make_engine(dev)
for name, t in bench.gpu_random_weights(torch, dev, seed=0):
    _lib.check(lib.showo_load_weight(model._engine, name.encode(), _lib.ptr(t), t.numel(), 1), name)
_lib.check(lib.showo_weights_complete(model._engine), "complete")
model._streamed = True
n_new = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.manual_seed(0)
ids = torch.randint(0, 50000, (16, 276), device=dev)
descs = [(0, 0, 0, 0, 259)] * 16
for _ in range(2):
    toks, _ = model.mmu_generate_batched(ids, attention_mask=descs, max_new_tokens=n_new, top_k=1)
torch.cuda.synchronize()
if os.environ.get("MMU_TIMING"):
    # ms per decode step from the difference of two generation lengths (prefill cancels)
    def run(n):
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.mmu_generate_batched(ids, attention_mask=descs, max_new_tokens=n, top_k=1)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    a, b = run(8), run(72)
    print("decode ms/step %.4f  (prefill+8 steps %.3f ms)  env %s" % ((b - a) / 64, a, {k: v for k, v in os.environ.items() if k.startswith("SHOWO_")}))
if os.environ.get("MMU_SAVE"):
    t24, _ = model.mmu_generate_batched(ids, attention_mask=descs, max_new_tokens=24, top_k=1)
    torch.save(t24.cpu(), os.environ["MMU_SAVE"])
    if os.environ.get("MMU_COMPARE"):
        other = torch.load(os.environ["MMU_COMPARE"])
        first_diff = [(int((a != b).nonzero()[0]) if (a != b).any() else 24) for a, b in zip(t24.cpu(), other)]
        print("tokens equal to %s up to step (per row, 24 = all): %s" % (os.environ["MMU_COMPARE"], first_diff))
print("done", toks.shape)
