"""One-off measurement of BASELINE.json configs[3]'s per-GPU share (t2i 512x512: N = 1024 image tokens, L = 1155, 18 steps,
CFG 5, 8 images per GPU, full-size random-init model) through the same public API as bench.py.  Prints one JSON line."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import showo_b200
from showo_b200 import _lib

bench.N_TOK, bench.L_SEQ = 1024, 129 + 1 + 1024 + 1
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.require_gpu()
model = showo_b200.Showo(False, bench.V, 50295, materialize=False)
model._make_engine(dev)
for name, t in bench.gpu_random_weights(torch, dev, seed=0):
    _lib.check(lib.showo_load_weight(model._engine, name.encode(), _lib.ptr(t), t.numel(), 1), name)
_lib.check(lib.showo_weights_complete(model._engine), "complete")
model._streamed = True
vq = showo_b200.MAGVITv2(materialize=False)
vq.load_weights(bench.gpu_random_magvit_weights(torch, dev, seed=1), device=dev)
cfg = bench.t2i_config()
cond_h, unc_h, descs = bench.synth_prompts(torch, 8, seed=1234)
cond_d, unc_d = cond_h.to(dev), unc_h.to(dev)
ids = torch.empty_like(cond_d)


def step():
    ids.copy_(cond_d)
    codes = model.t2i_generate(ids, unc_d, descs, guidance_scale=bench.CFG_W, timesteps=bench.T_STEPS, config=cfg)
    return vq.decode_code_uint8(torch.clamp(codes, 0, bench.CODEBOOK - 1))


imgs = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    imgs = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
N, L, P, T = 1024, 1155, 129, bench.T_STEPS
f_img = T * 2 * ((N + 2) * (bench.G_TOK + bench.A_PAIR * L) + N * 2 * bench.D * bench.CODEBOOK) + 2 * P * (bench.G_TOK + bench.A_PAIR * P / 2)
print(json.dumps({"workload": "t2i 512x512, 18 steps, CFG 5, batch 8, 1 GPU", "images_per_s": round(8 / (ms / 1e3), 3),
                  "ms_per_batch": round(ms, 2), "image_shape": list(imgs.shape),
                  "algorithmic_tflops": round((f_img + 1205e9) * 8 / (ms / 1e3) / 1e12, 1), "launches": model.kernel_launches()}))
