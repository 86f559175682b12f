"""Where does the end-to-end (reference call shape) step spend its time?  (not a test)  python tests/e2e_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import showo_b200  # noqa: E402
from showo_b200 import _lib, masks  # noqa: E402


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
    cfg = bench.t2i_config()
    cond, unc, descs = bench.synth_prompts(torch, 8, 1234)
    cond_d, unc_d = cond.to(dev), unc.to(dev)

    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    mask = bench.dense_mask_like_reference(torch, torch.cat([cond_d, unc_d]))
    print("dense mask build ms", timed(lambda: bench.dense_mask_like_reference(torch, torch.cat([cond_d, unc_d]))))
    print("descriptors_from_dense ms", timed(lambda: masks.descriptors_from_dense(mask)))
    d2 = masks.descriptors_from_dense(mask)
    print("descriptors equal", d2 == descs, d2[:2], descs[:2])
    print("layout", model._t2i_layout(cond_d, unc_d, mask, 5.0, cfg)[:5], model._t2i_layout(cond_d, unc_d, descs, 5.0, cfg)[:5])
    ids = cond_d.clone()
    print("t2i_generate(descs) ms", timed(lambda: model.t2i_generate(ids.copy_(cond_d), unc_d, descs, guidance_scale=5.0, timesteps=18, config=cfg), 3))
    print("t2i_generate(dense) ms", timed(lambda: model.t2i_generate(ids.copy_(cond_d), unc_d, mask, guidance_scale=5.0, timesteps=18, config=cfg), 3))
    print("t2i_generate(descs from dense) ms", timed(lambda: model.t2i_generate(ids.copy_(cond_d), unc_d, d2, guidance_scale=5.0, timesteps=18, config=cfg), 3))


if __name__ == "__main__":
    main()
