"""GPU parity tests of the CLIP ViT vision tower on the engine (clip_forward, csrc/clip.cu) through the drop-in class
show-o_b200/clip_tower.py, against oracle/clip_oracle.py (itself pinned to the live `transformers` model in
tests/test_oracle_golden.py) on seeded weights and pixels.

Tolerance: bf16 operands with fp32 accumulation and an fp32 residual stream against an fp32 reference; the features are the
penultimate block's residual stream (std about 1 with these weights).  Bounds are stated next to each assert, the observed errors are
written to gpurun_out/parity_observed.json.
"""
import json
import os

import numpy as np
import pytest
import torch

import showo_b200
from oracle import clip_oracle as CO
from showo_b200 import CLIPVisionTower

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(key, value):
    p = os.path.join(ROOT, "gpurun_out", "parity_observed.json")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[key] = value
    json.dump(d, open(p, "w"), indent=1)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _pixels(seed, B, S):
    r = np.random.Generator(np.random.Philox(seed))
    return torch.from_numpy(r.standard_normal(size=(B, 3, S, S), dtype=np.float32))       # image-processor output is ~N(0, 1) per channel


def _tower(d, dev, seed=2):
    W = CO.make_clip_weights(d, seed=seed)
    t = CLIPVisionTower(dict(image_size=d.image_size, patch_size=d.patch_size, hidden=d.hidden, n_layers=d.n_layers, n_heads=d.n_heads, ffn=d.ffn))
    t.load_weights(W, device=dev)
    return W, t


@pytest.mark.parametrize("geo", [dict(image_size=56, patch_size=14, hidden=256, n_layers=3, n_heads=4, ffn=512),        # 17 tokens: mma.sync attention only
                                 dict(image_size=224, patch_size=14, hidden=128, n_layers=2, n_heads=2, ffn=256)])      # 257 tokens: two tcgen05 tiles + 1 tail row
def test_clip_tower_small_geometries_against_the_oracle(dev, geo):
    d = CO.ClipDims(**geo)
    W, t = _tower(d, dev)
    x = _pixels(7, 3, d.image_size)
    with torch.no_grad():
        hs = CO.hidden_states(x, W, d)
    got = t(x.to(dev)).cpu()
    ref = hs[-2][:, 1:]
    err = (got - ref).abs()
    print(f"clip {geo['image_size']}px/{d.n_layers}L: max {err.max():.4f} mean {err.mean():.5f} (ref std {ref.std():.3f})")
    _record(f"clip_small_{geo['image_size']}", {"max_abs": float(err.max()), "mean_abs": float(err.mean()), "ref_std": float(ref.std())})
    assert got.shape == ref.shape and err.max().item() < 0.025 and err.mean().item() < 0.004      # observed 0.0106 / 0.0018
    # the other selections of clip_encoder.py:29-37 and every hidden_states index
    t.select_feature = "cls_patch"
    assert (t(x.to(dev)).cpu() - hs[-2]).abs().max().item() < 0.025
    t.select_feature = "patch"
    for sel in (0, 1, -1):
        t.select_layer = sel
        assert (t(x.to(dev)).cpu() - hs[sel][:, 1:]).abs().max().item() < 0.025, sel
    t.select_layer = -2
    # list input (clip_encoder.py:41-46), dtype of the input kept, batch independence bit for bit
    fl = t([x[0].to(dev), x[1].to(dev).half()])
    assert fl[0].shape == (1, d.n_tokens - 1, d.hidden) and fl[1].dtype == torch.float16
    assert torch.equal(fl[0][0].cpu(), got[0])
    assert t.kernel_launches() > 0
    t.select_feature = "pooled"
    with pytest.raises(ValueError):
        t(x.to(dev))


def test_clip_vit_l14_336_full_size_against_the_oracle(dev):
    """openai/clip-vit-large-patch14-336 geometry (24 layers, hidden 1024, 16 heads, MLP 4096, 577 tokens), seeded weights: the features
    the MMU path consumes ([B, 576, 1024], hidden_states[-2] without CLS) vs the fp32 oracle; then through Showo.mm_projector like
    inference_mmu.py:128-131 does."""
    d = CO.ClipDims()
    W, t = _tower(d, dev)
    x = _pixels(11, 2, 336)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = CO.tower_features(x, W, d)
    got = t(x.to(dev))
    assert got.shape == (2, 576, 1024) and torch.isfinite(got).all()
    err = (got.cpu() - ref).abs()
    rel = float((got.cpu() - ref).norm() / ref.norm())
    print(f"clip ViT-L/14-336: rel L2 {rel:.5f}, max {err.max():.4f}, mean {err.mean():.5f} (ref std {ref.std():.3f}, max {ref.abs().max():.2f}); "
          f"{t.kernel_launches()} launches")
    _record("clip_vit_l14_336", {"rel_l2": rel, "max_abs": float(err.max()), "mean_abs": float(err.mean()), "ref_std": float(ref.std()),
                                 "ref_max": float(ref.abs().max()), "launches": t.kernel_launches()})
    assert rel < 0.012 and err.max().item() < 0.16          # observed rel L2 0.0058, max 0.076 (ref std 2.7, max 13.1)
    # timing of one batch of 16 images (a config-3 MMU batch), device-side
    xb = _pixels(12, 16, 336).to(dev)
    t(xb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        t(xb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flop = 16 * 23 * (577 * 2 * (4 * 1024 * 1024 + 2 * 1024 * 4096) + 4 * 577 * 577 * 1024) + 16 * 576 * 2 * 1024 * 588
    print(f"clip ViT-L/14-336, 16 images: {ms:.2f} ms = {16e3 / ms:.0f} images/s, {flop / ms / 1e9:.0f} TFLOP/s")
    _record("clip_vit_l14_336_speed", {"ms_per_16_images": ms, "tflops": flop / ms / 1e9})
