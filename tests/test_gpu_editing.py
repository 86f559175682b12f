"""GPU tests of the inpainting / extrapolation modes (inference_t2i.py:80-284) end to end through the C ABI:
MAGVITv2.get_code -> token bookkeeping (show-o_b200/editing.py, pinned to the script's own lines on CPU by
tests/test_host_logic.py) -> Showo.t2i_generate with partially known image tokens -> MAGVITv2.decode_code(shape=(h, w)).

Against the oracle: the oracle runs the same flow from the same token canvas (the engine's own codes, so that LFQ sign flips of the
bf16 encoder do not enter) and every denoise step is replayed teacher-forced -- logits within the tolerance of test_gpu_parity.py,
the sampler on the oracle's logits bit-exact (this is where the known-token branches `unknown_map` / `finfo.max` of
modeling_showo.py:153-164 decide), decisions from engine logits only where the oracle's margin is below 2x the logit error.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import editing_stubs as ES  # noqa: E402
import fixtures as FX  # noqa: E402
import showo_b200  # noqa: E402
from oracle import magvit_oracle as MO  # noqa: E402
from oracle import showo_oracle as O  # noqa: E402
from showo_b200 import _lib, editing, masks as M, train_inputs as TI  # noqa: E402

pytestmark = pytest.mark.gpu
VOC = O.ShowoVocab()
TOL_TINY = 0.03


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def engines(dev):
    dims = O.PhiDims(**FX.TINY)
    W = O.make_showo_weights(dims, seed=3)
    m = showo_b200.Showo(False, dims.vocab_size, VOC.llm_vocab_size, phi_dims=FX.TINY, materialize=False)
    m.load_weights(W, device=dev)
    vq = showo_b200.MAGVITv2(materialize=False)
    vq.load_weights(MO.make_magvit_weights(1), device=dev)
    return dims, W, m, vq


def _prompting():
    return TI.UniversalPrompting(ES.FakeTokenizer(), max_text_len=128, ignore_id=-100, cond_dropout_prob=0.1)


def test_inpainting_flow_end_to_end_against_the_oracle(engines, dev):
    dims, W, m, vq = engines
    lib = _lib.require_gpu()
    case = dict(mode="inpainting", R=256, B=2, w=2.0, prompt="a red fox", seed=51)
    cfg = ES.make_config(case)
    T = cfg.training.generation_timesteps
    up = _prompting()
    image, mask_img = ES.pixels(case["seed"], 256).to(dev), ES.mask_pixels(case["seed"] + 100, 256).to(dev)
    prompts = [case["prompt"]] * case["B"]
    # (a) the flow on the engines: known tokens survive, the regenerated ones are codes, the image has the right shape
    gen, images = editing.inpaint(m, vq, up, prompts, image, mask_img, cfg, mask_token_id=VOC.mask_token_id,
                                  generator=torch.Generator(device=dev).manual_seed(9))
    codes = vq.get_code(image[None]).expand(case["B"], -1)
    regen = editing.inpainting_token_mask(mask_img, 256, case["B"])
    assert 0 < int(regen.sum()) < regen.numel()
    assert torch.equal(gen[~regen], codes[~regen]) and int(gen.min()) >= 0 and int(gen.max()) < 8192
    assert images.shape == (case["B"], 3, 256, 256) and torch.isfinite(images).all()
    assert torch.equal(regen.cpu(), editing.inpainting_token_mask(mask_img.cpu(), 256, case["B"]))     # bicubic + threshold: same grid on both devices
    # (b) the oracle runs the same canvas; every step replayed teacher-forced
    tokens = (codes + len(up.text_tokenizer)).cpu().clone()
    tokens[regen.cpu()] = VOC.mask_token_id
    cond, uncond, descs = editing.t2i_gen_inputs(up, prompts, tokens, case["w"])
    dense = O.create_attention_mask_predict_next(torch.cat([cond, uncond]))
    assert M.descriptors_from_dense(dense) == descs            # the closed form the flow hands over == the script's dense mask
    trace = []
    with torch.no_grad():
        ref_gen = O.t2i_generate(W, dims, VOC, cond.clone(), uncond.clone(), dense, guidance_scale=case["w"], timesteps=T,
                                 generator=torch.Generator().manual_seed(21), trace=trace)
    assert torch.equal(ref_gen[~regen.cpu()], codes.cpu()[~regen.cpu()])
    floors, temps = showo_b200.step_schedule(showo_b200.cosine_schedule, T, 256, 1.0)
    B, w = case["B"], case["w"]
    worst = 0.0
    for s, tr in enumerate(trace):
        ids_in = tr.input_ids_in.to(dev)
        sl = m.t2i_step_logits(ids_in, uncond.to(dev), descs, guidance_scale=w, config=cfg)
        lg = ((1 + w) * sl[:B] - w * sl[B:]).cpu()
        err = (lg - tr.logits).abs().max().item()
        worst = max(worst, err)
        assert err < (1 + 2 * w) * TOL_TINY, (s, err)
        ex, un = tr.expo.to(dev), tr.uniform.to(dev)
        for lc, lu, ww, exact in ((tr.logits.contiguous().to(dev), None, 0.0, True), (sl[:B].contiguous(), sl[B:].contiguous(), w, False)):
            ids_d = ids_in.clone()
            out = torch.zeros(B, 256, dtype=torch.int64, device=dev)
            mk = torch.zeros(B, 256, dtype=torch.uint8, device=dev)
            _lib.check(lib.showo_sampler_step(_lib.ptr(lc), _lib.ptr(lu), B, 256, 8192, ww, _lib.ptr(ids_d), 387, 130, VOC.image_offset,
                                              VOC.mask_token_id, floors[s], temps[s], _lib.ptr(ex), _lib.ptr(un), 0, s, _lib.ptr(out),
                                              _lib.ptr(mk), _lib.current_stream_ptr()))
            if exact:
                assert torch.equal(out.cpu(), tr.sampled_ids) and torch.equal(mk.cpu().bool(), tr.masking), s
            else:
                race = (tr.logits.reshape(-1, 8192) - torch.log(tr.expo)).topk(2, -1).values
                margin = (race[:, 0] - race[:, 1]).view(B, 256)
                diff = out.cpu() != tr.sampled_ids
                assert (margin[diff] <= 2 * err).all(), (s, margin[diff], err)
    print(f"inpainting flow: {int(regen[0].sum())} of 256 tokens regenerated per image, worst combined-logit error {worst:.4f} over {T} steps")


def test_extrapolation_flow_end_to_end(engines, dev):
    """two rounds to the right (inference_t2i.py:166-284): the grid grows 16x16 -> 16x24 -> 16x32, the original image's tokens are
    untouched in the first 16 columns (kept halves are known tokens of t2i_generate), and decode_code(shape=(16, 32)) gives a
    256 x 512 image that matches the oracle's decoder on the same grid."""
    dims, W, m, vq = engines
    case = dict(mode="extrapolation", R=256, B=2, w=0.0, prompt="a lake *** a forest", direction="right *** right", offset=0, seed=52)
    cfg = ES.make_config(case)
    cfg.training.generation_timesteps = 3
    up = _prompting()
    image = ES.pixels(case["seed"], 256).to(dev)
    grid, images = editing.extrapolate(m, vq, up, case["prompt"].split(" *** "), case["direction"].split(" *** "), image, cfg, offset=0,
                                       mask_token_id=VOC.mask_token_id, generator=torch.Generator(device=dev).manual_seed(4))
    codes = vq.get_code(image[None]).reshape(1, 16, 16)
    assert grid.shape == (2, 16, 32) and int(grid.min()) >= 0 and int(grid.max()) < 8192
    assert torch.equal(grid[:, :, :16], codes.expand(2, -1, -1))
    assert images.shape == (2, 3, 256, 512) and torch.isfinite(images).all()
    with torch.no_grad():
        ref = MO.decode_code(grid[:1].reshape(1, -1).cpu(), MO.make_magvit_weights(1), shape=(16, 32))
    d = (images[:1].cpu() - ref).abs()
    print(f"extrapolation: decode of the 16 x 32 grid vs oracle: max {d.max():.4f} mean {d.mean():.5f}")
    assert d.mean().item() < 0.012 and d.max().item() < 0.25
    # upwards with an offset: 16x16 -> (16 + 8 + 2) x 16; to the left with an odd offset: 16 x 25 -- grids that are no multiple of the
    # convolution kernel's 8 x 16 pixel tile, decoded against the oracle
    grid_u, img_u = editing.extrapolate(m, vq, up, ["sky"], ["up"], image, cfg, offset=2, mask_token_id=VOC.mask_token_id)
    assert grid_u.shape == (2, 26, 16) and img_u.shape == (2, 3, 416, 256)
    assert torch.equal(grid_u[:, 10:, :], codes.expand(2, -1, -1))
    grid_l, img_l = editing.extrapolate(m, vq, up, ["hills"], ["left"], image, cfg, offset=1, mask_token_id=VOC.mask_token_id)
    assert grid_l.shape == (2, 16, 25) and img_l.shape == (2, 3, 256, 400)
    assert torch.equal(grid_l[:, :, 9:], codes.expand(2, -1, -1))
    for g_, im_ in ((grid_u, img_u), (grid_l, img_l)):
        with torch.no_grad():
            ref = MO.decode_code(g_[:1].reshape(1, -1).cpu(), MO.make_magvit_weights(1), shape=tuple(g_.shape[1:]))
        d = (im_[:1].cpu() - ref).abs()
        print(f"extrapolation: decode of the {g_.shape[1]} x {g_.shape[2]} grid vs oracle: max {d.max():.4f} mean {d.mean():.5f}")
        assert d.mean().item() < 0.012 and d.max().item() < 0.25
