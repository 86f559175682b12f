"""Stub models / tokenizer / config shared by make_golden_editing.py (which runs the reference's own script blocks on them) and
tests/test_host_logic.py (which runs show-o_b200/editing.py on them): deterministic, CPU-only, no arithmetic of the hot path."""
from __future__ import annotations

import numpy as np
import torch

MASK_ID = 58497
TEXT_VOCAB = 50305


class Cfg(dict):
    """attribute + .get access like the OmegaConf tree the script reads"""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class FakeTokenizer:
    """the id layout of the Show-o tokenizer with a deterministic str -> ids map"""
    bos_token_id = 50256
    eos_token_id = 50256
    pad_token_id = 50295
    _ids = {"[PAD]": 50295, "<|soi|>": 50296, "<|eoi|>": 50297, "<|sov|>": 50298, "<|eov|>": 50299, "<|t2i|>": 50300,
            "<|mmu|>": 50301, "<|t2v|>": 50302, "<|v2v|>": 50303, "<|lvg|>": 50304}

    def add_special_tokens(self, d):
        return 0

    def add_tokens(self, toks):
        return 0

    def convert_tokens_to_ids(self, t):
        return [self._ids[x] for x in t] if isinstance(t, (list, tuple)) else self._ids[t]

    def __len__(self):
        return TEXT_VOCAB

    @staticmethod
    def encode(text):
        return [(ord(c) * 131 + 7 * i) % 50000 for i, c in enumerate(text)]

    def __call__(self, texts, **kw):
        return {"input_ids": [self.encode(t) for t in texts]}


class StubVQ:
    """get_code: a deterministic function of the pixels; decode_code: records what it was asked to decode"""

    def __init__(self):
        self.decoded = []

    def get_code(self, x):
        B, _, R, _ = x.shape
        W = R // 16
        p = x.reshape(B, 3, W, 16, W, 16).mean(dim=(1, 3, 5))
        return ((p + 1.0) * 4000.0).long().clamp(0, 8191).reshape(B, W * W)

    def decode_code(self, ids, shape=None):
        self.decoded.append((ids.clone(), shape))
        n = ids.shape[1]
        h, w = shape if shape is not None else (int(n ** 0.5), int(n ** 0.5))
        return torch.zeros(ids.shape[0], 3, 16 * h, 16 * w)


class StubShowo:
    """t2i_generate: records its inputs and returns a deterministic fill that keeps the known tokens (modeling_showo.py:153-154)"""

    def __init__(self, num_vq_tokens):
        self.calls = []
        self.N = num_vq_tokens
        self.config = Cfg(mask_token_id=MASK_ID)

    def t2i_generate(self, input_ids=None, uncond_input_ids=None, attention_mask=None, **kw):
        self.calls.append(dict(input_ids=input_ids.clone(), uncond_input_ids=None if uncond_input_ids is None else uncond_input_ids.clone(),
                               attention_mask=attention_mask, kw={k: v for k, v in kw.items() if k in ("guidance_scale", "temperature", "timesteps", "seq_len", "noise_type")}))
        img = input_ids[:, -(self.N + 1):-1]
        B, N = img.shape
        fake = (torch.arange(B)[:, None] * 977 + torch.arange(N)[None, :] * 131 + 17 * len(self.calls)) % 8192
        return torch.where(img == MASK_ID, fake, img - TEXT_VOCAB)


def pixels(seed, R):
    r = np.random.Generator(np.random.Philox(seed))
    return torch.from_numpy(r.uniform(-1, 1, size=(3, R, R)).astype("float32"))


def mask_pixels(seed, R):
    """a soft-edged blob in [0, 1] so that the bicubic down-sampling + 0.5 threshold has something to decide"""
    r = np.random.Generator(np.random.Philox(seed))
    yy, xx = np.meshgrid(np.arange(R), np.arange(R), indexing="ij")
    cy, cx, rad = r.uniform(0.3, 0.7) * R, r.uniform(0.3, 0.7) * R, r.uniform(0.2, 0.35) * R
    d = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
    m = np.clip((rad - d) / 6.0 + 0.5, 0, 1) + r.uniform(-0.05, 0.05, size=(R, R))
    return torch.from_numpy(np.clip(m, 0, 1).astype("float32"))[None]


CASES = {
    "inpaint_cfg": dict(mode="inpainting", R=256, B=2, w=5.0, prompt="a red fox", seed=41),
    "inpaint_nocfg": dict(mode="inpainting", R=128, B=3, w=0.0, prompt="snow", seed=42),
    "extra_right_right": dict(mode="extrapolation", R=128, B=2, w=5.0, prompt="a lake *** a forest", direction="right *** right", offset=0, seed=43),
    "extra_left_left": dict(mode="extrapolation", R=128, B=1, w=2.0, prompt="hills *** clouds", direction="left *** left", offset=1, seed=44),
    "extra_up": dict(mode="extrapolation", R=128, B=2, w=0.0, prompt="sky", direction="up", offset=0, seed=45),
    "extra_up_up": dict(mode="extrapolation", R=128, B=1, w=5.0, prompt="sky *** stars", direction="up *** up", offset=2, seed=46),
}


def make_config(case):
    R = case["R"]
    N = (R // 16) ** 2
    return Cfg(mode=case["mode"], prompt=case["prompt"], batch_size=case["B"], guidance_scale=case["w"], generation_timesteps=4,
               image_path="image", inpainting_mask_path="mask", extra_direction=case.get("direction"), offset=case.get("offset", 0),
               dataset=Cfg(params=Cfg(resolution=R), preprocessing=Cfg(max_seq_length=128)),
               training=Cfg(batch_size=case["B"], guidance_scale=case["w"], generation_timesteps=4),
               model=Cfg(showo=Cfg(num_vq_tokens=N, codebook_size=8192, llm_vocab_size=50295, num_new_special_tokens=10)))
