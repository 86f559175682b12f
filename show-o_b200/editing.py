"""Inpainting and extrapolation on top of the t2i hot path: the two editing modes of the reference's
inference_t2i.py (:80-164 inpainting, :166-284 extrapolation) as functions.

Both are *callers* of the same three entry points as plain t2i -- `MAGVITv2.get_code`, `Showo.t2i_generate` with partially known
image tokens (the `unknown_map` / `finfo.max` branches of modeling_showo.py:153-164) and `MAGVITv2.decode_code(shape=(h, w))` on a
non-square token grid -- plus token-grid bookkeeping that the reference keeps inline in its script.  The bookkeeping below is pinned
bit for bit to those script lines (tests/golden/make_golden_editing.py executes the script's own blocks on stub models).

The attention mask goes to `t2i_generate` in closed form (`masks.descriptors_t2i`, the descriptors of
create_attention_mask_predict_next(rm_pad_in_image=True)); a dense `[2B, 1, L, L]` tensor like the script builds is accepted by
`t2i_generate` as well.  PIL / wandb plumbing of the script is the caller's business: images arrive as tensors.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import masks as M
from .schedules import get_mask_chedule

DIRECTIONS = ("left", "right", "up", "down")


def _get(cfg, name, default=None):
    g = getattr(cfg, "get", None)
    v = g(name, None) if callable(g) else getattr(cfg, name, None)
    return default if v is None else v


def inpainting_token_mask(mask_image: torch.Tensor, resolution: int, batch_size: int) -> torch.Tensor:
    """inference_t2i.py:102-110: the [1, R, R] inpainting mask in [0, 1] (1 = repaint) is brought to the token grid by BICUBIC
    interpolation and thresholded at 0.5 -> bool [batch_size, (R / 16)^2], True = token to regenerate."""
    m = mask_image
    if m.dim() == 3:
        m = m.unsqueeze(0)
    m = F.interpolate(m, size=resolution // 16, mode="bicubic")
    m = m.repeat(batch_size, 1, 1, 1)
    return (m >= 0.5).reshape(batch_size, -1)


def t2i_gen_inputs(uni_prompting, prompts, image_tokens: torch.Tensor, guidance_scale: float):
    """The rows every mode of inference_t2i.py hands to t2i_generate (:115-131, :250-265): conditional ids, unconditional ids (empty
    caption) when guidance_scale > 0, and the omni-mask descriptors of create_attention_mask_predict_next over their concatenation."""
    input_ids, _ = uni_prompting((prompts, image_tokens), "t2i_gen")
    pad, soi, eoi = (int(uni_prompting.sptids_dict[k]) for k in ("<|pad|>", "<|soi|>", "<|eoi|>"))
    if guidance_scale > 0:
        empty = [""] * len(prompts) if (len(prompts) and isinstance(prompts[0], str)) else [[] for _ in prompts]
        uncond_input_ids, _ = uni_prompting((empty, image_tokens), "t2i_gen")
        descs = M.descriptors_t2i(torch.cat([input_ids, uncond_input_ids], dim=0), pad, soi, eoi)
    else:
        uncond_input_ids = None
        descs = M.descriptors_t2i(input_ids, pad, soi, eoi)
    return input_ids, uncond_input_ids, descs


def _mask_schedule(config):
    ms = _get(config, "mask_schedule", None)
    if ms is not None:                                       # inference_t2i.py:133-138
        return get_mask_chedule(ms.schedule, **(_get(ms, "params", None) or {}))
    return get_mask_chedule(_get(config.training, "mask_schedule", "cosine"))


def _generate(model, uni_prompting, prompts, image_tokens, config, generator=None):
    tr = config.training
    input_ids, uncond_input_ids, descs = t2i_gen_inputs(uni_prompting, prompts, image_tokens, tr.guidance_scale)
    with torch.no_grad():
        gen = model.t2i_generate(input_ids=input_ids, uncond_input_ids=uncond_input_ids, attention_mask=descs,
                                 guidance_scale=tr.guidance_scale, temperature=_get(tr, "generation_temperature", 1.0),
                                 timesteps=tr.generation_timesteps, noise_schedule=_mask_schedule(config),
                                 noise_type=_get(tr, "noise_type", "mask"), seq_len=config.model.showo.num_vq_tokens,
                                 uni_prompting=uni_prompting, config=config, generator=generator)
    return torch.clamp(gen, max=config.model.showo.codebook_size - 1, min=0)


def inpaint(model, vq_model, uni_prompting, prompts, image: torch.Tensor, mask_image: torch.Tensor, config, mask_token_id: Optional[int] = None,
            generator=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """inference_t2i.py:80-158.  image [3, R, R] in [-1, 1], mask_image [1, R, R] in [0, 1]; `prompts`: one caption (str or token-id
    list) per output image.  Returns (token ids [B, N], images [B, 3, R, R] float, decode_code's range)."""
    B = len(prompts)
    R = config.dataset.params.resolution
    mask_token_id = model.config.mask_token_id if mask_token_id is None else mask_token_id
    images = image.unsqueeze(0).repeat(B, 1, 1, 1)
    regen = inpainting_token_mask(mask_image.to(image.device), R, B)
    tokens = vq_model.get_code(images) + len(uni_prompting.text_tokenizer)
    tokens[regen] = mask_token_id
    gen = _generate(model, uni_prompting, prompts, tokens, config, generator)
    return gen, vq_model.decode_code(gen)


def extrapolation_canvas(tokens: torch.Tensor, direction: str, offset: int, mask_token_id: int, W: Optional[int] = None) -> torch.Tensor:
    """inference_t2i.py:203-226: the next [B, W, W] canvas of one extrapolation round -- the half of the current token grid that
    borders `direction` slides to the opposite side and the freed half (W/2 + offset columns or rows) is filled with mask tokens.
    `tokens` [B, h, w] carries the text-vocabulary offset already; W = resolution / 16 (default: the grid's extent across the
    direction's axis)."""
    B = tokens.shape[0]
    horizontal = direction in ("left", "right")
    if W is None:
        W = tokens.shape[1] if horizontal else tokens.shape[2]
    keep = W // 2 - offset
    if horizontal:
        blank = torch.full((B, W, W // 2 + offset), mask_token_id, dtype=torch.int64, device=tokens.device)
        return torch.cat([blank, tokens[:, :, :keep]], dim=-1) if direction == "left" else torch.cat([tokens[:, :, -keep:], blank], dim=-1)
    blank = torch.full((B, W // 2 + offset, W), mask_token_id, dtype=torch.int64, device=tokens.device)
    # the script's final `else`: any direction that is not left / right / up extends downwards
    return torch.cat([blank, tokens[:, :keep, :]], dim=-2) if direction == "up" else torch.cat([tokens[:, -keep:, :], blank], dim=-2)


def extrapolation_merge(prev_codes: torch.Tensor, gen_codes: torch.Tensor, direction: str, offset: int, W: Optional[int] = None) -> torch.Tensor:
    """inference_t2i.py:198-201,268-275: glue the freshly generated [B, W, W] canvas to the part of the previous grid that slid out of
    it.  Both in codebook ids (no text-vocabulary offset).  The grid grows by W/2 + offset along the direction's axis.
    Departure, on purpose: for the downward direction the script concatenates `image_left_part` along the row axis (:274-275), which
    raises for every input (widths differ); the evident intent -- the rows above the kept half, `image_up_part` -- is what runs here."""
    if W is None:
        W = gen_codes.shape[1] if direction in ("left", "right") else gen_codes.shape[2]
    cut = W // 2 - offset
    if direction == "left":
        return torch.cat([gen_codes, prev_codes[:, :, cut:]], dim=-1)
    if direction == "right":
        return torch.cat([prev_codes[:, :, :-cut], gen_codes], dim=-1)
    if direction == "up":
        return torch.cat([gen_codes, prev_codes[:, cut:, :]], dim=-2)
    return torch.cat([prev_codes[:, :-cut, :], gen_codes], dim=-2)


def extrapolate(model, vq_model, uni_prompting, prompts: Sequence, directions: Sequence[str], image: torch.Tensor, config, offset: int = 0,
                mask_token_id: Optional[int] = None, generator=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """inference_t2i.py:166-284.  One round per (prompt, direction) pair: the canvas keeps half of the running grid, the other half is
    generated, and the result is glued back; the final (h, w) grid is decoded with decode_code(shape=(h, w)).  `image` [3, R, R].
    Returns (codebook ids [B, h, w], images [B, 3, 16 h, 16 w]).  Like the script, consecutive rounds have to extend the same axis
    (the canvas is always W x W with W = R / 16)."""
    B = config.training.batch_size
    R = config.dataset.params.resolution
    W = R // 16
    off_txt = len(uni_prompting.text_tokenizer)
    mask_token_id = model.config.mask_token_id if mask_token_id is None else mask_token_id
    grid = None
    for rnd, (prt, direction) in enumerate(zip(prompts, directions)):
        if rnd == 0:
            codes = vq_model.get_code(image.unsqueeze(0))
            grid = codes.reshape(1, W, W).repeat(B, 1, 1)
        canvas = extrapolation_canvas(grid + off_txt, direction, offset, mask_token_id, W).reshape(B, -1)
        gen = _generate(model, uni_prompting, [prt] * B, canvas, config, generator).reshape(B, W, W)
        grid = extrapolation_merge(grid, gen, direction, offset, W)
    _, h, w = grid.shape
    return grid, vq_model.decode_code(grid.reshape(B, -1), shape=(h, w))
