"""CPU oracle for MAGVIT-v2 encode/decode (TEST INFRASTRUCTURE ONLY — see oracle/showo_oracle.py header).

Functional fp32 restatement on torch CPU primitives (conv2d, group_norm, softmax) of
  * VQGANEncoder.forward            /root/reference/models/modeling_magvitv2.py:143-169
  * LFQuantizer.get_indices / get_codebook_entry   :186-221
  * VQGANDecoder.forward            :365-399
  * MAGVITv2.get_code / decode_code :423-433
  * ResnetBlock / AttnBlock / Upsample / Downsample / Normalize / swish
                                    /root/reference/models/common_modules.py:16-40,73-90,168-211,298-357
Weights: flat dict with the reference's state_dict key names (`encoder.down.0.block.0.norm1.weight`, ...).
Pinned against the live reference by tests/test_oracle_vs_reference.py and the fixtures in tests/golden/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class MagvitDims:
    ch: int = 128
    enc_ch_mult: List[int] = field(default_factory=lambda: [1, 2, 2, 4, 4])
    enc_res_blocks: List[int] = field(default_factory=lambda: [4, 3, 4, 3, 4])
    dec_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    dec_res_blocks: List[int] = field(default_factory=lambda: [4, 4, 3, 4, 3])
    z_channels: int = 13
    in_ch: int = 3
    out_ch: int = 3


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)                                   # common_modules.py:16-18


def group_norm(x: Tensor, W: Dict[str, Tensor], p: str) -> Tensor:
    return F.group_norm(x, 32, W[p + ".weight"], W[p + ".bias"], eps=1e-6)   # common_modules.py:21-24


def conv(x: Tensor, W: Dict[str, Tensor], p: str, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, W[p + ".weight"], W[p + ".bias"], stride=stride, padding=padding)


def resnet_block(x: Tensor, W: Dict[str, Tensor], p: str) -> Tensor:
    """common_modules.py:339-357 (temb is None, dropout 0)."""
    h = conv(swish(group_norm(x, W, p + ".norm1")), W, p + ".conv1")
    h = conv(swish(group_norm(h, W, p + ".norm2")), W, p + ".conv2")
    if (p + ".nin_shortcut.weight") in W:
        x = conv(x, W, p + ".nin_shortcut", padding=0)
    return x + h


def attn_block(x: Tensor, W: Dict[str, Tensor], p: str) -> Tensor:
    """common_modules.py:187-211: single head over HW tokens, scale C^-0.5."""
    h = group_norm(x, W, p + ".norm")
    q = conv(h, W, p + ".q", padding=0)
    k = conv(h, W, p + ".k", padding=0)
    v = conv(h, W, p + ".v", padding=0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + conv(h, W, p + ".proj_out", padding=0)


def encoder_forward(x: Tensor, W: Dict[str, Tensor], d: MagvitDims = MagvitDims()) -> Tensor:
    """VQGANEncoder.forward; returns pre-quantisation z [B,13,h,w] (after quant_conv)."""
    h = conv(x, W, "encoder.conv_in")
    nres = len(d.enc_ch_mult)
    for lvl in range(nres):
        for blk in range(d.enc_res_blocks[lvl]):
            h = resnet_block(h, W, f"encoder.down.{lvl}.block.{blk}")
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)            # common_modules.py:83-87
            h = conv(h, W, f"encoder.down.{lvl}.downsample.conv", stride=2, padding=0)
    h = resnet_block(h, W, "encoder.mid.block_1")
    h = attn_block(h, W, "encoder.mid.attn_1")
    h = resnet_block(h, W, "encoder.mid.block_2")
    h = conv(swish(group_norm(h, W, "encoder.norm_out")), W, "encoder.conv_out")
    return conv(h, W, "encoder.quant_conv", padding=0)


def lfq_indices(z: Tensor) -> Tensor:
    """sign -> 13-bit index, channel 0 = MSB (modeling_magvitv2.py:186-206,236-238).  z [B,13,h,w] -> [B,h*w] int64."""
    bits = (z > 0).long()
    pw = 2 ** torch.arange(z.shape[1] - 1, -1, -1)
    return (bits * pw.view(1, -1, 1, 1)).sum(1).reshape(z.shape[0], -1)


def lfq_entry(indices: Tensor, h: int, w: int, e_dim: int = 13) -> Tensor:
    """get_codebook_entry (modeling_magvitv2.py:208-221): index -> +-1 vectors, [B,13,h,w]."""
    b = indices.shape[0]
    bits = (indices.reshape(-1, 1) >> torch.arange(e_dim - 1, -1, -1)) & 1
    zq = bits.float() * 2 - 1
    return zq.view(b, h, w, e_dim).permute(0, 3, 1, 2).contiguous()


def get_code(pixel_values: Tensor, W: Dict[str, Tensor], d: MagvitDims = MagvitDims()) -> Tensor:
    return lfq_indices(encoder_forward(pixel_values, W, d))


def decoder_forward(z: Tensor, W: Dict[str, Tensor], d: MagvitDims = MagvitDims()) -> Tensor:
    """VQGANDecoder.forward."""
    h = conv(z, W, "decoder.post_quant_conv", padding=0)
    h = conv(h, W, "decoder.conv_in")
    h = resnet_block(h, W, "decoder.mid.block_1")
    h = attn_block(h, W, "decoder.mid.attn_1")
    h = resnet_block(h, W, "decoder.mid.block_2")
    nres = len(d.dec_ch_mult)
    for lvl in reversed(range(nres)):
        for blk in range(d.dec_res_blocks[lvl]):
            h = resnet_block(h, W, f"decoder.up.{lvl}.block.{blk}")
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")          # common_modules.py:36-40
            h = conv(h, W, f"decoder.up.{lvl}.upsample.conv")
    return conv(swish(group_norm(h, W, "decoder.norm_out")), W, "decoder.conv_out")


def decode_code(indices: Tensor, W: Dict[str, Tensor], shape=None, d: MagvitDims = MagvitDims()) -> Tensor:
    import math
    if shape is None:
        s = int(math.sqrt(indices.shape[-1]))
        shape = (s, s)
    return decoder_forward(lfq_entry(indices, shape[0], shape[1], d.z_channels), W, d)


# --------------------------------------------------------------------------- synthetic weights

def magvit_param_shapes(d: MagvitDims = MagvitDims()) -> Dict[str, tuple]:
    """Every parameter of MAGVITv2() with its shape, in the reference's state_dict naming."""
    S: Dict[str, tuple] = {}

    def conv_(p, o, i, k):
        S[p + ".weight"] = (o, i, k, k)
        S[p + ".bias"] = (o,)

    def norm_(p, c):
        S[p + ".weight"] = (c,)
        S[p + ".bias"] = (c,)

    def res_(p, i, o):
        norm_(p + ".norm1", i); conv_(p + ".conv1", o, i, 3)
        norm_(p + ".norm2", o); conv_(p + ".conv2", o, o, 3)
        if i != o:
            conv_(p + ".nin_shortcut", o, i, 1)

    def attn_(p, c):
        norm_(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv_(p + "." + n, c, c, 1)

    # encoder (modeling_magvitv2.py:59-141)
    conv_("encoder.conv_in", d.ch, d.in_ch, 3)
    in_mult = (1,) + tuple(d.enc_ch_mult)
    bi = d.ch
    for lvl in range(len(d.enc_ch_mult)):
        bi = d.ch * in_mult[lvl]
        bo = d.ch * d.enc_ch_mult[lvl]
        for blk in range(d.enc_res_blocks[lvl]):
            res_(f"encoder.down.{lvl}.block.{blk}", bi, bo)
            bi = bo
        if lvl != len(d.enc_ch_mult) - 1:
            conv_(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
    res_("encoder.mid.block_1", bi, bi); attn_("encoder.mid.attn_1", bi); res_("encoder.mid.block_2", bi, bi)
    norm_("encoder.norm_out", bi); conv_("encoder.conv_out", d.z_channels, bi, 3)
    conv_("encoder.quant_conv", d.z_channels, d.z_channels, 1)
    # decoder (:278-362)
    nres = len(d.dec_ch_mult)
    bi = d.ch * d.dec_ch_mult[nres - 1]
    conv_("decoder.conv_in", bi, d.z_channels, 3)
    res_("decoder.mid.block_1", bi, bi); attn_("decoder.mid.attn_1", bi); res_("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(nres)):
        bo = d.ch * d.dec_ch_mult[lvl]
        for blk in range(d.dec_res_blocks[lvl]):
            res_(f"decoder.up.{lvl}.block.{blk}", bi, bo)
            bi = bo
        if lvl != 0:
            conv_(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
    norm_("decoder.norm_out", bi); conv_("decoder.conv_out", d.out_ch, bi, 3)
    conv_("decoder.post_quant_conv", d.z_channels, d.z_channels, 1)
    return S


def make_magvit_weights(seed: int = 1, d: MagvitDims = MagvitDims()) -> Dict[str, Tensor]:
    """Deterministic random init (numpy Philox): conv weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like
    nn.Conv2d's default, biases likewise, GroupNorm weight 1 + N(0,0.05), bias N(0,0.05)."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(seed))
    W: Dict[str, Tensor] = {}
    shapes = magvit_param_shapes(d)
    for name, shp in shapes.items():
        if len(shp) == 4:
            bound = 1.0 / np.sqrt(shp[1] * shp[2] * shp[3])
            a = rng.uniform(-bound, bound, size=shp).astype(np.float32)
        elif ".norm" in name:
            a = rng.standard_normal(size=shp, dtype=np.float32) * 0.05
            if name.endswith(".weight"):
                a += 1.0
        else:
            wshape = shapes[name[:-5] + ".weight"]
            bound = 1.0 / np.sqrt(wshape[1] * wshape[2] * wshape[3])
            a = rng.uniform(-bound, bound, size=shp).astype(np.float32)
        W[name] = torch.from_numpy(np.ascontiguousarray(a))
    return W
