"""CPU oracle of the CLIP ViT vision tower -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by show-o_b200/).

The reference's `CLIPVisionTower` (models/clip_encoder.py:6-51) is a thin wrapper around the third-party dependency
`transformers` (requirements.txt pins transformers==4.41.1; checkpoint `openai/clip-vit-large-patch14-336`), absent from
/root/reference.  This file restates the published algorithm of `transformers.models.clip.modeling_clip`
(`CLIPVisionEmbeddings`, `CLIPEncoderLayer`, `CLIPAttention`, `CLIPMLP` with `quick_gelu`, `CLIPVisionTransformer`) on plain torch
CPU ops, and tests/test_oracle_golden.py pins it to the live `transformers.CLIPVisionModel` of this image (random-init configs,
same state_dict) -- the library travels to the GPU box, so the pin runs there too.

Call sites it is anchored on: models/clip_encoder.py:29-37 (`hidden_states[select_layer = -2]`, CLS dropped for 'patch'),
:39-51 (list or batched input, cast back to the input dtype), inference_mmu.py:100-131, training/train_w_clip_vit.py:532-537.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class ClipDims:
    image_size: int = 336
    patch_size: int = 14
    hidden: int = 1024
    n_layers: int = 24
    n_heads: int = 16
    ffn: int = 4096
    ln_eps: float = 1e-5

    @property
    def n_tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


def quick_gelu(x: Tensor) -> Tensor:
    """transformers.activations.QuickGELUActivation: x * sigmoid(1.702 x)"""
    return x * torch.sigmoid(1.702 * x)


def embeddings(pixels: Tensor, W: Dict[str, Tensor], d: ClipDims) -> Tensor:
    """CLIPVisionEmbeddings.forward: stride-P convolution without bias, CLS token in front, learned positions."""
    p = "vision_model.embeddings."
    pe = F.conv2d(pixels, W[p + "patch_embedding.weight"], stride=d.patch_size).flatten(2).transpose(1, 2)
    cls = W[p + "class_embedding"].expand(pixels.shape[0], 1, -1)
    return torch.cat([cls, pe], dim=1) + W[p + "position_embedding.weight"][None]


def layer(x: Tensor, W: Dict[str, Tensor], i: int, d: ClipDims) -> Tensor:
    """CLIPEncoderLayer.forward: pre-LN attention and MLP blocks, each with its own residual."""
    p = f"vision_model.encoder.layers.{i}."
    B, T, D = x.shape
    H, dh = d.n_heads, D // d.n_heads
    h = F.layer_norm(x, (D,), W[p + "layer_norm1.weight"], W[p + "layer_norm1.bias"], d.ln_eps)
    q = F.linear(h, W[p + "self_attn.q_proj.weight"], W[p + "self_attn.q_proj.bias"]) * dh ** -0.5      # CLIPAttention scales the query
    k = F.linear(h, W[p + "self_attn.k_proj.weight"], W[p + "self_attn.k_proj.bias"])
    v = F.linear(h, W[p + "self_attn.v_proj.weight"], W[p + "self_attn.v_proj.bias"])
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, T, D)
    x = x + F.linear(a, W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), W[p + "layer_norm2.weight"], W[p + "layer_norm2.bias"], d.ln_eps)
    h = quick_gelu(F.linear(h, W[p + "mlp.fc1.weight"], W[p + "mlp.fc1.bias"]))
    return x + F.linear(h, W[p + "mlp.fc2.weight"], W[p + "mlp.fc2.bias"])


def hidden_states(pixels: Tensor, W: Dict[str, Tensor], d: ClipDims):
    """CLIPVisionTransformer.forward(output_hidden_states=True).hidden_states: [embeddings after pre_layrnorm, layer 0 out, ...]"""
    x = embeddings(pixels.float(), W, d)
    x = F.layer_norm(x, (d.hidden,), W["vision_model.pre_layrnorm.weight"], W["vision_model.pre_layrnorm.bias"], d.ln_eps)
    out = [x]
    for i in range(d.n_layers):
        x = layer(x, W, i, d)
        out.append(x)
    return out


def tower_features(pixels: Tensor, W: Dict[str, Tensor], d: ClipDims, select_layer: int = -2, select_feature: str = "patch") -> Tensor:
    """CLIPVisionTower.forward + feature_select (models/clip_encoder.py:29-51)."""
    f = hidden_states(pixels, W, d)[select_layer]
    if select_feature == "patch":
        return f[:, 1:]
    if select_feature == "cls_patch":
        return f
    raise ValueError(f"Unexpected select feature: {select_feature}")


def make_clip_weights(d: ClipDims, seed: int = 2) -> Dict[str, Tensor]:
    """Seeded weights under CLIPVisionModel.state_dict() names (numpy Philox: identical on every box).  Scales are chosen so that the
    activations stay O(1) through the depth (N(0, 0.02) projections, LayerNorm 1 + N(0, 0.02), non-zero biases on purpose)."""
    import numpy as np
    r = np.random.Generator(np.random.Philox(seed))

    def n(shape, std):
        return torch.from_numpy(r.standard_normal(size=shape, dtype=np.float32) * std)
    D, Fd, P, T = d.hidden, d.ffn, d.patch_size, d.n_tokens
    W = {"vision_model.embeddings.class_embedding": n((D,), 0.5),
         "vision_model.embeddings.patch_embedding.weight": n((D, 3, P, P), 0.05),
         "vision_model.embeddings.position_embedding.weight": n((T, D), 0.3),
         "vision_model.pre_layrnorm.weight": 1.0 + n((D,), 0.02), "vision_model.pre_layrnorm.bias": n((D,), 0.02),
         "vision_model.post_layernorm.weight": 1.0 + n((D,), 0.02), "vision_model.post_layernorm.bias": n((D,), 0.02)}
    for i in range(d.n_layers):
        p = f"vision_model.encoder.layers.{i}."
        for name, (o, ii) in {"self_attn.q_proj": (D, D), "self_attn.k_proj": (D, D), "self_attn.v_proj": (D, D), "self_attn.out_proj": (D, D),
                              "mlp.fc1": (Fd, D), "mlp.fc2": (D, Fd)}.items():
            W[p + name + ".weight"] = n((o, ii), 0.02)
            W[p + name + ".bias"] = n((o,), 0.02)
        for name in ("layer_norm1", "layer_norm2"):
            W[p + name + ".weight"] = 1.0 + n((D,), 0.02)
            W[p + name + ".bias"] = n((D,), 0.02)
    return W
