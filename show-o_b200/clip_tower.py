"""`CLIPVisionTower` of the import surface `from models import Showo, MAGVITv2, get_mask_chedule, CLIPVisionTower`
(models/__init__.py:1-4, models/clip_encoder.py:6-82; callers inference_mmu.py:73-74,100-133, training/train_w_clip_vit.py:532-537),
running on the engine (`clip_forward`, csrc/clip.cu): patch-embedding GEMM, 23 pre-LN transformer blocks on the tcgen05 GEMM and
attention kernels, `hidden_states[-2]` with the CLS token dropped -> [B, 576, 1024] for 336 x 336 inputs.

`transformers` is used for what it is in the reference too -- reading the checkpoint / config and the host-side image processor
(PIL resize / crop / normalise) -- not for any arithmetic on the tower.  No CPU fallback: the forward needs an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib


class _Cfg(dict):
    __getattr__ = dict.get


def _dims_from_config(cfg) -> Dict[str, int]:
    return dict(image_size=int(cfg.image_size), patch_size=int(cfg.patch_size), hidden=int(cfg.hidden_size), n_layers=int(cfg.num_hidden_layers),
                n_heads=int(cfg.num_attention_heads), ffn=int(cfg.intermediate_size), ln_eps=float(getattr(cfg, "layer_norm_eps", 1e-5)))


class CLIPVisionTower(nn.Module):
    """Reference signature: CLIPVisionTower(vision_tower) with `vision_tower` a hub name / local directory (models/clip_encoder.py:7-16).
    Also accepted: a `transformers.CLIPVisionConfig` or a dict of dims (image_size, patch_size, hidden, n_layers, n_heads, ffn) -- no
    checkpoint is read then and the weights arrive through `load_weights(state_dict)` (CLIPVisionModel.state_dict() names)."""

    def __init__(self, vision_tower, *, load: bool = True):
        super().__init__()
        self.is_loaded = False
        self.select_layer = -2            # penultimate transformer block (clip_encoder.py:13)
        self.select_feature = "patch"     # drop the CLS token (:14)
        self._engine = None
        self._engine_device = None
        self.register_buffer("_anchor", torch.zeros(()), persistent=False)        # follows .to(device): where the engine lives; not in state_dict()
        self.image_processor = None
        state = None
        if isinstance(vision_tower, dict):
            self.vision_tower_name = "<dims>"
            self._dims = dict(vision_tower)
            self._dims.setdefault("ln_eps", 1e-5)
            self._hf_config = None
        elif isinstance(vision_tower, str):
            from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel
            self.vision_tower_name = vision_tower
            self._hf_config = CLIPVisionConfig.from_pretrained(vision_tower)
            self._dims = _dims_from_config(self._hf_config)
            if load:
                self.image_processor = CLIPImageProcessor.from_pretrained(vision_tower)
                state = CLIPVisionModel.from_pretrained(vision_tower).state_dict()          # checkpoint reader only
        else:                                                                                # a CLIPVisionConfig
            self.vision_tower_name = "<config>"
            self._hf_config = vision_tower
            self._dims = _dims_from_config(vision_tower)
        if self._dims["hidden"] != self._dims["n_heads"] * 64 or self._dims["hidden"] % 128 != 0:
            raise ValueError(f"CLIPVisionTower on the engine needs head_dim 64 and hidden % 128 == 0 (got {self._dims})")
        self._weights: Optional[Dict[str, torch.Tensor]] = None
        if state is not None:
            self.load_weights(state)

    # ------------------------------------------------------------------ weights
    def load_weights(self, weights: Dict[str, torch.Tensor], device=None):
        """Keep a CLIPVisionModel.state_dict() (fp32, host) and stream it into the engine when a CUDA device is known."""
        self._weights = {k: v.detach().float().contiguous() for k, v in weights.items() if k.startswith("vision_model.") and "position_ids" not in k}
        self.is_loaded = True
        if self._engine is not None:
            _lib.load().clip_engine_destroy(self._engine)
            self._engine = None
        if device is not None:
            self._anchor = self._anchor.to(device)
        if self._anchor.is_cuda:
            self._sync()
        return self

    def _sync(self):
        dev = self._anchor.device
        if dev.type != "cuda":
            raise _lib.ShowoError("CLIPVisionTower must live on a CUDA (B200) device: show-o_b200 has no CPU fallback")
        if self._engine is not None and self._engine_device == dev:
            return self._engine
        if self._weights is None:
            raise _lib.ShowoError("CLIPVisionTower: no weights (construct from a checkpoint or call load_weights)")
        lib = _lib.require_gpu()
        if self._engine is not None:
            lib.clip_engine_destroy(self._engine)
            self._engine = None
        d = self._dims
        cfg = _lib.ClipConfig(d["image_size"], d["patch_size"], d["hidden"], d["n_layers"], d["n_heads"], d["ffn"], float(d["ln_eps"]))
        h = C.c_void_p()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _lib.check(lib.clip_engine_create(C.byref(cfg), idx, C.byref(h)), "clip_engine_create")
        self._engine, self._engine_device = h, dev
        for name, t in self._weights.items():
            _lib.check(lib.clip_load_weight(h, name.encode(), _lib.ptr(t), t.numel(), int(t.is_cuda)), f"clip_load_weight({name})")
        _lib.check(lib.clip_weights_complete(h), "clip_weights_complete")
        return h

    def load_model(self, device_map=None):                   # clip_encoder.py:18-27: loading happens in __init__ here
        if self.is_loaded:
            print("{} is already loaded, `load_model` called again, skipping.".format(self.vision_tower_name))

    def __del__(self):
        try:
            if self._engine is not None:
                _lib.load().clip_engine_destroy(self._engine)
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def _features(self, images: torch.Tensor) -> torch.Tensor:
        if self.select_feature not in ("patch", "cls_patch"):
            raise ValueError(f"Unexpected select feature: {self.select_feature}")
        if images.device != self._anchor.device:
            images = images.to(self._anchor.device)          # `images.to(device=self.device, ...)` of the reference
        eng = self._sync()
        lib = _lib.require_gpu()
        x = images.float().contiguous()
        B, c, s1, s2 = x.shape
        if c != 3 or s1 != self._dims["image_size"] or s2 != self._dims["image_size"]:
            raise ValueError(f"CLIPVisionTower expects [B, 3, {self._dims['image_size']}, {self._dims['image_size']}] images, got {tuple(x.shape)}")
        drop = self.select_feature == "patch"
        T = self.num_patches + (0 if drop else 1)
        out = torch.empty(B, T, self.hidden_size, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.clip_forward(eng, _lib.ptr(x), B, int(self.select_layer), int(drop), _lib.ptr(out), _lib.current_stream_ptr()),
                       "clip_forward")
        return out

    @torch.no_grad()
    def forward(self, images):
        """clip_encoder.py:39-51: a list of [3, S, S] images or a batch [B, 3, S, S]; features come back in the input's dtype."""
        if isinstance(images, (list, tuple)):
            return [self._features(im.unsqueeze(0)).to(im.dtype) for im in images]
        return self._features(images).to(images.dtype)

    def kernel_launches(self) -> int:
        return int(_lib.load().clip_kernel_launches(self._engine)) if self._engine is not None else 0

    # ------------------------------------------------------------------ the reference's properties (:53-82)
    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return torch.float32

    @property
    def device(self):
        return self._anchor.device

    @property
    def config(self):
        if self._hf_config is not None:
            return self._hf_config
        d = self._dims
        return _Cfg(image_size=d["image_size"], patch_size=d["patch_size"], hidden_size=d["hidden"], num_hidden_layers=d["n_layers"],
                    num_attention_heads=d["n_heads"], intermediate_size=d["ffn"], layer_norm_eps=d["ln_eps"])

    @property
    def hidden_size(self):
        return self._dims["hidden"]

    @property
    def num_patches_per_side(self):
        return self._dims["image_size"] // self._dims["patch_size"]

    @property
    def num_patches(self):
        return self.num_patches_per_side ** 2
