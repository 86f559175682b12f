"""`CLIPVisionTower` for the import surface `from models import Showo, MAGVITv2, get_mask_chedule, CLIPVisionTower`
(models/__init__.py:1-4, used by inference_mmu.py:73-74,133).

The tower is a frozen third-party network (HF `transformers` CLIP ViT-L/14-336) that runs ONCE per image in front of the
hot path; SURVEY.md section 8 keeps it out of scope ("features synthetic", f-3 next).  This class is therefore only a thin
delegate to `transformers.CLIPVisionModel` with the reference's conventions (models/clip_encoder.py:39-51): features of the
penultimate layer, CLS token dropped -> [B, 576, 1024] for 336 x 336 inputs.  Nothing here is on the measured path.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class CLIPVisionTower(nn.Module):
    select_layer = -2            # penultimate transformer block
    select_feature = "patch"     # drop the CLS token

    def __init__(self, vision_tower, *, load: bool = True):
        """vision_tower: hub name / local directory of a CLIP vision checkpoint, or a `CLIPVisionConfig` (random init,
        used by the offline tests)."""
        super().__init__()
        from transformers import CLIPVisionConfig, CLIPVisionModel
        self.vision_tower_name = vision_tower if isinstance(vision_tower, str) else "<config>"
        if isinstance(vision_tower, CLIPVisionConfig):
            self.vision_tower = CLIPVisionModel(vision_tower)
            self.image_processor = None
        elif load:
            from transformers import CLIPImageProcessor
            self.vision_tower = CLIPVisionModel.from_pretrained(vision_tower)
            self.image_processor = CLIPImageProcessor.from_pretrained(vision_tower)
        else:
            self.vision_tower = CLIPVisionModel(CLIPVisionConfig.from_pretrained(vision_tower))
            self.image_processor = None
        self.vision_tower.requires_grad_(False)
        self.is_loaded = True

    def _select(self, out):
        feats = out.hidden_states[self.select_layer]
        return feats[:, 1:] if self.select_feature == "patch" else feats

    @torch.no_grad()
    def forward(self, images):
        if isinstance(images, (list, tuple)):
            return [self.forward(im.unsqueeze(0)) for im in images]
        out = self.vision_tower(images.to(device=self.device, dtype=self.dtype), output_hidden_states=True)
        return self._select(out).to(images.dtype)

    @property
    def dtype(self):
        return next(self.vision_tower.parameters()).dtype

    @property
    def device(self):
        return next(self.vision_tower.parameters()).device

    @property
    def config(self):
        return self.vision_tower.config

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches_per_side(self):
        return self.config.image_size // self.config.patch_size

    @property
    def num_patches(self):
        return self.num_patches_per_side ** 2

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)
