"""Reading the checkpoints the reference's `save_pretrained` / HF hub layout produces (models/modeling_utils.py): a
directory with `config.json` plus ONE of
  pytorch_model.bin | diffusion_pytorch_model.safetensors | model.safetensors
or a sharded set described by `<name>.index.json` ({"weight_map": {tensor name: shard file}}).
Host-side plumbing only: tensors come back on the CPU and are packed into the engine by `load_weights` / `_sync_engine`.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import torch

_SINGLE = ("pytorch_model.bin", "diffusion_pytorch_model.safetensors", "model.safetensors", "diffusion_pytorch_model.bin")
_INDEX = ("diffusion_pytorch_model.safetensors.index.json", "model.safetensors.index.json", "pytorch_model.bin.index.json")


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def read_config(path: str) -> dict:
    p = os.path.join(path, "config.json")
    if not os.path.exists(p):
        return {}
    return {k: v for k, v in json.load(open(p)).items() if not k.startswith("_")}


def read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """State dict of the checkpoint directory (or single file) `path`."""
    if os.path.isfile(path):
        return _load_file(path)
    for name in _INDEX:
        ip = os.path.join(path, name)
        if os.path.exists(ip):
            weight_map = json.load(open(ip))["weight_map"]
            sd: Dict[str, torch.Tensor] = {}
            for shard in sorted(set(weight_map.values())):
                sd.update(_load_file(os.path.join(path, shard)))
            missing = [k for k in weight_map if k not in sd]
            if missing:
                raise KeyError(f"{ip}: tensors listed in the index but absent from the shards: {missing[:5]}")
            return sd
    for name in _SINGLE:
        fp = os.path.join(path, name)
        if os.path.exists(fp):
            return _load_file(fp)
    raise FileNotFoundError(f"no checkpoint file under {path} (looked for {', '.join(_SINGLE + _INDEX)})")


def check_keys(expected, got, what: str, ignore_substrings=("rotary_emb.inv_freq",)):
    """Every expected key must be present; extra keys are allowed only if they match `ignore_substrings`."""
    expected, got = set(expected), set(got)
    missing = sorted(expected - got)
    extra = sorted(k for k in got - expected if not any(s in k for s in ignore_substrings))
    if missing or extra:
        raise KeyError(f"{what}: checkpoint does not match the model -- missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                       f"unexpected {extra[:5]}{'...' if len(extra) > 5 else ''}")


def write_checkpoint(path: str, state_dict: Dict[str, torch.Tensor], config: dict | None = None, max_shard_bytes: int = 5 << 30,
                     safe_serialization: bool = True):
    """The inverse: `config.json` + `diffusion_pytorch_model.safetensors` (sharded with an index above `max_shard_bytes`), or -- with
    safe_serialization=False, what training/train.py:710 asks for -- one pickled `diffusion_pytorch_model.bin`."""
    os.makedirs(path, exist_ok=True)
    if config is not None:
        json.dump(config, open(os.path.join(path, "config.json"), "w"), indent=2, sort_keys=True)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in state_dict.items()}
    if not safe_serialization:
        torch.save(sd, os.path.join(path, "diffusion_pytorch_model.bin"))
        return
    from safetensors.torch import save_file
    shards, cur, cur_bytes = [], {}, 0
    for k in sorted(sd):
        nbytes = sd[k].numel() * sd[k].element_size()
        if cur and cur_bytes + nbytes > max_shard_bytes:
            shards.append(cur); cur, cur_bytes = {}, 0
        cur[k] = sd[k]; cur_bytes += nbytes
    shards.append(cur)
    if len(shards) == 1:
        save_file(shards[0], os.path.join(path, "diffusion_pytorch_model.safetensors"))
        return
    weight_map = {}
    for i, sh in enumerate(shards):
        fn = f"diffusion_pytorch_model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(path, fn))
        weight_map.update({k: fn for k in sh})
    json.dump({"metadata": {"total_size": sum(v.numel() * v.element_size() for v in sd.values())}, "weight_map": weight_map},
              open(os.path.join(path, "diffusion_pytorch_model.safetensors.index.json"), "w"), indent=2)
