"""Load the UNMODIFIED reference Python (/root/reference) as the pin for the oracle.

Only usable in the build container (the GPU box has no /root/reference).  Recipe = SURVEY.md
Appendix A: register an empty `models` package whose __path__ points at the reference, pre-seed
stubs for `models.modeling_utils` (needs diffusers) and `models.misc` (needs omegaconf), then import
the reference's own modeling_showo / modeling_magvitv2 / sampling / prompting_utils files.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

import torch
import torch.nn as nn

REF = os.environ.get("SHOWO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models"))


def _install_stubs():
    if "models" in sys.modules and getattr(sys.modules["models"], "_showo_ref_stub", False):
        return
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    pkg._showo_ref_stub = True
    sys.modules["models"] = pkg

    mu = types.ModuleType("models.modeling_utils")

    class _Cfg(dict):
        __getattr__ = dict.get

        def __setattr__(self, k, v):
            self[k] = v

    class ConfigMixin:
        def register_to_config(self, **kw):
            if not hasattr(self, "_cfg"):
                object.__setattr__(self, "_cfg", _Cfg())
            self._cfg.update(kw)

        @property
        def config(self):
            if not hasattr(self, "_cfg"):
                object.__setattr__(self, "_cfg", _Cfg())
            return self._cfg

    class ModelMixin(nn.Module):
        def __getattr__(self, name):
            try:
                return super().__getattr__(name)
            except AttributeError:
                cfg = self.__dict__.get("_cfg")
                if cfg is not None and name in cfg:
                    return cfg[name]
                raise

    def register_to_config(init):
        import functools
        import inspect

        @functools.wraps(init)
        def inner(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            kw = {k: v for k, v in bound.arguments.items() if k not in ("self", "kwargs")}
            kw.update(bound.arguments.get("kwargs", {}))
            object.__setattr__(self, "_cfg", _Cfg(kw))
            init(self, *args, **kwargs)

        return inner

    mu.ConfigMixin, mu.ModelMixin, mu.register_to_config = ConfigMixin, ModelMixin, register_to_config
    sys.modules["models.modeling_utils"] = mu

    misc = types.ModuleType("models.misc")
    import typing
    for n in ("Any", "Callable", "Dict", "Iterable", "List", "NamedTuple", "NewType", "Optional", "Sized",
              "Tuple", "Type", "TypeVar", "Union"):
        setattr(misc, n, getattr(typing, n))
    misc.__all__ = [n for n in dir(misc) if not n.startswith("_")]
    sys.modules["models.misc"] = misc


def load_modules():
    """Returns SimpleNamespace(showo=models.modeling_showo, magvit=models.modeling_magvitv2, sampling=..., prompting=...)."""
    assert available(), "reference tree not present"
    _install_stubs()
    sm = importlib.import_module("models.modeling_showo")
    mv = importlib.import_module("models.modeling_magvitv2")
    sp = importlib.import_module("models.sampling")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    pu = importlib.import_module("training.prompting_utils")
    return SimpleNamespace(showo=sm, magvit=mv, sampling=sp, prompting=pu)


def build_showo(dims, weights, w_clip_vit=False):
    """Instantiate the reference Showo with PhiConfig(dims) and load `weights` (reference key names)."""
    from transformers import PhiConfig
    mods = load_modules()
    cfg = PhiConfig(vocab_size=dims.vocab_size, hidden_size=dims.hidden, intermediate_size=dims.ffn,
                    num_hidden_layers=dims.n_layers, num_attention_heads=dims.n_heads,
                    max_position_embeddings=dims.max_pos, layer_norm_eps=dims.ln_eps)
    cfg.rope_theta = dims.rope_theta
    cfg.partial_rotary_factor = dims.rotary_dim / dims.head_dim
    cfg.rope_scaling = None
    cfg._attn_implementation = "sdpa"
    mods.showo.AutoConfig.from_pretrained = staticmethod(lambda *a, **k: cfg)
    prev = torch.get_default_dtype()
    model = mods.showo.Showo(w_clip_vit=w_clip_vit, vocab_size=dims.vocab_size, llm_vocab_size=50295,
                             llm_model_path="x", codebook_size=8192, num_vq_tokens=256)
    torch.set_default_dtype(prev)
    missing, unexpected = model.load_state_dict(weights, strict=False)
    missing = [m for m in missing if "rotary_emb" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    model.eval()
    return model, mods


def t2i_config(voc):
    NS = SimpleNamespace
    return NS(model=NS(showo=NS(num_vq_tokens=voc.num_vq_tokens, num_new_special_tokens=voc.num_new_special_tokens,
                                llm_vocab_size=voc.llm_vocab_size)),
              dataset=NS(preprocessing=NS(max_seq_length=voc.max_text_len)))


def build_magvit(weights=None):
    mods = load_modules()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        vq = mods.magvit.MAGVITv2()
    if weights is not None:
        sd = vq.state_dict()
        missing = [k for k in sd if k not in weights and not k.startswith("quantize.")]
        assert not missing, missing
        vq.load_state_dict(weights, strict=False)
    vq.eval()
    return vq, mods
