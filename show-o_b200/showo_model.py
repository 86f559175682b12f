"""`Showo`: drop-in for the reference's models.Showo (models/modeling_showo.py:23-240) on top of libshowo_b200.so.

Same constructor arguments, state_dict key names, `forward` / `t2i_generate` / `mmu_generate` signatures and return
conventions, so inference_t2i.py / inference_mmu.py can import it unchanged.  The torch modules below
are parameter containers only (they give `state_dict()`, `named_parameters()`, `.to()`, and the
`model.showo.model.embed_tokens(ids)` call the scripts make from outside); all arithmetic of the hot path runs in the
CUDA library through the C ABI.  No CPU fallback: calling a compute method without a B200 raises ShowoError.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, masks
from .schedules import cosine_schedule, step_schedule


class _Cfg(dict):
    """Minimal stand-in for diffusers' FrozenDict config (`model.config.mask_token_id`, ...)."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class _Attn(nn.Module):
    def __init__(self, d, dh, eps):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.dense = (nn.Linear(d, d) for _ in range(4))
        self.q_layernorm = nn.LayerNorm(dh, eps=eps)
        self.k_layernorm = nn.LayerNorm(dh, eps=eps)


class _Mlp(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.fc1 = nn.Linear(d, f)
        self.fc2 = nn.Linear(f, d)


class _Layer(nn.Module):
    def __init__(self, d, f, dh, eps):
        super().__init__()
        self.self_attn = _Attn(d, dh, eps)
        self.mlp = _Mlp(d, f)
        self.input_layernorm = nn.LayerNorm(d, eps=eps)


class _PhiModel(nn.Module):
    def __init__(self, v, d, f, nl, dh, eps):
        super().__init__()
        self.embed_tokens = nn.Embedding(v, d)
        self.layers = nn.ModuleList([_Layer(d, f, dh, eps) for _ in range(nl)])
        self.final_layernorm = nn.LayerNorm(d, eps=eps)


class _PhiForCausalLM(nn.Module):
    def __init__(self, v, d, f, nl, dh, eps):
        super().__init__()
        self.model = _PhiModel(v, d, f, nl, dh, eps)
        self.lm_head = nn.Linear(d, v)

    def resize_token_embeddings(self, new_v: int):
        """train.py:195 calls this; rows are copied, new rows N(0, 0.02) (phi.py:833-842)."""
        old_e, old_h = self.model.embed_tokens, self.lm_head
        d = old_e.embedding_dim
        if new_v == old_e.num_embeddings:
            return old_e
        e = nn.Embedding(new_v, d).to(old_e.weight.device, old_e.weight.dtype)
        h = nn.Linear(d, new_v).to(old_h.weight.device, old_h.weight.dtype)
        with torch.no_grad():
            e.weight.normal_(0, 0.02)
            h.weight.normal_(0, 0.02)
            h.bias.zero_()
            n = min(new_v, old_e.num_embeddings)
            e.weight[:n] = old_e.weight[:n]
            h.weight[:n] = old_h.weight[:n]
            h.bias[:n] = old_h.bias[:n]
        self.model.embed_tokens, self.lm_head = e, h
        return e


class _MMProjector(nn.Sequential):
    """Parameter container of Showo.mm_projector (modeling_showo.py:49-54) whose call runs on the engine (showo_mm_projector):
    `model.mm_projector(images_embeddings)` as inference_mmu.py:131 does it.  Under autograd (training/train_w_clip_vit.py:599-601
    trains the projector through `input_embeddings`) the call is differentiable: backward = showo_mm_projector_backward, the four
    parameters receive their gradients; the CLIP features themselves are frozen in the reference and get none."""

    def __init__(self, owner):
        super().__init__(nn.Linear(1024, 2048), nn.GELU(), nn.Linear(2048, 2048))
        object.__setattr__(self, "_owner", owner)

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if x.requires_grad:
                raise _lib.ShowoError("mm_projector: the gradient wrt the CLIP features is not provided (the vision tower is frozen in "
                                      "train_w_clip_vit.py:199-203); detach the features")
            names = ["mm_projector." + k for k, _ in self.named_parameters()]
            return _ProjectStep.apply(self._owner, x, names, *self.parameters())
        return self._owner._project(x)


class _ProjectStep(torch.autograd.Function):
    """autograd bridge of Showo.mm_projector: forward = showo_mm_projector (keeps its activations), backward =
    showo_mm_projector_backward + showo_read_grad.  Plumbing only."""

    @staticmethod
    def forward(ctx, model, x, names, *params):
        ctx.model, ctx.names, ctx.params, ctx.dtype = model, names, params, x.dtype
        return model._project(x)

    @staticmethod
    def backward(ctx, gy):
        m = ctx.model
        m.mm_projector_backward(gy)
        grads = [m.read_grad(n, like=p) if p.requires_grad else None for n, p in zip(ctx.names, ctx.params)]
        return (None, None, None, *grads)


class _TrainStep(torch.autograd.Function):
    """autograd bridge of the training step: forward = showo_train_forward, backward = showo_backward + showo_read_grad.
    (Plumbing only: no arithmetic happens in torch.)"""

    @staticmethod
    def forward(ctx, model, ids, emb, descs, labels, terms, names, *params):
        logits, losses = model.train_forward(ids, emb, descs, labels, terms)
        ctx.model, ctx.names, ctx.params = model, names, params
        ctx.emb = emb if (emb is not None and emb.requires_grad) else None
        ctx.mark_non_differentiable(logits)
        return logits, losses[0, 0], losses[1, 0], losses[2, 0]

    @staticmethod
    def backward(ctx, _g_logits, g1, g2, g3):
        m = ctx.model
        dev = m._engine_device
        g = torch.stack([x.float().reshape(()) if x is not None else torch.zeros((), device=dev) for x in (g1, g2, g3)])
        demb = m.backward(g, ctx.emb)
        grads = [m.read_grad(n, like=p) if p.requires_grad else None for n, p in zip(ctx.names, ctx.params)]
        return (None, None, demb, None, None, None, None, *grads)


class Showo(nn.Module):
    """Reference signature: Showo(w_clip_vit, vocab_size, llm_vocab_size, llm_model_path='', codebook_size=8192,
    num_vq_tokens=256, load_from_showo=True, **kwargs)   (modeling_showo.py:27-37).

    Extra keyword-only arguments select a non-default backbone geometry for tests (`phi_dims=dict(hidden=..,
    n_layers=.., n_heads=.., ffn=..)`) and `materialize=False` skips allocating the fp32 torch parameters when
    the weights will be streamed straight into the engine with `load_weights`."""

    def __init__(self, w_clip_vit, vocab_size, llm_vocab_size, llm_model_path="", codebook_size=8192,
                 num_vq_tokens=256, load_from_showo=True, *, phi_dims: Optional[dict] = None,
                 materialize: bool = True, num_new_special_tokens: int = 10, **kwargs):
        super().__init__()
        dims = dict(hidden=2048, n_layers=24, n_heads=32, ffn=8192, rotary_dim=32, max_pos=2048, ln_eps=1e-5,
                    rope_theta=10000.0)
        dims.update(phi_dims or {})
        self._dims = SimpleNamespace(**dims)
        self.vocab_size = vocab_size
        self.output_size = vocab_size
        self.w_clip_vit = w_clip_vit
        self.config = _Cfg(w_clip_vit=w_clip_vit, vocab_size=vocab_size, llm_vocab_size=llm_vocab_size,
                           llm_model_path=llm_model_path, codebook_size=codebook_size, num_vq_tokens=num_vq_tokens,
                           load_from_showo=load_from_showo, mask_token_id=vocab_size - 1)
        self._num_new_special_tokens = num_new_special_tokens
        d = self._dims
        if materialize:
            self.showo = _PhiForCausalLM(vocab_size, d.hidden, d.ffn, d.n_layers, d.hidden // d.n_heads, d.ln_eps)
            self.showo.apply(self._init_weights)
        else:
            self.showo = None
        if w_clip_vit:
            self.mm_projector = _MMProjector(self)
        self._engine = None
        self._engine_versions = None
        self._engine_device = None
        self._streamed = False

    @staticmethod
    def _init_weights(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

    @classmethod
    def from_pretrained(cls, path, **kwargs):
        """`config.json` + weights written by the reference's save_pretrained / the hub layout (single file or sharded
        safetensors / .bin; inference_t2i.py:67).  The checkpoint must carry every parameter of the model."""
        from . import checkpoint
        cfg = checkpoint.read_config(path)
        cfg.update(kwargs)
        model = cls(**cfg)
        sd = checkpoint.read_state_dict(path)
        checkpoint.check_keys(model.state_dict().keys(), sd.keys(), f"Showo.from_pretrained({path})")
        model.load_state_dict({k: v for k, v in sd.items() if "rotary_emb.inv_freq" not in k}, strict=True)
        return model

    def save_pretrained(self, path, max_shard_bytes: int = 5 << 30, safe_serialization: bool = True, **kwargs):
        """config.json + safetensors (or, safe_serialization=False like training/train.py:710, a pickled .bin) in the layout
        `from_pretrained` (and the reference's ModelMixin) reads back."""
        from . import checkpoint
        if self.showo is None:
            raise _lib.ShowoError("Showo.save_pretrained: the weights were streamed into the engine (materialize=False); "
                                  "there are no torch parameters to write")
        cfg = {"_class_name": "Showo", "w_clip_vit": bool(self.config.w_clip_vit), "vocab_size": int(self.vocab_size),
               "llm_vocab_size": int(self.config.llm_vocab_size), "llm_model_path": "", "codebook_size": int(self.config.codebook_size),
               "num_vq_tokens": int(self.config.num_vq_tokens), "load_from_showo": False}
        checkpoint.write_checkpoint(path, self.state_dict(), cfg, max_shard_bytes, safe_serialization=safe_serialization)

    # ------------------------------------------------------------------ engine plumbing
    @property
    def device(self):
        if self.showo is not None:
            return self.showo.lm_head.weight.device
        return self._engine_device or torch.device("cuda", torch.cuda.current_device())

    def _make_engine(self, device: torch.device):
        lib = _lib.require_gpu()
        d = self._dims
        cfg = _lib.Config(self.vocab_size, d.hidden, d.n_layers, d.n_heads, d.ffn, d.rotary_dim, d.max_pos, d.ln_eps,
                          d.rope_theta, self.config.llm_vocab_size, self._num_new_special_tokens,
                          self.config.codebook_size)
        h = C.c_void_p()
        _lib.check(lib.showo_engine_create(C.byref(cfg), device.index or 0, C.byref(h)), "showo_engine_create")
        self._engine, self._engine_device = h, device
        return h

    def load_weights(self, weights: Dict[str, torch.Tensor], device=None):
        """Stream a reference-format state_dict (fp32, host or device tensors) straight into the engine."""
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.require_gpu()
        if self._engine is None:
            self._make_engine(device)
        for name, t in weights.items():
            if not (name.startswith("showo.") or name.startswith("mm_projector.")):
                continue
            if "rotary_emb" in name:
                continue
            t = t.detach()
            if t.dtype != torch.float32:
                t = t.float()
            t = t.contiguous()
            _lib.check(lib.showo_load_weight(self._engine, name.encode(), _lib.ptr(t), t.numel(), int(t.is_cuda)),
                       f"showo_load_weight({name})")
        _lib.check(lib.showo_weights_complete(self._engine), "showo_weights_complete")
        self._streamed = True
        return self

    def _param_key(self, p):
        return (p._version, p.data_ptr(), p.device)

    def refresh_engine(self):
        """Force a full reload of the engine's bf16 weights from the torch parameters.  Needed after writes that autograd's
        version counter cannot see (`p.data.copy_(...)`, `p.data.add_(...)`, EMA swaps through `.data`); everything that goes
        through `load_state_dict`, `.to()`, or an in-place op on the parameter itself is picked up automatically."""
        self._engine_versions = None
        return self._sync_engine()

    def _sync_engine(self):
        """(Re)load the engine's bf16 copy of every parameter whose (version, storage) changed since the last upload
        (load_state_dict, optimizer step, .to()): only those tensors are re-packed."""
        if self.showo is None:
            if not self._streamed:
                raise _lib.ShowoError("no weights: call load_weights() or construct with materialize=True")
            return self._engine
        dev = self.showo.lm_head.weight.device
        if dev.type != "cuda":
            raise _lib.ShowoError("Showo must live on a CUDA (B200) device: show-o_b200 has no CPU fallback")
        if self._engine is not None and self._engine_device != dev:
            _lib.load().showo_engine_destroy(self._engine)
            self._engine, self._engine_versions = None, None
        keys = {"showo." + k: self._param_key(p) for k, p in self.showo.named_parameters()}
        if self.w_clip_vit:
            keys.update({"mm_projector." + k: self._param_key(p) for k, p in self.mm_projector.named_parameters()})
        prev = self._engine_versions or {}
        stale = {k: v for k, v in keys.items() if prev.get(k) != v}
        if self._engine is not None and not stale:
            return self._engine
        params = dict(self.showo.named_parameters())
        if self.w_clip_vit:
            params.update({"../mm_projector." + k: p for k, p in self.mm_projector.named_parameters()})
        with torch.cuda.device(dev):
            # the upload runs on the legacy default stream: fence it against whatever the caller's stream still has in flight
            torch.cuda.current_stream().synchronize()
            self.load_weights({k: (params[k[len("showo."):]] if k.startswith("showo.") else params["../" + k]) for k in stale}, device=dev)
            torch.cuda.synchronize(dev)
        self._engine_versions = keys
        return self._engine

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine_versions = None          # .to() / .cuda() / .float(): storages moved
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._engine_versions = None
        return out

    def __del__(self):
        try:
            if self._engine is not None:
                _lib.load().showo_engine_destroy(self._engine)
        except Exception:
            pass

    def kernel_launches(self) -> int:
        return int(_lib.load().showo_kernel_launches(self._engine)) if self._engine is not None else 0

    def _mask_descs(self, attention_mask, n_seq: int):
        if attention_mask is None:
            return masks.descriptors_causal(n_seq)
        if isinstance(attention_mask, (list, tuple)):       # already descriptors
            return list(attention_mask)
        return masks.descriptors_from_dense(attention_mask)

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _check_ids(ids, what):
        if ids.dtype != torch.int64 or not ids.is_cuda:
            raise _lib.ShowoError(f"{what}: token ids must be a CUDA int64 tensor (got {ids.dtype} on {ids.device}); "
                                  "the kernels read them as const int64_t* (out-of-range ids are clamped to [0, vocab) by the "
                                  "embedding gather where torch would raise)")
        return ids.contiguous()

    def _loss_terms(self, B, L, batch_size_t2i, batch_size_lm, batch_size_mmu, max_seq_length):
        """Row ranges of the three F.cross_entropy terms (modeling_showo.py:81-100), quirks included: logits[-batch_size_mmu:]
        is the WHOLE batch when batch_size_mmu == 0 (python's -0), while an empty t2i / lm slice gives a NaN mean."""
        P = max_seq_length + 1
        bt = min(max(batch_size_t2i, 0), B)
        lm0 = min(batch_size_t2i, B)
        lm1 = min(batch_size_t2i + batch_size_lm, B)
        mmu0 = B - batch_size_mmu if 0 < batch_size_mmu <= B else 0
        return ((0, bt, P, max(L - P, 0), 0), (lm0, max(lm1 - lm0, 0), 0, L - 1, 1), (mmu0, B - mmu0, 0, L - 1, 1))

    def forward(self, input_ids, input_embeddings=None, attention_mask=None, labels=None, label_smoothing=0.0,
                batch_size_t2i=0, batch_size_lm=0, batch_size_mmu=0, max_seq_length=128, labels_mask_text=None,
                labels_mask_image=None, **kwargs):
        """modeling_showo.py:59-102.  Returns logits fp32 [B,L,V] (and the three CE losses when labels are given).
        With labels and autograd enabled (training/train.py:589-612) the call runs the engine's training forward, which
        keeps the activations, and `loss.backward()` runs showo_backward: parameters (and `input_embeddings`) receive
        gradients; the returned logits are a non-differentiable output (the reference only uses them for logging)."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        if input_embeddings is None:
            B, L = input_ids.shape
            ids = self._check_ids(input_ids, "Showo.forward")
            emb = None
            dev = ids.device
        else:
            B, L, _ = input_embeddings.shape
            ids = None
            emb = input_embeddings.float().contiguous()
            dev = emb.device
        descs = self._mask_descs(attention_mask, B)
        if labels is not None and torch.is_grad_enabled():
            terms = self._loss_terms(B, L, batch_size_t2i, batch_size_lm, batch_size_mmu, max_seq_length)
            lab = labels.to(dev).long().contiguous()
            params = list(self.showo.named_parameters()) if self.showo is not None else []
            names = ["showo." + k for k, _ in params]
            return _TrainStep.apply(self, ids, emb, descs, lab, terms, names, *[p for _, p in params])
        logits = torch.empty(B, L, self.vocab_size, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.showo_forward(eng, _lib.ptr(ids), _lib.ptr(emb), B, L, _lib.masks_array(descs),
                                         _lib.ptr(logits), _lib.current_stream_ptr()), "showo_forward")
        if labels is not None:
            V = self.output_size
            lab = labels.to(dev).long().contiguous()
            out = torch.empty(3, 2, dtype=torch.float32, device=dev)
            terms = self._loss_terms(B, L, batch_size_t2i, batch_size_lm, batch_size_mmu, max_seq_length)
            with torch.cuda.device(dev):
                for i, (b0, nb, t0, nt, shift) in enumerate(terms):
                    _lib.check(lib.showo_cross_entropy(_lib.ptr(logits), _lib.ptr(lab), L, V, b0, nb, t0, max(nt, 0), shift, -100,
                                                       _lib.ptr(out[i]), _lib.current_stream_ptr()), "showo_cross_entropy")
            loss_t2i, loss_lm, loss_mmu = out[0, 0], out[1, 0], out[2, 0]
            return logits, loss_t2i, loss_lm, loss_mmu
        return logits

    @torch.no_grad()
    def forward_fp32(self, input_ids=None, input_embeddings=None, attention_mask=None):
        """Showo.forward without labels on the engine's fp32 VERIFICATION path (showo_forward_fp32: fp32 activations, the fp32 master
        weights, CUDA cores, one Linear at a time) -- for the parity tests' stricter claims, not a product path.  Needs
        `enable_optimizer()` (that is what makes the engine keep fp32 masters)."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        if input_embeddings is None:
            B, L = input_ids.shape
            ids, emb, dev = self._check_ids(input_ids, "Showo.forward_fp32"), None, input_ids.device
        else:
            B, L, _ = input_embeddings.shape
            ids, emb, dev = None, input_embeddings.float().contiguous(), input_embeddings.device
        descs = self._mask_descs(attention_mask, B)
        logits = torch.empty(B, L, self.vocab_size, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.showo_forward_fp32(eng, _lib.ptr(ids), _lib.ptr(emb), B, L, _lib.masks_array(descs), _lib.ptr(logits),
                                              _lib.current_stream_ptr()), "showo_forward_fp32")
        return logits

    def train_forward(self, input_ids=None, input_embeddings=None, attention_mask=None, labels=None, terms=None,
                      want_logits=True):
        """The engine's training forward without autograd plumbing (used by _TrainStep and by the benchmarks): returns
        (logits or None, losses [3, 2] = {mean, count} per term)."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        if input_embeddings is None:
            B, L = input_ids.shape
            ids, emb, dev = self._check_ids(input_ids, "Showo.train_forward"), None, input_ids.device
        elif input_ids is None:
            B, L, _ = input_embeddings.shape
            ids, emb, dev = None, input_embeddings.detach().float().contiguous(), input_embeddings.device
        else:
            # both = the mixed rows of train_w_clip_vit.py:532-537 without the torch-side embed / cat: positions with ids >= 0 are looked
            # up in the engine's table, positions with ids < 0 take input_embeddings[b, t] (the mm_projector output)
            B, L = input_ids.shape
            if tuple(input_embeddings.shape[:2]) != (B, L):
                raise ValueError(f"input_embeddings {tuple(input_embeddings.shape)} does not match input_ids {tuple(input_ids.shape)}")
            ids, emb, dev = self._check_ids(input_ids, "Showo.train_forward"), input_embeddings.detach().float().contiguous(), input_ids.device
        descs = self._mask_descs(attention_mask, B)
        lab = self._check_ids(labels, "Showo.train_forward(labels)")
        logits = torch.empty(B, L, self.vocab_size, dtype=torch.float32, device=dev) if want_logits else None
        losses = torch.empty(3, 2, dtype=torch.float32, device=dev)
        flat = (C.c_int32 * 15)(*[int(v) for t in terms for v in t])
        with torch.cuda.device(dev):
            _lib.check(lib.showo_train_forward(eng, _lib.ptr(ids), _lib.ptr(emb), B, L, _lib.masks_array(descs), _lib.ptr(lab), flat,
                                               -100, _lib.ptr(logits), _lib.ptr(losses), _lib.current_stream_ptr()), "showo_train_forward")
        return logits, losses

    def backward(self, loss_grads, want_input_grad_like=None):
        """showo_backward for loss = sum_i loss_grads[i] * loss_i of the last train_forward; returns d loss / d input_embeddings
        when `want_input_grad_like` (a [B, L, hidden] tensor) is given."""
        lib = _lib.require_gpu()
        g = torch.as_tensor(loss_grads, dtype=torch.float32, device=self._engine_device).contiguous()
        demb = torch.empty_like(want_input_grad_like, dtype=torch.float32) if want_input_grad_like is not None else None
        with torch.cuda.device(self._engine_device):
            _lib.check(lib.showo_backward(self._engine, _lib.ptr(g), _lib.ptr(demb), _lib.current_stream_ptr()), "showo_backward")
        return demb

    def backward_phase(self, phase: int, loss_grads, want_input_grad_like=None):
        """One phase of showo_backward (-1: loss / head / final LN, then layers n_layers-1 .. 0, then -2: embedding)."""
        lib = _lib.require_gpu()
        g = torch.as_tensor(loss_grads, dtype=torch.float32, device=self._engine_device).contiguous()
        demb = torch.empty_like(want_input_grad_like, dtype=torch.float32) if (want_input_grad_like is not None and phase == -2) else None
        with torch.cuda.device(self._engine_device):
            _lib.check(lib.showo_backward_phase(self._engine, int(phase), _lib.ptr(g), _lib.ptr(demb), _lib.current_stream_ptr()),
                       "showo_backward_phase")
        return demb

    def grad_buffer(self) -> torch.Tensor:
        """Zero-copy fp32 view of the engine's gradient buffer (packed weight layout) -- what a data-parallel all-reduce runs on."""
        lib = _lib.require_gpu()
        base, n = C.c_void_p(), C.c_int64()
        _lib.check(lib.showo_grad_buffer(self._engine, C.byref(base), C.byref(n)), "showo_grad_buffer")

        class _View:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<f4", "data": (int(base.value), False), "version": 2}
        return torch.as_tensor(_View(), device=self._engine_device)

    def grad_range(self, phase: int):
        lib = _lib.require_gpu()
        b, e_ = C.c_int64(), C.c_int64()
        _lib.check(lib.showo_grad_range(self._engine, int(phase), C.byref(b), C.byref(e_)), "showo_grad_range")
        return int(b.value), int(e_.value)

    def backward_overlapped(self, loss_grads, group=None, comm_stream=None, average: bool = True, input_grad_like=None,
                            projector_rows=None):
        """Data-parallel backward: runs the phases in order and all-reduces each phase's gradient range on `comm_stream` as soon
        as the phase has been enqueued (per-layer buckets; the layer's bytes move over NCCL while the next layer's backward runs).
        `input_grad_like` ([B, L, hidden]) asks for the gradient wrt the input embeddings (kept in `self.last_input_grad`);
        `projector_rows` (an index expression into it, e.g. (slice(4, 8), slice(30, 606))) names the positions the last mm_projector
        call produced: its backward runs after the embedding phase and its 6.3 M gradients are the last bucket.
        Returns the CUDA event after which every gradient is reduced."""
        import torch.distributed as dist
        dev = self._engine_device
        cur = torch.cuda.current_stream(dev)
        comm = comm_stream or torch.cuda.Stream(dev)
        G = self.grad_buffer()
        op = dist.ReduceOp.AVG if average else dist.ReduceOp.SUM
        phases = [-1] + list(range(self._dims.n_layers - 1, -1, -1)) + [-2]
        demb = None
        for ph in phases:
            out = self.backward_phase(ph, loss_grads, input_grad_like)
            demb = out if out is not None else demb
            ready = torch.cuda.Event()
            ready.record(cur)
            b, e_ = self.grad_range(ph)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                dist.all_reduce(G[b:e_], op=op, group=group)
        self.last_input_grad = demb
        if projector_rows is not None:
            if demb is None:
                raise ValueError("projector_rows needs input_grad_like")
            self.mm_projector_backward(demb[projector_rows])
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                dist.all_reduce(self.mm_projector_grad_buffer(), op=op, group=group)
        done = torch.cuda.Event()
        done.record(comm)
        return done

    # ------------------------------------------------------------------ mm_projector (training/train_w_clip_vit.py:599-601)
    def mm_projector_backward(self, grad_output: torch.Tensor):
        """showo_mm_projector_backward for the rows of the LAST mm_projector call: grad_output [..., 2048] = the gradient wrt its output
        (the projector's slice of `backward(..., want_input_grad_like=)`).  Gradients: read_grad("mm_projector.0.weight") ...,
        mm_projector_grad_buffer() for the data-parallel all-reduce; adamw_step() then updates the four tensors as well."""
        lib = _lib.require_gpu()
        g = grad_output.detach().float().contiguous()
        with torch.cuda.device(self._engine_device):
            _lib.check(lib.showo_mm_projector_backward(self._engine, _lib.ptr(g), g.numel() // 2048, _lib.current_stream_ptr()),
                       "showo_mm_projector_backward")

    def mm_projector_grad_buffer(self) -> torch.Tensor:
        """Zero-copy fp32 view of the projector's gradients [0.weight | 0.bias | 2.weight | 2.bias] (6,295,552 elements)."""
        lib = _lib.require_gpu()
        base, n = C.c_void_p(), C.c_int64()
        _lib.check(lib.showo_mm_projector_grad_buffer(self._engine, C.byref(base), C.byref(n)), "showo_mm_projector_grad_buffer")

        class _View:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<f4", "data": (int(base.value), False), "version": 2}
        return torch.as_tensor(_View(), device=self._engine_device)

    @torch.no_grad()
    def _project(self, x: torch.Tensor) -> torch.Tensor:
        """mm_projector(x): x [..., 1024] CLIP-ViT features -> [..., 2048] embeddings (inference_mmu.py:128-131)."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        xf = x.float().contiguous()
        out = torch.empty(*xf.shape[:-1], 2048, dtype=torch.float32, device=xf.device)
        with torch.cuda.device(xf.device):
            _lib.check(lib.showo_mm_projector(eng, _lib.ptr(xf), xf.numel() // 1024, _lib.ptr(out), _lib.current_stream_ptr()), "showo_mm_projector")
        return out.to(x.dtype) if x.dtype != torch.float32 else out

    # ------------------------------------------------------------------ optimizer (training/train.py:211-236, :617)
    def enable_optimizer(self, device=None):
        """Keep fp32 master weights + Adam moments in the engine.  For a materialized model every parameter is re-streamed; a model
        built with materialize=False must call this BEFORE load_weights()."""
        lib = _lib.require_gpu()
        if self._engine is None:
            dev = torch.device(device) if device is not None else (self.device if self.showo is not None else torch.device("cuda", torch.cuda.current_device()))
            self._make_engine(dev)
        _lib.check(lib.showo_optimizer_enable(self._engine), "showo_optimizer_enable")
        self._streamed = False
        self._engine_versions = None
        if self.showo is not None:
            self._sync_engine()
        return self

    def adamw_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
        """optimizer.step() on the engine's masters with the gradients of the last backward (decay on every non-bias parameter)."""
        lib = _lib.require_gpu()
        with torch.cuda.device(self._engine_device):
            _lib.check(lib.showo_adamw_step(self._engine, float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                            _lib.current_stream_ptr()), "showo_adamw_step")

    def read_param(self, name: str, like: Optional[torch.Tensor] = None, shape=None):
        """fp32 master value of one parameter (reference state_dict name) after engine-side optimizer steps."""
        lib = _lib.require_gpu()
        out = torch.empty(like.shape if like is not None else shape, dtype=torch.float32, device=self._engine_device)
        with torch.cuda.device(self._engine_device):
            _lib.check(lib.showo_read_param(self._engine, name.encode(), _lib.ptr(out), out.numel(), _lib.current_stream_ptr()),
                       f"showo_read_param({name})")
        return out

    @torch.no_grad()
    def pull_parameters(self):
        """Copy the engine's fp32 masters back into the torch parameters (before state_dict() / save_pretrained after training)."""
        for k, p in self.showo.named_parameters():
            p.copy_(self.read_param("showo." + k, like=p))
        self._engine_versions = {"showo." + k: self._param_key(p) for k, p in self.showo.named_parameters()}
        if self.w_clip_vit:
            for k, p in self.mm_projector.named_parameters():
                p.copy_(self.read_param("mm_projector." + k, like=p))
            self._engine_versions.update({"mm_projector." + k: self._param_key(p) for k, p in self.mm_projector.named_parameters()})

    def read_grad(self, name: str, like: Optional[torch.Tensor] = None, shape=None):
        """Gradient of one parameter (reference state_dict name) as a fresh fp32 tensor."""
        lib = _lib.require_gpu()
        out = torch.empty(like.shape if like is not None else shape, dtype=torch.float32, device=self._engine_device)
        with torch.cuda.device(self._engine_device):
            _lib.check(lib.showo_read_grad(self._engine, name.encode(), _lib.ptr(out), out.numel(), _lib.current_stream_ptr()),
                       f"showo_read_grad({name})")
        return out

    # ------------------------------------------------------------------ t2i
    def _t2i_layout(self, input_ids, uncond_input_ids, attention_mask, guidance_scale, config):
        N = config.model.showo.num_vq_tokens
        P = config.dataset.preprocessing.max_seq_length + 1
        B, L = input_ids.shape
        cfg_on = uncond_input_ids is not None and guidance_scale > 0
        n_seq = 2 * B if cfg_on else B
        if isinstance(attention_mask, (list, tuple)):
            descs = list(attention_mask)
        else:
            am = attention_mask
            if am is not None and am.shape[0] != n_seq:
                am = am[:n_seq]
            descs = self._mask_descs(am, n_seq)
        assert len(descs) >= n_seq, "attention_mask must cover cond (and uncond) rows"
        descs = descs[:n_seq]
        # the text prefix is step-invariant iff its rows are purely causal and nothing later is visible to them
        reusable = (L == P + N + 2) and all((d[2] <= d[1] or d[1] >= P) and d[4] <= d[3] for d in descs)
        return N, (P if reusable else 0), B, L, cfg_on, descs

    @torch.no_grad()
    def t2i_generate(self, input_ids: torch.LongTensor = None, uncond_input_ids: torch.LongTensor = None,
                     attention_mask=None, temperature=1.0, timesteps=18, guidance_scale=0,
                     noise_schedule=cosine_schedule, generator: torch.Generator = None, config=None, rng_row_offset: int = 0, **kwargs):
        """modeling_showo.py:104-181.  `input_ids` is updated in place; returns LongTensor[B, N] of codes.

        `rng_row_offset` (data parallel): global index of this call's first batch row; the kernel's Philox noise is keyed by
        (seed, global row, token, step), so with one seed on every rank a row's image does not depend on the split over GPUs.

        With `generator` given the categorical / gumbel noise is drawn by torch from it in the reference's order
        ([B*N,C] exponentials then [B,N] uniforms per step) and handed to the kernel, so a run is reproducible
        against the reference on identical logits; with generator=None the kernel uses its own Philox stream."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        N, P, B, L, cfg_on, descs = self._t2i_layout(input_ids, uncond_input_ids, attention_mask, guidance_scale, config)
        dev = input_ids.device
        C_ = self.config.codebook_size
        assert input_ids.is_contiguous() and input_ids.dtype == torch.int64
        floors, temps = step_schedule(noise_schedule, timesteps, N, float(temperature))
        floors_a = (C.c_int32 * timesteps)(*floors)
        temps_a = (C.c_float * timesteps)(*temps)
        expo = unif = None
        if generator is not None:
            expo = torch.empty(timesteps, B * N, C_, dtype=torch.float32, device=dev)
            unif = torch.empty(timesteps, B, N, dtype=torch.float32, device=dev)
            for s in range(timesteps):
                expo[s].exponential_(1, generator=generator)
                unif[s].uniform_(0, 1, generator=generator)
            seed = 0
        else:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        unc = uncond_input_ids.contiguous() if cfg_on else None
        out = torch.empty(B, N, dtype=torch.int64, device=dev)
        _lib.check(lib.showo_set_rng_row_base(eng, int(rng_row_offset)), "showo_set_rng_row_base")
        with torch.cuda.device(dev):
            _lib.check(lib.showo_t2i_generate(eng, _lib.ptr(input_ids), _lib.ptr(unc), B, L, N, P,
                                              _lib.masks_array(descs), timesteps, float(guidance_scale), floors_a,
                                              temps_a, _lib.ptr(expo), _lib.ptr(unif), seed, _lib.ptr(out),
                                              _lib.current_stream_ptr()), "showo_t2i_generate")
        return out

    @torch.no_grad()
    def t2i_step_logits(self, input_ids, uncond_input_ids=None, attention_mask=None, guidance_scale=0, config=None):
        """Parity helper: sliced logits [n_branch*B, N, C] of ONE denoise-step forward (cond rows, then uncond)."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        N, P, B, L, cfg_on, descs = self._t2i_layout(input_ids, uncond_input_ids, attention_mask, guidance_scale, config)
        nb = 2 if cfg_on else 1
        out = torch.empty(nb * B, N, self.config.codebook_size, dtype=torch.float32, device=input_ids.device)
        unc = uncond_input_ids.contiguous() if cfg_on else None
        with torch.cuda.device(input_ids.device):
            _lib.check(lib.showo_t2i_logits(eng, _lib.ptr(input_ids.contiguous()), _lib.ptr(unc), B, L, N, P,
                                            _lib.masks_array(descs), _lib.ptr(out), _lib.current_stream_ptr()),
                       "showo_t2i_logits")
        return out

    # ------------------------------------------------------------------ mmu
    @torch.no_grad()
    def mmu_generate_batched(self, idx=None, input_embeddings=None, attention_mask=None, max_new_tokens=100,
                             temperature=1.0, top_k=None, eot_token=None, generator: torch.Generator = None, rng_row_offset: int = 0):
        """Batched KV-cached decode: returns (tokens [B, max_new_tokens] int64, lengths [B] int32).
        `rng_row_offset`: global index of the first row (sampled decode, Philox noise keyed by the global row; see t2i_generate).

        top_k=1 is greedy (inference_mmu.py:81); top_k=None / k>1 draw from softmax(top-k-filtered logits / temperature)
        (modeling_showo.py:219-228).  With `generator` the Exp(1) noise of torch.multinomial is drawn by torch
        ([B, V] per token) and handed to the kernel; otherwise the kernel's Philox stream is seeded from torch's
        global generator."""
        lib = _lib.require_gpu()
        eng = self._sync_engine()
        if input_embeddings is None:
            B, L0 = idx.shape
            ids, emb, dev = self._check_ids(idx, "Showo.mmu_generate"), None, idx.device
        else:
            B, L0, _ = input_embeddings.shape
            ids, emb, dev = None, input_embeddings.float().contiguous(), input_embeddings.device
        descs = self._mask_descs(attention_mask, B)
        toks = torch.zeros(B, max_new_tokens, dtype=torch.int64, device=dev)
        lens = torch.zeros(B, dtype=torch.int32, device=dev)
        eot = -1 if eot_token is None else int(eot_token)
        k = 0 if top_k is None else int(top_k)
        expo, seed = None, 0
        if k != 1:
            if generator is not None:
                expo = torch.empty(max_new_tokens, B, self.config.vocab_size, dtype=torch.float32, device=dev)
                for t in range(max_new_tokens):
                    expo[t].exponential_(1, generator=generator)
            else:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        _lib.check(lib.showo_set_rng_row_base(eng, int(rng_row_offset)), "showo_set_rng_row_base")
        with torch.cuda.device(dev):
            _lib.check(lib.showo_mmu_generate(eng, _lib.ptr(ids), _lib.ptr(emb), B, L0, _lib.masks_array(descs),
                                              max_new_tokens, k, float(temperature), eot, seed, _lib.ptr(expo),
                                              _lib.ptr(toks), _lib.ptr(lens), _lib.current_stream_ptr()),
                       "showo_mmu_generate")
        return toks, lens

    @torch.no_grad()
    def mmu_generate(self, idx=None, input_embeddings=None, attention_mask=None, max_new_tokens=100, temperature=1.0,
                     top_k=None, eot_token=None, generator: torch.Generator = None):
        """modeling_showo.py:183-240: list of 0-d LongTensors for batch row 0, cut after eot_token.  (The reference
        only works for B == 1; for B > 1 use mmu_generate_batched.)"""
        toks, lens = self.mmu_generate_batched(idx, input_embeddings, attention_mask, max_new_tokens, temperature,
                                               top_k, eot_token, generator)
        n = int(lens[0].item())
        return [toks[0, i] for i in range(n)]
