"""CPU oracle for the Show-o hot path (TEST INFRASTRUCTURE ONLY).

This file is a functional fp32 restatement, on torch CPU primitives (matmul,
softmax, exponential_/uniform_ draws), of the reference algorithm for

  * the Phi-1.5 backbone as forked by Show-o        (/root/reference/models/phi.py)
  * Showo.forward / t2i_generate / mmu_generate     (/root/reference/models/modeling_showo.py)
  * the MaskGIT sampler helpers                     (/root/reference/models/sampling.py)
  * the dense omni attention-mask builders          (/root/reference/training/prompting_utils.py)
  * the FlexAttention predicate form of the mask    (/root/reference/training/omni_attention.py)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product path (show-o_b200/) never
does: it fails loudly when the CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so
the pin is the reference's own Python, imported unmodified in the build
container by `tests/golden/ref_loader.py`; `tests/golden/make_golden.py` ran it
to produce the committed fixtures and `tests/test_oracle_vs_reference.py`
re-checks this restatement against the live reference whenever
/root/reference is present.  Third-party arithmetic under the path is torch
(reference pins torch==2.2.1, requirements.txt:195; this image has 2.11).

Weights are passed as a flat dict using the reference's state_dict key names
(`showo.model.layers.{i}.self_attn.q_proj.weight`, ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config

@dataclass
class PhiDims:
    """Backbone geometry (defaults = Phi-1.5 as instantiated by modeling_showo.py:42-47)."""
    vocab_size: int = 58498
    hidden: int = 2048
    n_layers: int = 24
    n_heads: int = 32
    ffn: int = 8192
    rotary_dim: int = 32          # partial_rotary_factor 0.5 * head_dim 64  (phi.py:277-281)
    ln_eps: float = 1e-5          # PhiConfig.layer_norm_eps                  (phi.py:744)
    rope_theta: float = 10000.0
    max_pos: int = 2048

    @property
    def head_dim(self) -> int:
        return self.hidden // self.n_heads


@dataclass
class ShowoVocab:
    """Vocabulary arithmetic of configs/showo_demo.yaml:19-24 (50295 + 10 + 8192 + 1)."""
    llm_vocab_size: int = 50295
    num_new_special_tokens: int = 10
    codebook_size: int = 8192
    num_vq_tokens: int = 256
    max_text_len: int = 128        # dataset.preprocessing.max_seq_length

    @property
    def image_offset(self) -> int:
        return self.llm_vocab_size + self.num_new_special_tokens

    @property
    def vocab_size(self) -> int:
        return self.image_offset + self.codebook_size + 1

    @property
    def mask_token_id(self) -> int:
        return self.vocab_size - 1            # modeling_showo.py:40


# special token ids with the Phi-1.5 tokenizer (prompting_utils.py:20,26-32)
BOS = EOS = 50256
PAD, SOI, EOI, SOV, EOV, T2I, MMU, T2V, V2V, LVG = range(50295, 50305)


# --------------------------------------------------------------------------- primitives

def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim (phi.py:744,776; q/k layernorm phi.py:265-271)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_new(x: Tensor) -> Tensor:
    """ACT2FN['gelu_new'] used by PhiMLP (phi.py:204)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def rotary_tables(dims: PhiDims, n_pos: int) -> tuple[Tensor, Tensor]:
    """cos/sin caches of PhiRotaryEmbedding (phi.py:79-112): emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (dims.rope_theta ** (torch.arange(0, dims.rotary_dim, 2, dtype=torch.int64).float()
                                          / dims.rotary_dim))
    t = torch.arange(n_pos, dtype=torch.int64).float()
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def apply_partial_rotary(x: Tensor, cos: Tensor, sin: Tensor, rot: int) -> Tensor:
    """x [B,H,L,dh]; rotate the first `rot` dims with rotate_half pairing (i, i+rot/2) (phi.py:163-196,680-694)."""
    xr, xp = x[..., :rot], x[..., rot:]
    half = rot // 2
    rotated = torch.cat((-xr[..., half:], xr[..., :half]), dim=-1)
    xr = xr * cos + rotated * sin
    return torch.cat((xr, xp), dim=-1)


# --------------------------------------------------------------------------- backbone

def phi_layer(x: Tensor, add_mask: Optional[Tensor], W: Dict[str, Tensor], i: int, dims: PhiDims,
              cos: Tensor, sin: Tensor) -> Tensor:
    """PhiDecoderLayer.forward (phi.py:774-790): y = Attn(LN x) + MLP(LN x) + x, one shared pre-LN."""
    p = f"showo.model.layers.{i}."
    B, L, D = x.shape
    H, dh = dims.n_heads, dims.head_dim
    xh = layer_norm(x, W[p + "input_layernorm.weight"], W[p + "input_layernorm.bias"], dims.ln_eps)

    def lin(name: str, t: Tensor) -> Tensor:
        return t @ W[p + name + ".weight"].t() + W[p + name + ".bias"]

    # PhiSdpaAttention.forward (phi.py:657-727)
    q = lin("self_attn.q_proj", xh).view(B, L, H, dh).transpose(1, 2)
    k = lin("self_attn.k_proj", xh).view(B, L, H, dh).transpose(1, 2)
    v = lin("self_attn.v_proj", xh).view(B, L, H, dh).transpose(1, 2)
    q = layer_norm(q, W[p + "self_attn.q_layernorm.weight"], W[p + "self_attn.q_layernorm.bias"], dims.ln_eps)
    k = layer_norm(k, W[p + "self_attn.k_layernorm.weight"], W[p + "self_attn.k_layernorm.bias"], dims.ln_eps)
    q = apply_partial_rotary(q, cos[:L], sin[:L], dims.rotary_dim)   # position_ids = arange(L) for every row (phi.py:998-1003)
    k = apply_partial_rotary(k, cos[:L], sin[:L], dims.rotary_dim)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if add_mask is not None:
        s = s + add_mask
    pr = torch.softmax(s, dim=-1)
    a = (pr @ v).transpose(1, 2).reshape(B, L, D)
    attn_out = lin("self_attn.dense", a)
    # PhiMLP.forward (phi.py:208-212)
    mlp_out = lin("mlp.fc2", gelu_new(lin("mlp.fc1", xh)))
    return attn_out + mlp_out + x


def showo_logits(W: Dict[str, Tensor], dims: PhiDims, input_ids: Optional[Tensor] = None,
                 input_embeddings: Optional[Tensor] = None, add_mask: Optional[Tensor] = None,
                 return_hidden: bool = False):
    """Showo.forward without labels (modeling_showo.py:76-79) -> PhiForCausalLM.forward (phi.py:1123-1183).

    add_mask: additive float mask [B,1,L,L] (0 / large negative) passed straight through (phi.py:1009-1026).
    Returns fp32 logits [B,L,V].
    """
    if input_embeddings is None:
        x = W["showo.model.embed_tokens.weight"][input_ids]
    else:
        x = input_embeddings
    L = x.shape[1]
    cos, sin = rotary_tables(dims, max(L, 1))
    hiddens = [x]
    for i in range(dims.n_layers):
        x = phi_layer(x, add_mask, W, i, dims, cos, sin)
        hiddens.append(x)
    x = layer_norm(x, W["showo.model.final_layernorm.weight"], W["showo.model.final_layernorm.bias"], dims.ln_eps)
    logits = (x @ W["showo.lm_head.weight"].t() + W["showo.lm_head.bias"]).float()
    if return_hidden:
        return logits, hiddens
    return logits


def showo_losses(logits: Tensor, labels: Tensor, batch_size_t2i: int, batch_size_lm: int, batch_size_mmu: int,
                 max_seq_length: int):
    """The three cross-entropies of Showo.forward (modeling_showo.py:81-100)."""
    import torch.nn.functional as F
    V = logits.shape[-1]
    loss_t2i = F.cross_entropy(logits[:batch_size_t2i, max_seq_length + 1:].reshape(-1, V),
                               labels[:batch_size_t2i, max_seq_length + 1:].reshape(-1), ignore_index=-100)
    sl = slice(batch_size_t2i, batch_size_t2i + batch_size_lm)
    loss_lm = F.cross_entropy(logits[sl, :-1].reshape(-1, V), labels[sl, 1:].reshape(-1), ignore_index=-100)
    loss_mmu = F.cross_entropy(logits[-batch_size_mmu:, :-1].reshape(-1, V),
                               labels[-batch_size_mmu:, 1:].reshape(-1), ignore_index=-100)
    return loss_t2i, loss_lm, loss_mmu


# --------------------------------------------------------------------------- masks

NEG_I64 = float(torch.iinfo(torch.int64).min)     # the additive value the reference produces (prompting_utils.py:503-507)


def mask_allowed_t2i(sequence: Tensor, pad_id: int = PAD, soi_id: int = SOI, eoi_id: int = EOI,
                     rm_pad_in_image: bool = True) -> Tensor:
    """Boolean 'may attend' matrix [N,L,L] of create_attention_mask_predict_next (prompting_utils.py:466-511)."""
    N, L = sequence.shape
    is_pad = sequence == pad_id
    is_soi = sequence == soi_id
    is_eoi = sequence == eoi_id
    in_img = (torch.cumsum(is_soi, 1) > torch.cumsum(is_eoi, 1)) | is_soi | is_eoi
    q = torch.arange(L)[:, None]
    k = torch.arange(L)[None, :]
    causal = (k <= q)
    allowed = torch.zeros(N, L, L, dtype=torch.bool)
    for i in range(N):
        text_rows = causal.clone()
        img_rows = torch.ones(L, L, dtype=torch.bool)
        if rm_pad_in_image:
            pads = torch.nonzero(is_pad[i]).flatten()
            if pads.numel() > 0:
                last = int(pads[-1])
                text_rows[last + 1:, :last + 1] = False          # :493-497
            sois = torch.nonzero(is_soi[i]).flatten()
            if sois.numel() > 0:
                s0 = int(sois[0])
                img_rows[s0:, is_pad[i]] = False                 # :498-500
        allowed[i] = torch.where(in_img[i][:, None], img_rows, text_rows)
    return allowed


def additive_from_allowed(allowed: Tensor) -> Tensor:
    """1.0 - mask, masked_fill(iinfo(int64).min) -> float32 [N,1,L,L] (prompting_utils.py:502-509)."""
    add = torch.zeros(allowed.shape, dtype=torch.float32)
    add[~allowed] = NEG_I64
    return add.unsqueeze(1)


def create_attention_mask_predict_next(sequence: Tensor, pad_id: int = PAD, soi_id: int = SOI, eoi_id: int = EOI,
                                       rm_pad_in_image: bool = True) -> Tensor:
    return additive_from_allowed(mask_allowed_t2i(sequence, pad_id, soi_id, eoi_id, rm_pad_in_image))


def mask_allowed_mmu(sequence: Tensor, eoi_id: int = EOI) -> Tensor:
    """create_attention_mask_for_mmu (prompting_utils.py:591-604): causal + every row sees columns <= eoi of ROW 0."""
    N, L = sequence.shape
    q = torch.arange(L)[:, None]
    k = torch.arange(L)[None, :]
    e0 = int(torch.nonzero(sequence == eoi_id)[0][1])      # torch.where(...)[1][0]: first hit in row-major order
    allowed = (k <= q) | (k <= e0)
    return allowed[None].expand(N, L, L).clone()


def create_attention_mask_for_mmu(sequence: Tensor, eoi_id: int = EOI) -> Tensor:
    return additive_from_allowed(mask_allowed_mmu(sequence, eoi_id))


def mask_allowed_mmu_vit(N: int, L: int, system_prompt_len: int = 0, n_vis: int = 576) -> Tensor:
    """create_attention_mask_for_mmu_vit (prompting_utils.py:606-624)."""
    q = torch.arange(L)[:, None]
    k = torch.arange(L)[None, :]
    b = 1 + system_prompt_len + 1
    allowed = (k <= q) | ((k >= b) & (k < b + n_vis))
    return allowed[None].expand(N, L, L).clone()


def omni_predicate(q: Tensor, k: Tensor, pad_end: int, full_begin: int, full_end: int,
                   win_begin: int, win_end: int) -> Tensor:
    """Closed form used by the CUDA kernels (per-sequence descriptor), restating the mask_mods of
    omni_attention.py:48-96: causal, OR query inside the bidirectional image span, OR key inside the
    always-visible window (mmu image prefix / vit features); minus left-pad columns for rows past the pads."""
    ok = (k <= q) | ((q >= full_begin) & (q < full_end)) | ((k >= win_begin) & (k < win_end))
    ok = ok & ~((k < pad_end) & (q >= pad_end))
    return ok


# --------------------------------------------------------------------------- sampler

def cosine_schedule(t: Tensor) -> Tensor:
    """sampling.py:39-40 (fp32 0-d tensor in, fp32 out)."""
    return torch.cos(t * math.pi * 0.5)


def log_clamped(t: Tensor, eps: float = 1e-20) -> Tensor:
    """sampling.py:10-11."""
    return torch.log(t.clamp(min=eps))


def mask_by_random_topk(mask_len: Tensor, probs: Tensor, temperature: float, uniform: Tensor) -> Tensor:
    """sampling.py:31-36 with the uniform draw supplied by the caller (`zeros_like(p).uniform_(0,1)`)."""
    gumbel = -log_clamped(-log_clamped(uniform))
    confidence = log_clamped(probs) + temperature * gumbel
    sorted_confidence = torch.sort(confidence, dim=-1).values
    cut_off = torch.gather(sorted_confidence, 1, mask_len.long())
    return confidence < cut_off


def categorical_from_exponential(probs2d: Tensor, expo: Tensor) -> Tensor:
    """torch.multinomial(p, 1) == argmax(p / q), q ~ Exp(1) drawn by empty_like(p).exponential_(1) (SURVEY 8a-12 [probed])."""
    return torch.argmax(probs2d / expo, dim=-1)


@dataclass
class StepTrace:
    input_ids_in: Tensor       # [B,L] ids fed to the model at this step (cond rows)
    logits: Tensor             # [B,N,C] post-CFG image-vocab logits
    expo: Tensor               # [B*N,C] Exp(1) noise
    uniform: Tensor            # [B,N] U(0,1) noise
    sampled_ids: Tensor        # [B,N] after where(unknown,...)
    masking: Tensor            # [B,N] bool
    mask_len: Tensor           # [B,1]
    temperature: float         # the compounded temperature handed to mask_by_random_topk


def t2i_sample_step(logits: Tensor, ids_minus: Tensor, step: int, timesteps: int, temperature: float,
                    mask_token_id: int, num_vq_tokens: int, expo: Tensor, uniform: Tensor,
                    noise_schedule: Callable = cosine_schedule):
    """One pass of the body of the denoise loop after the logits are known (modeling_showo.py:149-179).

    logits [B,N,C] fp32 (already CFG-combined and sliced); ids_minus [B,N] current codes or mask_token_id.
    Returns sampled_ids, masking, mask_len, new_temperature.
    """
    probs = logits.softmax(dim=-1)
    sampled = categorical_from_exponential(probs.reshape(-1, logits.size(-1)), expo).view(*logits.shape[:-1])
    unknown = ids_minus == mask_token_id
    sampled = torch.where(unknown, sampled, ids_minus)
    ratio = 1.0 * (step + 1) / timesteps
    mask_ratio = noise_schedule(torch.tensor(ratio))
    sel = torch.gather(probs, -1, sampled.long()[..., None]).squeeze(-1)
    sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
    mask_len = (num_vq_tokens * mask_ratio).floor().unsqueeze(0)
    mask_len = torch.max(torch.tensor([1]), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len))
    temperature = temperature * (1.0 - ratio)
    masking = mask_by_random_topk(mask_len, sel, temperature, uniform)
    return sampled, masking, mask_len, temperature


def t2i_generate(W: Dict[str, Tensor], dims: PhiDims, voc: ShowoVocab, input_ids: Tensor,
                 uncond_input_ids: Optional[Tensor], attention_mask: Tensor, temperature: float = 1.0,
                 timesteps: int = 18, guidance_scale: float = 0.0, generator: Optional[torch.Generator] = None,
                 noise_schedule: Callable = cosine_schedule, trace: Optional[List[StepTrace]] = None,
                 logits_fn: Optional[Callable] = None) -> Tensor:
    """Showo.t2i_generate (modeling_showo.py:104-181).  Mutates `input_ids` in place like the reference.

    Noise is consumed in the reference's order: per step `[B*N,C]` exponentials (torch.multinomial) then
    `[B,N]` uniforms (gumbel_noise), from `generator` (or the default generator when None).
    """
    N = voc.num_vq_tokens
    off = voc.image_offset
    mask_id = voc.mask_token_id
    P = voc.max_text_len + 1
    fwd = logits_fn or (lambda ids, m: showo_logits(W, dims, input_ids=ids, add_mask=m))
    ids_minus = input_ids[:, -(N + 1):-1].clone()
    ids_minus = torch.where(ids_minus == mask_id, mask_id, ids_minus - off)
    if uncond_input_ids is not None:
        uncond_prefix = uncond_input_ids[:, :P]
    sampled = None
    for step in range(timesteps):
        ids_in = input_ids.clone()
        if uncond_input_ids is not None and guidance_scale > 0:
            uncond_input_ids = torch.cat([uncond_prefix, input_ids[:, P:]], dim=1)
            cond, uncond = fwd(torch.cat([input_ids, uncond_input_ids]), attention_mask).chunk(2)
            logits = (1 + guidance_scale) * cond - guidance_scale * uncond
        else:
            logits = fwd(input_ids, attention_mask)
        logits = logits[:, -(N + 1):-1, off:-1]
        expo = torch.empty(logits.shape[0] * N, logits.shape[-1]).exponential_(1, generator=generator)
        uniform = torch.zeros(logits.shape[0], N).uniform_(0, 1, generator=generator)
        sampled, masking, mask_len, temperature_new = t2i_sample_step(
            logits, ids_minus, step, timesteps, temperature, mask_id, N, expo, uniform, noise_schedule)
        if trace is not None:
            trace.append(StepTrace(ids_in, logits.clone(), expo, uniform, sampled.clone(), masking.clone(),
                                   mask_len.clone(), temperature_new))
        temperature = temperature_new
        input_ids[:, -(N + 1):-1] = torch.where(masking, mask_id, sampled + off)
        ids_minus = torch.where(masking, mask_id, sampled)
    return sampled


def mmu_generate(W: Dict[str, Tensor], dims: PhiDims, idx: Tensor, attention_mask: Tensor,
                 max_new_tokens: int = 100, temperature: float = 1.0, top_k: Optional[int] = None,
                 eot_token: Optional[int] = None, generator: Optional[torch.Generator] = None) -> List[Tensor]:
    """Showo.mmu_generate, ids branch (modeling_showo.py:183-240): B must be 1, full re-forward per token."""
    result = []
    for _ in range(max_new_tokens):
        logits = showo_logits(W, dims, input_ids=idx, add_mask=attention_mask)
        L = attention_mask.shape[-1]
        am = attention_mask.squeeze()
        am_a = torch.hstack([am, torch.zeros((L, 1)) + torch.finfo(logits.dtype).min])
        am_b = torch.vstack([am_a, torch.hstack([am[-1, :], torch.tensor([0])]).unsqueeze(0)])
        attention_mask = am_b
        logits = logits[:, -1, :] / temperature
        if top_k is not None:
            v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
            logits[logits < v[:, [-1]]] = -float("Inf")
        probs = torch.softmax(logits, dim=-1)
        expo = torch.empty_like(probs).exponential_(1, generator=generator)
        idx_next = categorical_from_exponential(probs, expo)[:, None]
        result.append(idx_next[0][0])
        idx = torch.cat((idx, idx_next), dim=1)
        if eot_token is not None and int(idx_next) == eot_token:
            break
    return result


# --------------------------------------------------------------------------- synthetic weights / prompts

def _normal(rng, shape, std: float) -> Tensor:
    import numpy as np
    a = rng.standard_normal(size=shape, dtype=np.float32)
    a *= std
    return torch.from_numpy(a)


def make_showo_weights(dims: PhiDims, seed: int = 0, w_clip_vit: bool = False) -> Dict[str, Tensor]:
    """Deterministic random-init weights: Linear/Embedding ~ N(0, 0.02), biases 0, LayerNorm 1/0
    (PhiPreTrainedModel._init_weights, phi.py:833-842), drawn from numpy's Philox so that every box
    regenerates identical tensors from the seed.  Biases and LN affine terms get a
    small non-trivial perturbation (N(0,0.02)) on purpose so that parity tests exercise them."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(seed))
    D, F, V = dims.hidden, dims.ffn, dims.vocab_size
    W: Dict[str, Tensor] = {}
    W["showo.model.embed_tokens.weight"] = _normal(rng, (V, D), 0.02)
    for i in range(dims.n_layers):
        p = f"showo.model.layers.{i}."
        for name, (o, ii) in {"self_attn.q_proj": (D, D), "self_attn.k_proj": (D, D), "self_attn.v_proj": (D, D),
                              "self_attn.dense": (D, D), "mlp.fc1": (F, D), "mlp.fc2": (D, F)}.items():
            W[p + name + ".weight"] = _normal(rng, (o, ii), 0.02)
            W[p + name + ".bias"] = _normal(rng, (o,), 0.02)
        for name, n in {"input_layernorm": D, "self_attn.q_layernorm": dims.head_dim,
                        "self_attn.k_layernorm": dims.head_dim}.items():
            W[p + name + ".weight"] = 1.0 + _normal(rng, (n,), 0.02)
            W[p + name + ".bias"] = _normal(rng, (n,), 0.02)
    W["showo.model.final_layernorm.weight"] = 1.0 + _normal(rng, (D,), 0.02)
    W["showo.model.final_layernorm.bias"] = _normal(rng, (D,), 0.02)
    W["showo.lm_head.weight"] = _normal(rng, (V, D), 0.02)
    W["showo.lm_head.bias"] = _normal(rng, (V,), 0.02)
    if w_clip_vit:
        W["mm_projector.0.weight"] = _normal(rng, (2048, 1024), 0.02)
        W["mm_projector.0.bias"] = _normal(rng, (2048,), 0.02)
        W["mm_projector.2.weight"] = _normal(rng, (2048, 2048), 0.02)
        W["mm_projector.2.bias"] = _normal(rng, (2048,), 0.02)
    return W


def make_t2i_prompts(batch: int, voc: ShowoVocab, seed: int = 1234, min_len: int = 8, max_len: int = 64):
    """Synthetic t2i_gen rows (prompting_utils.py:92-123; inference_t2i.py:293-300): left-padded
    [PAD..][t2i][bos] text [eos] (max_text_len+1 wide) + [soi] + N x mask_id + [eoi]; uncond = '' prompt."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(seed))
    P = voc.max_text_len + 1
    N = voc.num_vq_tokens
    L = P + 1 + N + 1
    cond = torch.full((batch, L), PAD, dtype=torch.int64)
    uncond = torch.full((batch, L), PAD, dtype=torch.int64)
    for b in range(batch):
        n = int(rng.integers(min_len, max_len + 1))
        text = torch.from_numpy(rng.integers(0, 50257, size=n).astype("int64"))
        row = torch.cat([torch.tensor([T2I, BOS]), text, torch.tensor([EOS])])
        cond[b, P - row.numel():P] = row
        urow = torch.tensor([T2I, BOS, EOS])
        uncond[b, P - 3:P] = urow
        for r in (cond, uncond):
            r[b, P] = SOI
            r[b, P + 1:P + 1 + N] = voc.mask_token_id
            r[b, P + 1 + N] = EOI
    return cond, uncond


def make_mmu_prompts(batch: int, voc: ShowoVocab, codes: Tensor, q_len: int = 16, seed: int = 77) -> Tensor:
    """Synthetic MMU rows (inference_mmu.py:153-164): [mmu][soi] codes+offset [eoi][bos] question ids."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(seed))
    rows = []
    for b in range(batch):
        qs = torch.from_numpy(rng.integers(0, 50257, size=q_len).astype("int64"))
        rows.append(torch.cat([torch.tensor([MMU, SOI]), codes[b] + voc.image_offset,
                               torch.tensor([EOI, BOS]), qs]))
    return torch.stack(rows)
