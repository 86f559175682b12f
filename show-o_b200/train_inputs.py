"""Device-side producer of the training step's t2i rows (training/train.py:468-488): the reference's
`mask_or_random_replace_tokens` (training/utils.py:77-154) and `UniversalPrompting.t2i_prompt`
(training/prompting_utils.py:39-90) as ONE kernel launch behind `showo_t2i_train_prep`.

Same call shapes as the reference so that train.py's `prepare_inputs_and_labels` reads unchanged:

    input_ids, labels, loss_weight, mask_prob = mask_or_random_replace_tokens(image_tokens, mask_id, config, mask_schedule)
    input_ids, masks, labels = prompting.t2i_prompt(text_ids, input_ids, labels)

and the fused form `prompting.t2i_train_rows(text_ids, image_tokens, ...)` that does both without the [B, N] round trip and
also returns the omni-mask descriptors of the rows.  Tokenisation stays host work: `text_ids` is a list of token-id lists.
Noise: with `generator=` the uniforms are drawn by torch in the reference's order and handed to the kernel (parity mode);
otherwise the kernel's own Philox stream is seeded from torch's global generator.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from .schedules import _PowSchedule, cosine_schedule, linear_schedule

IGNORE_ID = -100


def _schedule_code(mask_schedule):
    if mask_schedule is None or mask_schedule is cosine_schedule:
        return 0, 0.0
    if mask_schedule is linear_schedule:
        return 1, 0.0
    if isinstance(mask_schedule, _PowSchedule):
        return 2, float(mask_schedule.exponent)
    return 3, 0.0                      # arbitrary callable: evaluated by the caller's torch code on the B timesteps


def _get(cfg, name, default=None):
    g = getattr(cfg, "get", None)
    return g(name, default) if callable(g) else getattr(cfg, name, default)


def _check_supported(config):
    tr = config.training
    if _get(tr, "mask_contiguous_region_prob", None) is not None:
        raise NotImplementedError("mask_contiguous_region_prob: the rectangle branch (training/utils.py:104-131) is host-side "
                                  "python `random` code in the reference; not provided on the device")
    if _get(tr, "predict_all_tokens", False) or not _get(tr, "noise_type", "mask"):
        raise NotImplementedError("only noise_type='mask' with predict_all_tokens=False (every shipped config) is provided")


def _launch(mode, image_tokens, text_ids, max_text_len, special, min_rate, drop_prob, sched, sched_param, timesteps, rand,
            drop_probs, seed, masked_in=None, labels_in=None, want_descs=True):
    lib = _lib.require_gpu()
    ref = image_tokens if image_tokens is not None else masked_in
    dev = ref.device
    if dev.type != "cuda":
        raise _lib.ShowoError("t2i training inputs are produced on the device: tensors must be CUDA (no CPU fallback)")
    B, N = ref.shape
    txt = lens = None
    stride = 0
    if mode & 2:
        stride = max(1, max(len(t) for t in text_ids))
        host = torch.zeros(B, stride, dtype=torch.int64)
        for i, t in enumerate(text_ids):
            if len(t):
                host[i, :len(t)] = torch.as_tensor(list(t), dtype=torch.int64)
        txt = host.to(dev, non_blocking=True)
        lens = torch.tensor([len(t) for t in text_ids], dtype=torch.int32).to(dev, non_blocking=True)
    L = max_text_len + 1 + N + 2
    if mode & 2:
        ids = torch.empty(B, L, dtype=torch.int64, device=dev)
        lab = torch.empty(B, L, dtype=torch.int64, device=dev)
        ones = torch.empty(B, L + 1, dtype=torch.int64, device=dev)
        descs = torch.empty(B, 5, dtype=torch.int32, device=dev) if want_descs else None
    else:
        ids = torch.empty(B, N, dtype=torch.int64, device=dev)
        lab = torch.empty(B, N, dtype=torch.int64, device=dev)
        ones = descs = None
    mp = torch.empty(B, dtype=torch.float32, device=dev) if mode & 1 else None
    sp = (C.c_int64 * 8)(*[int(v) for v in special])
    f32 = lambda t: None if t is None else t.to(dev, torch.float32).contiguous()   # noqa: E731
    timesteps, rand, drop_probs = f32(timesteps), f32(rand), f32(drop_probs)
    it = None if image_tokens is None else image_tokens.to(torch.int64).contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.showo_t2i_train_prep(_lib.ptr(it), B, N, _lib.ptr(txt), _lib.ptr(lens), stride, max_text_len, sp,
                                            float(min_rate), float(drop_prob), sched, float(sched_param), _lib.ptr(timesteps),
                                            _lib.ptr(rand), _lib.ptr(drop_probs), seed, mode, _lib.ptr(masked_in), _lib.ptr(labels_in),
                                            _lib.ptr(ids), _lib.ptr(lab), _lib.ptr(ones), _lib.ptr(descs), _lib.ptr(mp),
                                            _lib.current_stream_ptr()), "showo_t2i_train_prep")
    return ids, lab, ones, descs, mp


def _seed():
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _draw(shape, generator, dev):
    gd = generator.device if generator is not None else dev
    return torch.rand(*shape, device=gd, generator=generator)


def mask_or_random_replace_tokens(image_tokens, mask_id, config, mask_schedule, is_train=True, generator: Optional[torch.Generator] = None,
                                  noise=None):
    """training/utils.py:77-154.  Returns (input_ids [B,N], labels [B,N], loss_weight=None, mask_prob [B]).
    `noise=(timesteps [B], rand [B,N])` hands the reference's own uniform draws to the kernel (parity tests)."""
    _check_supported(config)
    if not is_train and _get(config.training, "eval_mask_ratios", None):
        raise NotImplementedError("eval_mask_ratios uses python's `random.choices` on the host in the reference")
    B, N = image_tokens.shape
    dev = image_tokens.device
    sched, param = _schedule_code(mask_schedule)
    timesteps = rand = None
    if noise is not None:
        timesteps, rand = noise
        if sched == 3:
            timesteps = mask_schedule(timesteps)
    elif generator is not None or sched == 3:
        timesteps = _draw((B,), generator, dev)
        if sched == 3:
            timesteps = mask_schedule(timesteps)
        if generator is not None:
            rand = _draw((B, N), generator, dev)
    ids, lab, _, _, mp = _launch(1, image_tokens, None, 0, (0, 0, 0, 0, 0, 0, int(mask_id), IGNORE_ID),
                                 _get(config.training, "min_masking_rate", 0.0), 0.0, sched, param, timesteps, rand, None, _seed())
    return ids, lab, None, mp


class UniversalPrompting:
    """The t2i training slice of training/prompting_utils.py:9-90: same constructor, `t2i_prompt` on the device.
    `text_tokenizer` only has to provide bos_token_id / eos_token_id / pad_token_id, convert_tokens_to_ids and __len__ like the
    reference uses it; text arrives already tokenised (lists of ids), as `__call__` hands it to `t2i_prompt` (:431-436)."""

    def __init__(self, text_tokenizer, special_tokens=("<|soi|>", "<|eoi|>", "<|sov|>", "<|eov|>", "<|t2i|>", "<|mmu|>", "<|t2v|>", "<|v2v|>", "<|lvg|>"),
                 max_text_len=8000, max_seq_len=377, ignore_id=-100, cond_dropout_prob=0.1):
        self.text_tokenizer = text_tokenizer
        self.text_tokenizer.add_special_tokens({"pad_token": "[PAD]"})
        self.text_tokenizer.add_tokens(list(special_tokens))
        self.sptids_dict = {t: torch.tensor(text_tokenizer.convert_tokens_to_ids([t])) for t in special_tokens}
        self.sptids_dict["<|sot|>"] = torch.tensor([text_tokenizer.bos_token_id])
        self.sptids_dict["<|eot|>"] = torch.tensor([text_tokenizer.eos_token_id])
        self.sptids_dict["<|pad|>"] = torch.tensor([text_tokenizer.pad_token_id])
        self.max_text_len = max_text_len + 1          # "plus 1 because at this time we add a task token before" (:33-34)
        self.pad_id = text_tokenizer.convert_tokens_to_ids("[PAD]")
        self.ignore_id = ignore_id
        self.cond_dropout_prob = cond_dropout_prob

    def _special(self, mask_id=0):
        tk = self.text_tokenizer
        return (self.pad_id, tk.bos_token_id, tk.eos_token_id, int(self.sptids_dict["<|t2i|>"]), int(self.sptids_dict["<|soi|>"]),
                int(self.sptids_dict["<|eoi|>"]), int(mask_id), self.ignore_id)

    def t2i_prompt(self, text_ids: Sequence[Sequence[int]], image_ids, labels, generator: Optional[torch.Generator] = None, probs=None):
        """:39-90 -> (input_ids [B,L], attention_masks [B,L+1] (all ones, the reference's quirk), labels [B,L]).
        `probs` = the reference's torch.rand(len(text_ids)) for parity runs."""
        if probs is None and generator is not None:
            probs = _draw((len(text_ids),), generator, image_ids.device)
        ids, lab, ones, _, _ = _launch(2, None, text_ids, self.max_text_len - 1, self._special(), 0.0, self.cond_dropout_prob, 3, 0.0,
                                       None, None, probs, _seed(), masked_in=image_ids.to(torch.int64).contiguous(),
                                       labels_in=labels.to(torch.int64).contiguous(), want_descs=False)
        return ids, ones, lab

    def t2i_train_rows(self, text_ids: Sequence[Sequence[int]], image_tokens, mask_id, config, mask_schedule=None,
                       generator: Optional[torch.Generator] = None, noise=None):
        """mask_or_random_replace_tokens + t2i_prompt fused (train.py:476-486): returns (input_ids [B,L], labels [B,L],
        mask_prob [B], descriptors int32 [B,5] on the device = create_attention_mask_predict_next of these rows in closed form)."""
        _check_supported(config)
        B, N = image_tokens.shape
        dev = image_tokens.device
        sched, param = _schedule_code(mask_schedule)
        timesteps = rand = probs = None
        if noise is not None:
            timesteps, rand, probs = noise
            if sched == 3:
                timesteps = mask_schedule(timesteps)
        elif generator is not None or sched == 3:
            timesteps = _draw((B,), generator, dev)
            if sched == 3:
                timesteps = mask_schedule(timesteps)
            if generator is not None:
                rand = _draw((B, N), generator, dev)
                probs = _draw((len(text_ids),), generator, dev)
        ids, lab, _, descs, mp = _launch(3, image_tokens, text_ids, self.max_text_len - 1, self._special(mask_id),
                                         _get(config.training, "min_masking_rate", 0.0), self.cond_dropout_prob, sched, param,
                                         timesteps, rand, probs, _seed())
        return ids, lab, mp, descs

    def t2i_gen_prompt(self, text_ids: Sequence[Sequence[int]], image_ids):
        """:92-123, the rows of the generation scripts (inference_t2i.py:115,118,250,253): left-padded
        [pad.. <|t2i|> bos text eos] (max_text_len wide, truncated with a closing eos when longer) + <|soi|> + image ids + <|eoi|>;
        an empty caption is [bos] (the unconditional row).  Returns (input_ids [B, L], text attention masks [B, max_text_len]).
        Host-side assembly on the caller's device, like the reference: this runs once per generation call, not per step."""
        tk = self.text_tokenizer
        dev = image_ids.device
        T = self.max_text_len
        rows, masks = [], []
        for i, t in enumerate(text_ids):
            t = [int(v) for v in t]
            if len(t) == 0:
                t = [tk.bos_token_id]
            elif t[0] != tk.bos_token_id:
                t = [tk.bos_token_id] + t
            ids = [int(self.sptids_dict["<|t2i|>"])] + t + [tk.eos_token_id]
            if T >= len(ids):
                n_pad = T - len(ids)
                ids = [self.pad_id] * n_pad + ids
                # the reference computes the pad count AFTER padding (T - len(padded) = 0): the mask row is all ones, kept
                msk = [1] * len(ids)
            else:
                ids = ids[:T - 1] + [tk.eos_token_id]
                msk = [1] * len(ids)
            rows.append(torch.cat([torch.tensor(ids, dtype=torch.int64, device=dev), self.sptids_dict["<|soi|>"].to(dev),
                                   image_ids[i].to(torch.int64), self.sptids_dict["<|eoi|>"].to(dev)]))
            masks.append(torch.tensor(msk, dtype=torch.int64, device=dev))
        return torch.stack(rows), torch.stack(masks)

    def __call__(self, input, task, padding=True, config=None):
        if task == "t2i":
            return self.t2i_prompt(input[0], input[1], input[2])
        if task == "t2i_gen":
            text = input[0]
            if len(text) and isinstance(text[0], str):        # :426-428: the reference tokenises here
                text = self.text_tokenizer(list(text))["input_ids"]
            return self.t2i_gen_prompt([list(t) for t in text], input[1])
        raise NotImplementedError(f"task {task!r}: only the t2i rows (training: on the device, generation: t2i_gen) are provided")
