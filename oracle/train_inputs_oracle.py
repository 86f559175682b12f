"""TEST INFRASTRUCTURE (never imported by the product path): numpy restatement of the reference's t2i training-input
producer, pinned against tests/golden/train_prep.npz (outputs of the unmodified reference, tests/golden/make_golden_prep.py).

  mask_image_tokens  <- training/utils.py:77-154   mask_or_random_replace_tokens (noise_type "mask", no contiguous region,
                                                   predict_all_tokens off)
  t2i_prompt_rows    <- training/prompting_utils.py:39-90   UniversalPrompting.t2i_prompt
"""
from __future__ import annotations

import numpy as np

IGNORE = -100


def cosine_mask_prob(timesteps: np.ndarray) -> np.ndarray:
    """models/sampling.py:39-40 on fp32: cos((t * pi) * 0.5)."""
    t = timesteps.astype(np.float32)
    return np.cos((t * np.float32(np.pi)) * np.float32(0.5)).astype(np.float32)


def mask_image_tokens(image_tokens, mask_id, mask_prob, rand, min_masking_rate=0.0):
    """utils.py:88-103,133-152: returns (input_ids, labels, mask_prob clipped)."""
    B, N = image_tokens.shape
    mp = np.maximum(mask_prob.astype(np.float32), np.float32(min_masking_rate))                  # .clip(min_masking_rate)
    n = np.maximum(np.rint(np.float32(N) * mp), 1.0)                                              # .round().clamp(min=1)
    perm = np.argsort(rand, axis=-1, kind="stable")                                              # torch.rand(B, N).argsort(-1)
    mask = perm < n[:, None]                                                                     # NB the permutation, not the rank
    ids = np.where(mask, mask_id, image_tokens)
    labels = np.where(mask, image_tokens, IGNORE)
    return ids, labels, mp


def t2i_prompt_rows(text_ids, image_ids, labels, probs, *, max_text_len, pad, bos, eos, task, soi, eoi, cond_dropout_prob):
    """prompting_utils.py:39-90 with self.max_text_len = max_text_len + 1.  Returns (ids [B,L], attention_masks [B,L+1], labels [B,L])."""
    P = max_text_len + 1
    rows, labs, masks = [], [], []
    for i, t in enumerate(text_ids):
        t = list(t)
        if len(t) == 0:
            t = [bos]
        elif t[0] != bos:
            t = [bos] + t
        tmp = [task] + t + [eos]
        if probs[i] < cond_dropout_prob:
            tmp = [task, bos, eos]
        if P >= len(tmp):
            tmp = [pad] * (P - len(tmp)) + tmp
            m = [0] * (P - len(tmp)) + [1] * (len(tmp) + image_ids.shape[-1] + 3)    # evaluated AFTER padding: all ones, L + 1 long
        else:
            tmp = tmp[:P - 1] + [eos]
            m = [1] * (len(tmp) + image_ids.shape[-1] + 3)
        lab = np.concatenate([np.asarray(tmp, dtype=np.int64), [soi], labels[i], [eoi]])
        lab = np.where(lab == pad, IGNORE, lab)
        rows.append(np.concatenate([np.asarray(tmp, dtype=np.int64), [soi], image_ids[i], [eoi]]))
        labs.append(lab)
        masks.append(np.asarray(m, dtype=np.int64))
    return np.stack(rows), np.stack(masks), np.stack(labs)


def t2i_descriptors(ids, pad, soi, eoi):
    """closed form of create_attention_mask_predict_next(rm_pad_in_image=True) for these rows (show-o_b200/masks.py)."""
    out = []
    for row in ids:
        pads = np.nonzero(row == pad)[0]
        sois, eois = np.nonzero(row == soi)[0], np.nonzero(row == eoi)[0]
        out.append((int(pads[-1]) + 1 if pads.size else 0, int(sois[0]), int(eois[-1]) + 1, 0, 0))
    return np.asarray(out, dtype=np.int32)
