"""Omni attention mask: dense additive tensor  <->  per-sequence closed form (showo_seq_mask_t).

The reference builds a dense float `[B,1,L,L]` tensor (training/prompting_utils.py:466-511,591-624) and feeds it
to SDPA; the kernels here evaluate the equivalent predicate of training/omni_attention.py:48-96 in registers:

    allowed(q,k) = (k <= q  or  full_begin <= q < full_end  or  win_begin <= k < win_end)
                   and not (k < pad_end and q >= pad_end)

`descriptors_from_dense` recovers the five integers per sequence from the caller's dense mask and VERIFIES them
against it on every non-pad query row (pad rows' outputs are never read by any other row); a mask
that the closed form cannot express raises instead of being silently approximated.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

Desc = Tuple[int, int, int, int, int]


def predicate(L: int, desc: Desc, device=None) -> torch.Tensor:
    pad_end, fb, fe, wb, we = desc
    q = torch.arange(L, device=device)[:, None]
    k = torch.arange(L, device=device)[None, :]
    ok = (k <= q) | ((q >= fb) & (q < fe)) | ((k >= wb) & (k < we))
    return ok & ~((k < pad_end) & (q >= pad_end))


def descriptors_t2i(input_ids: torch.Tensor, pad_id: int, soi_id: int, eoi_id: int) -> List[Desc]:
    """Closed form of create_attention_mask_predict_next(rm_pad_in_image=True) for left-padded rows with one image span."""
    ids = input_ids.detach().cpu()
    out = []
    for row in ids:
        pads = torch.nonzero(row == pad_id).flatten()
        pad_end = int(pads[-1]) + 1 if pads.numel() else 0
        sois = torch.nonzero(row == soi_id).flatten()
        eois = torch.nonzero(row == eoi_id).flatten()
        if sois.numel():
            fb = int(sois[0])
            fe = int(eois[-1]) + 1 if eois.numel() else row.numel()
        else:
            fb = fe = 0
        out.append((pad_end, fb, fe, 0, 0))
    return out


def descriptors_mmu(input_ids: torch.Tensor, eoi_id: int) -> List[Desc]:
    """create_attention_mask_for_mmu: every row sees columns <= eoi position OF ROW 0 (prompting_utils.py:594-595)."""
    ids = input_ids.detach().cpu()
    e0 = int(torch.nonzero(ids == eoi_id)[0][1])
    return [(0, 0, 0, 0, e0 + 1) for _ in range(ids.shape[0])]


def descriptors_mmu_vit(batch: int, system_prompt_len: int = 0, n_vis: int = 576) -> List[Desc]:
    b = 1 + system_prompt_len + 1
    return [(0, 0, 0, b, b + n_vis) for _ in range(batch)]


def descriptors_causal(batch: int) -> List[Desc]:
    return [(0, 0, 0, 0, 0) for _ in range(batch)]


def _descriptors_from_dense_cuda(m: torch.Tensor, verify: bool) -> List[Desc]:
    """CUDA tensors: one kernel (showo_mask_descriptors) derives and verifies the descriptors; 24 bytes per sequence come back."""
    import ctypes as C
    from . import _lib
    lib = _lib.require_gpu()
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    elif m.dtype not in (torch.float32, torch.uint8):
        m = m.float()
    if m.stride(-1) != 1 or m.stride(-2) != m.shape[-1]:
        m = m.contiguous()
    B, L, _ = m.shape
    descs = (_lib.SeqMask * B)()
    bad = (C.c_int32 * B)()
    with torch.cuda.device(m.device):
        _lib.check(lib.showo_mask_descriptors(_lib.ptr(m), m.element_size(), B, L, m.stride(0), descs, bad, _lib.current_stream_ptr()),
                   "showo_mask_descriptors")
    out = [(d.pad_end, d.full_begin, d.full_end, d.win_begin, d.win_end) for d in descs]
    if verify:
        for b in range(B):
            if bad[b]:
                raise NotImplementedError(
                    f"attention_mask row {b} is not an omni mask (causal / image-span / window / left-pad): "
                    f"derived descriptor {out[b]} does not reproduce it")
    return out


def descriptors_from_dense(attention_mask: torch.Tensor, verify: bool = True) -> List[Desc]:
    """attention_mask: additive [B,1,L,L] (0 = attend) or bool (True = attend).  One device->host copy of O(B*L) ints."""
    m = attention_mask
    if m.dim() == 4:
        m = m[:, 0]
    if m.dim() == 2:
        m = m[None]
    if m.is_cuda:
        return _descriptors_from_dense_cuda(m, verify)
    allowed = m if m.dtype == torch.bool else (m == 0)
    B, L, _ = allowed.shape
    dev = allowed.device
    ar = torch.arange(L, device=dev)
    last = allowed[:, L - 1, :]                                      # [B,L] the last row sees every non-pad column
    pad_end = torch.where(last.any(1), last.int().argmax(1), torch.full((B,), L, device=dev))
    sees_last = allowed[:, :, L - 1].clone()                         # rows that see the last column = bidirectional rows
    sees_last[:, L - 1] = False
    any_full = sees_last.any(1)
    fb = torch.where(any_full, sees_last.int().argmax(1), torch.zeros(B, dtype=torch.long, device=dev))
    fe = torch.where(any_full, torch.full((B,), L, device=dev), torch.zeros(B, dtype=torch.long, device=dev))
    # always-visible window: columns the first non-pad row sees beyond causality
    q0 = pad_end.clamp(max=L - 1)
    row0 = allowed[torch.arange(B, device=dev), q0]                  # [B,L]
    beyond = row0 & (ar[None, :] > q0[:, None])
    q0_full = (q0 >= fb) & (q0 < fe)
    beyond = beyond & ~q0_full[:, None]
    any_win = beyond.any(1)
    wb = torch.where(any_win, beyond.int().argmax(1), torch.zeros(B, dtype=torch.long, device=dev))
    we = torch.where(any_win, L - beyond.flip(1).int().argmax(1), torch.zeros(B, dtype=torch.long, device=dev))
    host = torch.stack([pad_end, fb, fe, wb, we], 1).cpu().tolist()
    descs = [tuple(int(v) for v in r) for r in host]
    if verify:
        for b, d in enumerate(descs):
            pred = predicate(L, d, dev)
            rows = ar >= d[0]
            if not torch.equal(pred[rows], allowed[b][rows]):
                raise NotImplementedError(
                    f"attention_mask row {b} is not an omni mask (causal / image-span / window / left-pad): "
                    f"derived descriptor {d} does not reproduce it")
    return descs
