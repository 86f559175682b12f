"""Generate the committed golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

    python tests/golden/make_golden.py        (build container only; needs /root/reference)

All inputs and weights are regenerated from seeds with numpy's Philox (identical on every box); only the
reference's OUTPUTS are stored.  Fixtures:
  masks.npz        reference dense masks (packed bits) for t2i / mmu / mmu_vit rows
  sampler.npz      reference multinomial + mask_by_random_topk outputs on seeded logits/noise
  tiny_t2i.npz     2-layer Showo (full vocabulary): reference logits slice, t2i_generate ids (CFG), mmu_generate tokens
  full_slice.npz   full-size Phi-1.5 Showo: reference logits slice / argmax / margins for one half-filled t2i row
  magvit.npz       full-size MAGVIT-v2: reference decode_code pixels (fp16), encoder z + codes
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_loader as R  # noqa: E402
from oracle import magvit_oracle as MO  # noqa: E402
from oracle import showo_oracle as O  # noqa: E402
from fixtures import (TINY, full_row_inputs, magvit_inputs, mask_rows, sampler_case, sampler_cases, tiny_mmu_inputs,  # noqa: E402
                      tiny_t2i_inputs)

torch.set_num_threads(8)


def main():
    mods = R.load_modules()
    voc = O.ShowoVocab()

    # ---------------------------------------------------------------- masks
    rows = mask_rows(voc)
    m_t2i = mods.prompting.create_attention_mask_predict_next(rows["t2i"], pad_id=O.PAD, soi_id=O.SOI, eoi_id=O.EOI,
                                                              rm_pad_in_image=True)
    m_mmu = mods.prompting.create_attention_mask_for_mmu(rows["mmu"], eoi_id=O.EOI)
    m_vit = mods.prompting.create_attention_mask_for_mmu_vit(torch.zeros(2, 700, 8), system_prompt_len=28)
    np.savez_compressed(os.path.join(HERE, "masks.npz"),
                        t2i=np.packbits((m_t2i[:, 0] == 0).numpy()), t2i_shape=np.array(m_t2i[:, 0].shape),
                        mmu=np.packbits((m_mmu[:, 0] == 0).numpy()), mmu_shape=np.array(m_mmu[:, 0].shape),
                        vit=np.packbits((m_vit[:, 0] == 0).numpy()), vit_shape=np.array(m_vit[:, 0].shape),
                        neg_value=np.array([float(m_t2i.min())]))
    print("masks ok", m_t2i.dtype, float(m_t2i.min()))

    # ---------------------------------------------------------------- sampler (reference functions only)
    out = {}
    for ci, case in enumerate(sampler_cases()):
        c = sampler_case(case, voc)
        probs = c["logits"].softmax(-1)
        # torch.multinomial == argmax(p / q), q = the exponential draw (verified here against the real op)
        g1 = torch.Generator().manual_seed(77)
        s_ref = torch.multinomial(probs.reshape(-1, probs.shape[-1]), 1, generator=g1)[:, 0]
        g2 = torch.Generator().manual_seed(77)
        q = torch.empty_like(probs.reshape(-1, probs.shape[-1])).exponential_(1, generator=g2)
        assert torch.equal(s_ref, torch.argmax(probs.reshape(-1, probs.shape[-1]) / q, -1))
        sampled = torch.argmax(probs.reshape(-1, probs.shape[-1]) / c["expo"], -1).view(c["B"], c["N"])
        unknown = c["ids_minus"] == voc.mask_token_id
        sampled = torch.where(unknown, sampled, c["ids_minus"])
        ratio = 1.0 * (case["step"] + 1) / case["T"]
        mask_ratio = mods.sampling.cosine_schedule(torch.tensor(ratio))
        sel = torch.gather(probs, -1, sampled.long()[..., None]).squeeze(-1)
        sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
        mask_len = (c["N"] * mask_ratio).floor().unsqueeze(0)
        mask_len = torch.max(torch.tensor([1]), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len))
        temperature = c["temp_in"] * (1.0 - ratio)

        class _G:      # feed OUR uniform draw through the reference's mask_by_random_topk
            pass
        orig = torch.Tensor.uniform_
        try:
            torch.Tensor.uniform_ = lambda self, a=0, b=1, generator=None: self.copy_(c["unif"])
            masking = mods.sampling.mask_by_random_topk(mask_len, sel, temperature)
        finally:
            torch.Tensor.uniform_ = orig
        out[f"sampled_{ci}"] = sampled.numpy().astype(np.int32)
        out[f"masking_{ci}"] = masking.numpy()
        out[f"mask_len_{ci}"] = mask_len.numpy().astype(np.int32)
        out[f"temp_{ci}"] = np.array([temperature], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler ok")

    # ---------------------------------------------------------------- tiny model
    dims = O.PhiDims(**TINY)
    W = O.make_showo_weights(dims, seed=3)
    model, _ = R.build_showo(dims, W)
    cond, uncond, mask = tiny_t2i_inputs(voc)
    with torch.no_grad():
        logits = model(torch.cat([cond, uncond]), attention_mask=mask)
    sl = logits[:, 130:386, voc.image_offset:-1]
    g = torch.Generator().manual_seed(21)
    c1 = cond.clone()
    noise_probe = torch.empty(8).exponential_(1, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        ids = model.t2i_generate(input_ids=c1, uncond_input_ids=uncond.clone(), attention_mask=mask, guidance_scale=5.0,
                                 timesteps=6, generator=g, config=R.t2i_config(voc))
    mm = tiny_mmu_inputs(voc)
    toks = []
    for b in range(mm.shape[0]):
        mk = mods.prompting.create_attention_mask_for_mmu(mm[b:b + 1], eoi_id=O.EOI)
        with torch.no_grad():
            r = model.mmu_generate(mm[b:b + 1], attention_mask=mk, max_new_tokens=8, top_k=1)
        toks.append(torch.stack(r).numpy())
    np.savez_compressed(os.path.join(HERE, "tiny_t2i.npz"), logits_pos=np.arange(0, 256, 16),
                        logits_slice=sl[:, ::16].numpy().astype(np.float32), argmax=sl.argmax(-1).numpy().astype(np.int32),
                        t2i_ids=ids.numpy().astype(np.int32), t2i_final_input_ids=c1.numpy(),
                        noise_probe=noise_probe.numpy(), mmu_tokens=np.stack(toks),
                        weight_probe=W["showo.model.layers.1.mlp.fc1.weight"][:4, :4].numpy())
    print("tiny ok", ids.shape)
    del model

    # ---------------------------------------------------------------- full-size slice
    t0 = time.time()
    dims = O.PhiDims()
    W = O.make_showo_weights(dims, seed=0)
    print("full weights", time.time() - t0)
    model, _ = R.build_showo(dims, W)
    ids_full, mask_full = full_row_inputs(voc)
    with torch.no_grad():
        lg = model(ids_full, attention_mask=mask_full)[:, 130:386, voc.image_offset:-1]
    top2 = lg.topk(2, -1).values
    np.savez_compressed(os.path.join(HERE, "full_slice.npz"), logits_pos=np.arange(0, 256, 16),
                        logits_slice=lg[:, ::16].numpy().astype(np.float32), argmax=lg.argmax(-1).numpy().astype(np.int32),
                        margin=(top2[..., 0] - top2[..., 1]).numpy(), logit_std=np.array([float(lg.std())]),
                        weight_probe=W["showo.model.layers.23.mlp.fc2.weight"][:4, :4].numpy())
    print("full ok", time.time() - t0, float(lg.std()))
    del model, W

    # ---------------------------------------------------------------- magvit
    Wm = MO.make_magvit_weights(1)
    vq, _ = R.build_magvit(Wm)
    codes_in, pixels_in = magvit_inputs()
    with torch.no_grad():
        dec = vq.decode_code(codes_in)
        z = vq.encoder(pixels_in)
        codes = vq.get_code(pixels_in)
    np.savez_compressed(os.path.join(HERE, "magvit.npz"), decode=dec.numpy().astype(np.float16), z=z.numpy(),
                        codes=codes.numpy().astype(np.int32),
                        weight_probe=Wm["decoder.conv_in.weight"][:2, :2, 0, 0].numpy())
    print("magvit ok")


if __name__ == "__main__":
    main()
