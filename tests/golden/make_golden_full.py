"""Full-size (24-layer, 1.45 B parameter) golden vectors for BASELINE.json configs[1], [2], [3] from the UNMODIFIED
reference (build container only; needs /root/reference; ~15 min on 8 cores).

    python tests/golden/make_golden_full.py [cfg1] [cfg2] [cfg3]

full_cfg1.npz  configs[1]  t2i 256x256, B = 8, CFG w = 5 (16 rows x 387), the first 3 of 18 denoise steps of the reference's
               own loop (reference forward + the reference-pinned sampler step, noise from numpy Philox so that the GPU box
               regenerates it bit for bit): per step the image ids going in, a slice of the cond / uncond logits, the sampled
               ids, the re-masking, the margin of every categorical draw (top-1 minus top-2 of logit - log(noise)) and the
               distance of every confidence to the re-masking cut-off.
full_cfg2.npz  configs[2]  MMU, 16 rows of L0 = 276, greedy: the first 8 tokens of 16 sequential B = 1 reference
               `mmu_generate` calls + the top-1 / top-2 logit margin behind every token (forward hook on the reference model).
full_cfg3.npz  configs[3]  geometry (N = 1024, L = 1155), one CFG pair, one half-filled denoise-step forward: logits slice,
               argmax and margins.
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
from oracle import showo_oracle as O  # noqa: E402
from fixtures import full_cfg1_inputs, full_cfg1_noise, full_cfg2_inputs, full_cfg3_inputs  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)


def main(which):
    t0 = time.time()
    voc = O.ShowoVocab()
    dims = O.PhiDims()
    W = O.make_showo_weights(dims, seed=0)
    model, mods = R.build_showo(dims, W)
    probe = W["showo.model.layers.23.mlp.fc2.weight"][:4, :4].numpy().copy()
    del W
    print("reference model ready", time.time() - t0, flush=True)

    if "cfg1" in which:
        B, T, w, N = 8, 18, 5.0, 256
        cond, uncond, mask = full_cfg1_inputs(voc)
        P = voc.max_text_len + 1
        ids = cond.clone()
        ids_minus = torch.full((B, N), voc.mask_token_id, dtype=torch.int64)
        temperature = 1.0
        out = dict(weight_probe=probe, logit_pos=np.arange(0, N, 64))
        for s in range(3):
            unc = torch.cat([uncond[:, :P], ids[:, P:]], 1)
            with torch.no_grad():
                lg = model(torch.cat([ids, unc]), attention_mask=mask)[:, -(N + 1):-1, voc.image_offset:-1]
            c, u = lg.chunk(2)
            logits = (1 + w) * c - w * u
            expo, unif = full_cfg1_noise(s, B, N)
            sampled, masking, mask_len, temp_new = O.t2i_sample_step(logits, ids_minus, s, T, temperature, voc.mask_token_id, N, expo, unif)
            race = (logits.reshape(-1, 8192) - torch.log(expo)).topk(2, -1).values
            # confidence exactly as mask_by_random_topk forms it (sampling.py:31-36)
            probs = logits.softmax(-1)
            sel = torch.gather(probs, -1, sampled[..., None]).squeeze(-1)
            sel = torch.where(ids_minus == voc.mask_token_id, sel, torch.finfo(sel.dtype).max)
            conf = O.log_clamped(sel) + temp_new * (-O.log_clamped(-O.log_clamped(unif)))
            cut = torch.gather(conf.sort(-1).values, 1, mask_len.long())
            out[f"s{s}_ids_in"] = ids[:, P + 1:P + 1 + N].numpy().astype(np.int32)
            out[f"s{s}_cond"] = c[:, ::64].numpy().astype(np.float32)
            out[f"s{s}_uncond"] = u[:, ::64].numpy().astype(np.float32)
            out[f"s{s}_sampled"] = sampled.numpy().astype(np.int32)
            out[f"s{s}_masking"] = masking.numpy()
            out[f"s{s}_race_margin"] = (race[:, 0] - race[:, 1]).view(B, N).numpy()
            out[f"s{s}_cut_dist"] = (conf - cut).abs().numpy()
            out[f"s{s}_logit_std"] = np.array([float(c.std()), float(logits.std())])
            temperature = temp_new
            ids[:, P + 1:P + 1 + N] = torch.where(masking, voc.mask_token_id, sampled + voc.image_offset)
            ids_minus = torch.where(masking, voc.mask_token_id, sampled)
            print("cfg1 step", s, time.time() - t0, "masked left", int(masking.sum()), flush=True)
        np.savez_compressed(os.path.join(HERE, "full_cfg1.npz"), **out)

    if "cfg2" in which:
        rows = full_cfg2_inputs(voc)
        n_new = 8
        toks = np.zeros((rows.shape[0], n_new), dtype=np.int32)
        margins = np.zeros((rows.shape[0], n_new), dtype=np.float32)
        grabbed = []
        hook = model.register_forward_hook(lambda m, a, o: grabbed.append((o[0] if isinstance(o, tuple) else o)[:, -1].detach().clone()))
        for b in range(rows.shape[0]):
            grabbed.clear()
            mk = mods.prompting.create_attention_mask_for_mmu(rows[b:b + 1], eoi_id=O.EOI)
            with torch.no_grad():
                r = model.mmu_generate(rows[b:b + 1], attention_mask=mk, max_new_tokens=n_new, top_k=1)
            toks[b] = torch.stack(r).numpy()
            for t in range(n_new):
                top2 = grabbed[t][0].topk(2).values
                margins[b, t] = float(top2[0] - top2[1])
            print("cfg2 row", b, toks[b].tolist(), time.time() - t0, flush=True)
        hook.remove()
        np.savez_compressed(os.path.join(HERE, "full_cfg2.npz"), tokens=toks, margins=margins, weight_probe=probe)

    if "cfg3" in which:
        voc3 = O.ShowoVocab(num_vq_tokens=1024)
        cond, uncond, mask = full_cfg3_inputs(voc3)
        with torch.no_grad():
            lg = model(torch.cat([cond, uncond]), attention_mask=mask)[:, -(1024 + 1):-1, voc3.image_offset:-1]
        top2 = lg.topk(2, -1).values
        np.savez_compressed(os.path.join(HERE, "full_cfg3.npz"), logit_pos=np.arange(0, 1024, 64),
                            logits_slice=lg[:, ::64].numpy().astype(np.float32), argmax=lg.argmax(-1).numpy().astype(np.int32),
                            margin=(top2[..., 0] - top2[..., 1]).numpy(), logit_std=np.array([float(lg.std())]), weight_probe=probe)
        print("cfg3 done", time.time() - t0, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["cfg1", "cfg2", "cfg3"])
