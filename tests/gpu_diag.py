"""One-shot GPU diagnostic: exercises every kernel against torch / the oracle and PRINTS error statistics
(no asserts) so one gpurun round trip gives a full picture.  Usage: python tests/gpu_diag.py [section ...]"""
import ctypes as C
import math
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import showo_b200  # noqa: E402
from showo_b200 import _lib, masks as M  # noqa: E402
from oracle import showo_oracle as O  # noqa: E402
from oracle import magvit_oracle as MO  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.require_gpu()
S = lambda: _lib.current_stream_ptr()  # noqa: E731


def stats(name, got, ref):
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    print(f"  {name}: max|d|={d.max().item():.4e} mean|d|={d.mean().item():.4e} ref_std={ref.std().item():.4e} "
          f"ref_absmax={ref.abs().max().item():.4e} nan={int(torch.isnan(got).sum())}", flush=True)
    return d.max().item()


def sec_gemm():
    print("== gemm")
    torch.manual_seed(0)
    for (Mm, N, K, bn) in [(128, 256, 64, 256), (128, 64, 128, 64), (300, 520, 192, 128), (1000, 2048, 2048, 256),
                           (4128, 6144, 2048, 256), (4128, 2048, 10240, 256), (16, 2048, 2048, 64), (4096, 8192, 2048, 256),
                           (257, 1000, 200, 64)]:
        A = (torch.randn(Mm, K, device=dev) * 0.5).bfloat16()
        Kp = (K + 7) // 8 * 8
        if Kp != K:
            continue
        Bw = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        bias = torch.randn(N, device=dev)
        ref = A.float() @ Bw.float().t() + bias
        # epi 2: f32 out
        out = torch.zeros(Mm, N, device=dev)
        rc = lib.showo_gemm_bf16(_lib.ptr(A), K, _lib.ptr(Bw), K, Mm, N, K, _lib.ptr(out), N, _lib.ptr(bias), None, 0, N, 2, bn, S())
        torch.cuda.synchronize()
        if rc:
            print("  rc", rc, lib.showo_last_error())
            continue
        stats(f"f32 M{Mm} N{N} K{K} bn{bn}", out, ref)
        # epi 0: bf16 + gelu on second half
        out16 = torch.zeros(Mm, N, device=dev, dtype=torch.bfloat16)
        gf = (N // 2) // 32 * 32
        lib.showo_gemm_bf16(_lib.ptr(A), K, _lib.ptr(Bw), K, Mm, N, K, _lib.ptr(out16), N, _lib.ptr(bias), None, 0, gf, 0, bn, S())
        ref16 = ref.clone()
        ref16[:, gf:] = O.gelu_new(ref[:, gf:])
        stats("   bf16+gelu", out16, ref16)
        # epi 1: resid
        res = torch.randn(Mm, N, device=dev)
        outr = res.clone()
        lib.showo_gemm_bf16(_lib.ptr(A), K, _lib.ptr(Bw), K, Mm, N, K, _lib.ptr(outr), N, _lib.ptr(bias), _lib.ptr(outr), N, N, 1, bn, S())
        stats("   resid", outr, ref + res)
    # timing of the two layer GEMMs
    for (Mm, N, K) in [(4128, 14336, 2048), (4128, 2048, 10240), (4096, 8192, 2048)]:
        A = (torch.randn(Mm, K, device=dev) * 0.5).bfloat16()
        Bw = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        out16 = torch.zeros(Mm, N, device=dev, dtype=torch.bfloat16)
        for bn in (128, 256):
            for _ in range(3):
                lib.showo_gemm_bf16(_lib.ptr(A), K, _lib.ptr(Bw), K, Mm, N, K, _lib.ptr(out16), N, None, None, 0, N, 0, bn, S())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                lib.showo_gemm_bf16(_lib.ptr(A), K, _lib.ptr(Bw), K, Mm, N, K, _lib.ptr(out16), N, None, None, 0, N, 0, bn, S())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"  time M{Mm} N{N} K{K} bn{bn}: {ms:.3f} ms  {2 * Mm * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            torch.matmul(A, Bw.t())
        t0.record()
        for _ in range(10):
            torch.matmul(A, Bw.t())
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"  cublas M{Mm} N{N} K{K}: {ms:.3f} ms  {2 * Mm * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


def sec_layernorm():
    print("== layernorm")
    for D in (256, 2048):
        x = torch.randn(777, D, device=dev) * 2 + 0.3
        g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
        out = torch.zeros(777, D, device=dev, dtype=torch.bfloat16)
        lib.showo_layernorm_test(_lib.ptr(x), _lib.ptr(g), _lib.ptr(b), 1e-5, _lib.ptr(out), 777, D, S())
        stats(f"D{D}", out, torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5))


def attn_ref(qkv, n_seq, rows, pos0, H, qg, qb, kg, kb, descs, kprev=None, vprev=None):
    """fp32 torch reference for showo_attention_test on the bf16-rounded inputs. Keys = rows at pos0.. (plus kprev)."""
    D = H * 64
    x = qkv.float().view(n_seq, rows, -1)
    k = x[..., :D].reshape(n_seq, rows, H, 64).transpose(1, 2)
    v = x[..., D:2 * D].reshape(n_seq, rows, H, 64).transpose(1, 2)
    q = x[..., 2 * D:3 * D].reshape(n_seq, rows, H, 64).transpose(1, 2)
    q = torch.nn.functional.layer_norm(q, (64,), qg, qb, 1e-5)
    k = torch.nn.functional.layer_norm(k, (64,), kg, kb, 1e-5)
    dims = O.PhiDims()
    cos, sin = O.rotary_tables(dims, pos0 + rows)
    cos, sin = cos.to(dev)[pos0:], sin.to(dev)[pos0:]
    q = O.apply_partial_rotary(q, cos, sin, 32)
    k = O.apply_partial_rotary(k, cos, sin, 32)
    q = q.bfloat16().float(); k = k.bfloat16().float()
    if kprev is not None:
        k = torch.cat([kprev, k], 2); v = torch.cat([vprev, v], 2)
    L = k.shape[2]
    s = (q @ k.transpose(-1, -2)) / 8.0
    for i, d in enumerate(descs):
        ok = M.predicate(L, d, dev)[pos0:pos0 + rows]
        s[i, :, ~ok] = float("-inf")
    p = torch.softmax(s, -1)
    o = (p @ v).transpose(1, 2).reshape(n_seq, rows, D)
    return o, k, v


def sec_attention():
    print("== attention")
    torch.manual_seed(1)
    H = 4; D = H * 64; ld = 3 * D + 128
    for (n_seq, rows, descs) in [
        (2, 387, [(100, 129, 387, 0, 0), (0, 129, 387, 0, 0)]),
        (2, 200, [(0, 0, 0, 0, 0), (0, 0, 0, 3, 150)]),
        (1, 64, [(5, 0, 0, 0, 0)]),
        (3, 130, [(20, 60, 130, 0, 0), (0, 0, 0, 0, 77), (129, 0, 0, 0, 0)]),
    ]:
        Lmax = (rows + 63) // 64 * 64
        qkv = (torch.randn(n_seq * rows, ld, device=dev)).bfloat16()
        qg = 1 + 0.1 * torch.randn(64, device=dev); qb = 0.1 * torch.randn(64, device=dev)
        kg = 1 + 0.1 * torch.randn(64, device=dev); kb = 0.1 * torch.randn(64, device=dev)
        ref, kk, vv = attn_ref(qkv, n_seq, rows, 0, H, qg, qb, kg, kb, descs)
        kc = torch.zeros(n_seq, H, Lmax, 64, device=dev, dtype=torch.bfloat16)
        vc = torch.zeros(n_seq, H, 64, Lmax, device=dev, dtype=torch.bfloat16)
        buf = qkv.clone()
        rc = lib.showo_attention_test(_lib.ptr(buf), ld, n_seq, rows, 0, H, _lib.ptr(qg), _lib.ptr(qb), _lib.ptr(kg), _lib.ptr(kb),
                                      1e-5, 10000.0, 32, _lib.ptr(kc), _lib.ptr(vc), Lmax, rows, _lib.masks_array(descs), S())
        torch.cuda.synchronize()
        if rc:
            print("  rc", rc, lib.showo_last_error()); continue
        got = buf.view(n_seq, rows, ld)[..., 2 * D:3 * D]
        stats(f"K cache seq{n_seq} rows{rows}", kc[:, :, :rows], kk)
        stats("   Vt cache", vc[:, :, :, :rows], vv.transpose(-1, -2))
        for i, d in enumerate(descs):
            stats(f"   out seq{i} desc{d} (non-pad rows)", got[i, d[0]:], ref[i, d[0]:])
    # step-style: prefix in cache, then image rows at pos0
    n_seq, P, R = 2, 129, 258
    L = P + R; Lmax = 448
    descs = [(40, P, L, 0, 0), (126, P, L, 0, 0)]
    qkv_full = torch.randn(n_seq * L, ld, device=dev).bfloat16()
    qg = torch.ones(64, device=dev); qb = torch.zeros(64, device=dev); kg = qg.clone(); kb = qb.clone()
    ref, _, _ = attn_ref(qkv_full, n_seq, L, 0, H, qg, qb, kg, kb, descs)
    kc = torch.zeros(n_seq, H, Lmax, 64, device=dev, dtype=torch.bfloat16)
    vc = torch.zeros(n_seq, H, 64, Lmax, device=dev, dtype=torch.bfloat16)
    full3 = qkv_full.view(n_seq, L, ld)
    pre = full3[:, :P].reshape(-1, ld).clone(); img = full3[:, P:].reshape(-1, ld).clone()
    lib.showo_attention_test(_lib.ptr(pre), ld, n_seq, P, 0, H, _lib.ptr(qg), _lib.ptr(qb), _lib.ptr(kg), _lib.ptr(kb),
                             1e-5, 10000.0, 32, _lib.ptr(kc), _lib.ptr(vc), Lmax, P, _lib.masks_array(descs), S())
    lib.showo_attention_test(_lib.ptr(img), ld, n_seq, R, P, H, _lib.ptr(qg), _lib.ptr(qb), _lib.ptr(kg), _lib.ptr(kb),
                             1e-5, 10000.0, 32, _lib.ptr(kc), _lib.ptr(vc), Lmax, L, _lib.masks_array(descs), S())
    torch.cuda.synchronize()
    for i, d in enumerate(descs):
        stats(f"step-style prefix rows seq{i}", pre.view(n_seq, P, ld)[i, d[0]:, 2 * D:3 * D], ref[i, d[0]:P])
        stats(f"step-style image rows seq{i}", img.view(n_seq, R, ld)[i, :, 2 * D:3 * D], ref[i, P:])
    # decode-style: 1 query at the end
    q1 = full3[:, L - 1:L].reshape(-1, ld).clone()
    kc2, vc2 = kc.clone(), vc.clone()
    lib.showo_attention_test(_lib.ptr(q1), ld, n_seq, 1, L - 1, H, _lib.ptr(qg), _lib.ptr(qb), _lib.ptr(kg), _lib.ptr(kb),
                             1e-5, 10000.0, 32, _lib.ptr(kc2), _lib.ptr(vc2), Lmax, L, _lib.masks_array(descs), S())
    torch.cuda.synchronize()
    stats("decode-style last row", q1.view(n_seq, 1, ld)[:, 0, 2 * D:3 * D], ref[:, L - 1])


def sec_sampler():
    print("== sampler")
    voc = O.ShowoVocab()
    g = torch.Generator().manual_seed(3)
    B, N, Cc = 3, 256, 8192
    for step, T, w in [(0, 18, 5.0), (7, 18, 5.0), (17, 18, 0.0), (3, 8, 2.0)]:
        cond = torch.randn(B, N, Cc, generator=g) * 1.5
        unc = torch.randn(B, N, Cc, generator=g) * 1.5
        ids_minus = torch.full((B, N), voc.mask_token_id, dtype=torch.int64)
        known = torch.rand(B, N, generator=g) < (step / T)
        ids_minus[known] = torch.randint(0, Cc, (int(known.sum()),), generator=g)
        expo = torch.empty(B * N, Cc).exponential_(1, generator=g)
        unif = torch.zeros(B, N).uniform_(0, 1, generator=g)
        logits = (1 + w) * cond - w * unc if w > 0 else cond
        floors, temps = showo_b200.step_schedule(showo_b200.cosine_schedule, T, N, 1.0)
        temp_prev = 1.0 if step == 0 else temps[step - 1]
        samp, masking, mask_len, tnew = O.t2i_sample_step(logits, ids_minus, step, T, temp_prev, voc.mask_token_id, N, expo, unif)
        L = 129 + N + 2
        ids = torch.full((B, L), 7, dtype=torch.int64)
        ids[:, 130:130 + N] = torch.where(ids_minus == voc.mask_token_id, ids_minus, ids_minus + voc.image_offset)
        ids_d = ids.to(dev)
        out = torch.zeros(B, N, dtype=torch.int64, device=dev)
        mk = torch.zeros(B, N, dtype=torch.uint8, device=dev)
        cond_d, unc_d, expo_d, unif_d = cond.to(dev), unc.to(dev), expo.to(dev), unif.to(dev)
        rc = lib.showo_sampler_step(_lib.ptr(cond_d), _lib.ptr(unc_d) if w > 0 else None, B, N, Cc, w, _lib.ptr(ids_d), L, 130,
                                    voc.image_offset, voc.mask_token_id, floors[step], temps[step], _lib.ptr(expo_d),
                                    _lib.ptr(unif_d), 0, step, _lib.ptr(out), _lib.ptr(mk), S())
        torch.cuda.synchronize()
        if rc:
            print("  rc", rc, lib.showo_last_error()); continue
        new_ids = torch.where(masking, voc.mask_token_id, samp + voc.image_offset)
        print(f"  step{step}/{T} w{w}: sampled mismatches {(out.cpu() != samp).sum().item()}/{B * N}  masking mismatches "
              f"{(mk.cpu().bool() != masking).sum().item()}  ids mismatches {(ids_d.cpu()[:, 130:130 + N] != new_ids).sum().item()} "
              f"mask_len {mask_len.flatten().tolist()} temp {tnew:.4f} vs {temps[step]:.4f} masked {int(masking.sum())}", flush=True)
    # philox mode smoke: valid codes, right number masked
    ids = torch.full((B, L), voc.mask_token_id, dtype=torch.int64, device=dev)
    out = torch.zeros(B, N, dtype=torch.int64, device=dev)
    mk = torch.zeros(B, N, dtype=torch.uint8, device=dev)
    cond_d = cond.to(dev)
    lib.showo_sampler_step(_lib.ptr(cond_d), None, B, N, Cc, 0.0, _lib.ptr(ids), L, 130, voc.image_offset, voc.mask_token_id,
                           200, 0.9, None, None, 1234, 0, _lib.ptr(out), _lib.ptr(mk), S())
    torch.cuda.synchronize()
    print("  philox: codes in range", int(out.min()), int(out.max()), "masked per row", mk.sum(1).tolist(),
          "hist-entropy-ish unique", out.unique().numel())


def tiny_model(n_layers=2):
    dims = O.PhiDims(hidden=256, n_layers=n_layers, n_heads=4, ffn=1024)
    W = O.make_showo_weights(dims, seed=3)
    voc = O.ShowoVocab()
    m = showo_b200.Showo(False, dims.vocab_size, voc.llm_vocab_size, phi_dims=dict(hidden=256, n_layers=n_layers, n_heads=4, ffn=1024),
                         materialize=False)
    m.load_weights(W, device=dev)
    return dims, voc, W, m


def t2i_cfg(voc):
    from types import SimpleNamespace as NS
    return NS(model=NS(showo=NS(num_vq_tokens=voc.num_vq_tokens, num_new_special_tokens=voc.num_new_special_tokens,
                                llm_vocab_size=voc.llm_vocab_size)),
              dataset=NS(preprocessing=NS(max_seq_length=voc.max_text_len)))


def sec_forward():
    print("== forward (tiny) vs oracle")
    dims, voc, W, m = tiny_model()
    cond, uncond = O.make_t2i_prompts(2, voc, seed=5)
    g = torch.Generator().manual_seed(0)
    fill = torch.rand(2, 256, generator=g) < 0.5
    codes = torch.randint(0, 8192, (2, 256), generator=g) + voc.image_offset
    cond[:, 130:386] = torch.where(fill, codes, cond[:, 130:386])
    ids = torch.cat([cond, torch.cat([uncond[:, :129], cond[:, 129:]], 1)])
    mask = O.create_attention_mask_predict_next(ids)
    with torch.no_grad():
        ref, hid = O.showo_logits(W, dims, input_ids=ids, add_mask=mask, return_hidden=True)
    got = m(ids.to(dev), attention_mask=mask.to(dev))
    torch.cuda.synchronize()
    descs = M.descriptors_from_dense(mask.to(dev))
    print("  descs", descs)
    for b in range(ids.shape[0]):
        pe = descs[b][0]
        stats(f"logits row{b} (non-pad)", got[b, pe:].cpu(), ref[b, pe:])
    am_ref = ref[:, 130:386, voc.image_offset:-1].argmax(-1)
    am_got = got[:, 130:386, voc.image_offset:-1].argmax(-1).cpu()
    print("  argmax agree", (am_ref == am_got).float().mean().item())
    # step logits with prefix reuse vs full forward
    cfg = t2i_cfg(voc)
    sl = m.t2i_step_logits(cond.to(dev), uncond.to(dev), mask.to(dev), guidance_scale=5, config=cfg)
    torch.cuda.synchronize()
    full = got[:, 130:386, voc.image_offset:-1]
    stats("prefix-reuse step logits vs full forward (engine)", sl.cpu(), full.cpu())
    stats("prefix-reuse step logits vs oracle", sl.cpu(), ref[:, 130:386, voc.image_offset:-1])
    # mmu / lm masks
    g2 = torch.Generator().manual_seed(9)
    codes = torch.randint(0, 8192, (2, 256), generator=g2)
    mm = O.make_mmu_prompts(2, voc, codes, q_len=12)
    mk = O.create_attention_mask_for_mmu(mm)
    with torch.no_grad():
        ref2 = O.showo_logits(W, dims, input_ids=mm, add_mask=mk)
    got2 = m(mm.to(dev), attention_mask=mk.to(dev))
    stats("mmu-mask logits", got2.cpu(), ref2)
    return dims, voc, W, m


def sec_t2i():
    print("== t2i_generate (tiny), teacher-forced parity per step")
    dims, voc, W, m = tiny_model()
    cfg = t2i_cfg(voc)
    B, T, w = 2, 6, 5.0
    cond, uncond = O.make_t2i_prompts(B, voc, seed=11)
    mask = O.create_attention_mask_predict_next(torch.cat([cond, uncond]))
    trace = []
    g = torch.Generator().manual_seed(21)
    c1 = cond.clone()
    with torch.no_grad():
        ref = O.t2i_generate(W, dims, voc, c1, uncond.clone(), mask, guidance_scale=w, timesteps=T, generator=g, trace=trace)
    floors, temps = showo_b200.step_schedule(showo_b200.cosine_schedule, T, 256, 1.0)
    md = mask.to(dev)
    for s, tr in enumerate(trace):
        ids_in = tr.input_ids_in.to(dev)
        sl = m.t2i_step_logits(ids_in, uncond.to(dev), md, guidance_scale=w, config=cfg)
        lg = (1 + w) * sl[:B] - w * sl[B:]
        dmax = stats(f"step{s} post-CFG logits", lg.cpu(), tr.logits)
        ids_d = ids_in.clone()
        out = torch.zeros(B, 256, dtype=torch.int64, device=dev)
        mk = torch.zeros(B, 256, dtype=torch.uint8, device=dev)
        # sampler on ORACLE logits (exactness of the sampler) and on engine logits (end-to-end)
        for name, lc, lu, ww in (("oracle-logits", tr.logits.to(dev).contiguous(), None, 0.0), ("engine-logits", sl[:B].contiguous(), sl[B:].contiguous(), w)):
            ids_d = ids_in.clone()
            expo_d, unif_d = tr.expo.to(dev), tr.uniform.to(dev)
            lib.showo_sampler_step(_lib.ptr(lc), _lib.ptr(lu), B, 256, 8192, ww, _lib.ptr(ids_d), ids_d.shape[1], 130, voc.image_offset,
                                   voc.mask_token_id, floors[s], temps[s], _lib.ptr(expo_d), _lib.ptr(unif_d), 0, s,
                                   _lib.ptr(out), _lib.ptr(mk), S())
            torch.cuda.synchronize()
            print(f"    sampler[{name}] sampled mism {(out.cpu() != tr.sampled_ids).sum().item()} masking mism "
                  f"{(mk.cpu().bool() != tr.masking).sum().item()}", flush=True)
    # full loop, host noise from the same generator seed: compare with oracle end-to-end (bf16 => may diverge)
    g2 = torch.Generator(device=dev).manual_seed(21)
    c2 = cond.clone().to(dev)
    out = m.t2i_generate(c2, uncond.to(dev), md, guidance_scale=w, timesteps=T, generator=g2, config=cfg)
    torch.cuda.synchronize()
    print("  full loop out range", int(out.min()), int(out.max()), "launches", m.kernel_launches())
    out2 = m.t2i_generate(cond.clone().to(dev), uncond.to(dev), md, guidance_scale=w, timesteps=T, generator=None, config=cfg)
    print("  philox loop out range", int(out2.min()), int(out2.max()), "unique", out2.unique().numel())


def sec_mmu():
    print("== mmu_generate (tiny)")
    dims, voc, W, m = tiny_model()
    g = torch.Generator().manual_seed(4)
    B = 3
    codes = torch.randint(0, 8192, (B, 256), generator=g)
    ids = O.make_mmu_prompts(B, voc, codes, q_len=10)
    n_new = 12
    toks, lens = m.mmu_generate_batched(ids.to(dev), attention_mask=O.create_attention_mask_for_mmu(ids).to(dev),
                                        max_new_tokens=n_new, top_k=1)
    torch.cuda.synchronize()
    for b in range(B):
        mk = O.create_attention_mask_for_mmu(ids[b:b + 1])
        with torch.no_grad():
            ref = O.mmu_generate(W, dims, ids[b:b + 1], mk, max_new_tokens=n_new, top_k=1)
        ref = torch.stack(ref)
        print(f"  row{b} engine {toks[b].tolist()} oracle {ref.tolist()} agree {(toks[b].cpu() == ref).float().mean().item():.2f}", flush=True)


def sec_magvit():
    print("== magvit")
    W = MO.make_magvit_weights(1)
    vq = showo_b200.MAGVITv2(materialize=False)
    vq.load_weights(W, device=dev)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 8192, (2, 256), generator=g)
    t = time.time()
    with torch.no_grad():
        ref = MO.decode_code(ids, W)
    print("  oracle decode s", time.time() - t)
    got = vq.decode_code(ids.to(dev))
    torch.cuda.synchronize()
    stats("decode_code", got.cpu(), ref)
    u8 = vq.decode_code_uint8(ids.to(dev)).cpu()
    ref_u8 = (torch.clamp((ref + 1) / 2, 0, 1) * 255).permute(0, 2, 3, 1).numpy().astype("uint8")
    d8 = (u8.int() - torch.from_numpy(ref_u8).int()).abs()
    print("  uint8 image max diff", int(d8.max()), "mean", d8.float().mean().item())
    x = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    with torch.no_grad():
        z = MO.encoder_forward(x, W)
        ref_codes = MO.lfq_indices(z)
    codes = vq.get_code(x.to(dev)).cpu()
    bits_ref = (z > 0).reshape(2, 13, -1)
    bits_got = ((codes[:, None, :] >> torch.arange(12, -1, -1)[None, :, None]) & 1).bool()
    mism = bits_ref != bits_got
    zz = z.reshape(2, 13, -1).abs()
    print(f"  get_code: code agree {(codes == ref_codes).float().mean().item():.3f} bit mismatches {int(mism.sum())}/{mism.numel()} "
          f"max|z| at mismatch {zz[mism].max().item() if mism.any() else 0:.4f} z std {z.std().item():.3f}")
    for _ in range(2):
        vq.decode_code_uint8(torch.randint(0, 8192, (8, 256)).to(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ids8 = torch.randint(0, 8192, (8, 256)).to(dev)
    e0.record(); vq.decode_code_uint8(ids8); e1.record(); torch.cuda.synchronize()
    print("  decode B=8 ms", e0.elapsed_time(e1), "launches", vq.kernel_launches())


def sec_conv():
    print("== conv (implicit GEMM)")
    torch.manual_seed(5)
    for (NB, H, W_, cin, cout, taps) in [(2, 16, 16, 64, 512, 9), (1, 32, 32, 128, 128, 9), (2, 64, 64, 256, 128, 1), (1, 16, 16, 512, 13, 9)]:
        k = 3 if taps == 9 else 1
        x = torch.randn(NB, cin, H, W_, device=dev)
        w = torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)
        b = torch.randn(cout, device=dev)
        res = torch.randn(NB, cout, H, W_, device=dev)
        xb = x.permute(0, 2, 3, 1).contiguous().bfloat16()
        cout_pad = (cout + 63) // 64 * 64
        wp = torch.zeros(cout_pad, taps * cin, device=dev)
        wp[:cout] = w.permute(0, 2, 3, 1).reshape(cout, taps * cin)
        wp = wp.bfloat16()
        rb = res.permute(0, 2, 3, 1).contiguous().bfloat16()
        out = torch.zeros(NB * H * W_, cout, device=dev, dtype=torch.bfloat16)
        rc = lib.showo_conv_test(_lib.ptr(xb), _lib.ptr(wp), _lib.ptr(b), _lib.ptr(rb), _lib.ptr(out), NB, H, W_, cin, cout, taps, S())
        torch.cuda.synchronize()
        if rc:
            print("  rc", rc, lib.showo_last_error()); continue
        ref = torch.nn.functional.conv2d(xb.float().permute(0, 3, 1, 2), wp[:cout].float().view(cout, k, k, cin).permute(0, 3, 1, 2), b, padding=k // 2)
        ref = ref + rb.float().permute(0, 3, 1, 2)
        stats(f"conv NB{NB} {H}x{W_} {cin}->{cout} k{k}", out.view(NB, H, W_, cout).permute(0, 3, 1, 2), ref)


SECTIONS = {"gemm": sec_gemm, "layernorm": sec_layernorm, "attention": sec_attention, "sampler": sec_sampler,
            "forward": sec_forward, "t2i": sec_t2i, "mmu": sec_mmu, "conv": sec_conv, "magvit": sec_magvit}

if __name__ == "__main__":
    names = sys.argv[1:] or list(SECTIONS)
    print(torch.cuda.get_device_name(0), "sms", torch.cuda.get_device_properties(0).multi_processor_count)
    for n in names:
        try:
            SECTIONS[n]()
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("CUDA context broken:", e)
                break
