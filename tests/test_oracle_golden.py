"""CPU tests: the oracle (oracle/*.py) against the committed golden vectors that tests/golden/make_golden.py produced by
running the unmodified reference, plus closed-form properties of the path."""
import math
import os

import numpy as np
import pytest
import torch

import fixtures as FX
from oracle import magvit_oracle as MO
from oracle import showo_oracle as O

VOC = O.ShowoVocab()


def test_vocabulary_arithmetic():
    # configs/showo_demo.yaml:19-24: 50295 + 10 + 8192 + 1 = 58498, mask id = V - 1
    assert VOC.vocab_size == 58498 and VOC.mask_token_id == 58497 and VOC.image_offset == 50305
    assert (O.PAD, O.SOI, O.EOI, O.T2I, O.MMU) == (50295, 50296, 50297, 50300, 50301)


def test_masks_match_reference_golden():
    z = FX.load("masks.npz")
    rows = FX.mask_rows(VOC)
    ref_t2i = FX.unpack_mask(z, "t2i")
    assert torch.equal(O.mask_allowed_t2i(rows["t2i"]), ref_t2i)
    add = O.create_attention_mask_predict_next(rows["t2i"])
    assert add.dtype == torch.float32 and add.shape == (5, 1, 387, 387) and float(add.min()) == float(z["neg_value"][0])
    assert torch.equal(O.mask_allowed_mmu(rows["mmu"]), FX.unpack_mask(z, "mmu"))
    assert torch.equal(O.mask_allowed_mmu_vit(2, 700, system_prompt_len=28), FX.unpack_mask(z, "vit"))


def test_closed_form_predicate_equals_dense_masks_on_non_pad_rows():
    import showo_b200
    M = showo_b200.masks
    z = FX.load("masks.npz")
    rows = FX.mask_rows(VOC)
    ref = FX.unpack_mask(z, "t2i")
    descs = M.descriptors_t2i(rows["t2i"], O.PAD, O.SOI, O.EOI)
    for b, d in enumerate(descs):
        pred = M.predicate(387, d)
        assert torch.equal(pred[d[0]:], ref[b, d[0]:]), (b, d)
    assert descs[4][0] == 0 and descs[3][0] == 126                      # no-pad row, '' prompt row
    # recovered from the dense tensor alone (what Showo.forward does with the caller's attention_mask)
    assert M.descriptors_from_dense(O.additive_from_allowed(ref)) == descs
    mm = FX.unpack_mask(z, "mmu")
    d2 = M.descriptors_from_dense(O.additive_from_allowed(mm))
    assert M.descriptors_mmu(rows["mmu"], O.EOI) == [(0, 0, 0, 0, 259)] * 2            # eoi sits at position 258
    for b in range(2):      # the window recovered from the dense mask may start at 1 (column 0 is causal anyway)
        assert torch.equal(M.predicate(mm.shape[1], d2[b]), mm[b])
        assert torch.equal(M.predicate(mm.shape[1], (0, 0, 0, 0, 259)), mm[b])
    dv = M.descriptors_from_dense(O.additive_from_allowed(FX.unpack_mask(z, "vit")))
    for d in dv:
        assert torch.equal(M.predicate(700, d), FX.unpack_mask(z, "vit")[0])
    assert M.descriptors_from_dense(O.additive_from_allowed(torch.tril(torch.ones(1, 9, 9, dtype=torch.bool)))) == [(0, 0, 0, 0, 0)]
    bad = torch.tril(torch.ones(1, 9, 9, dtype=torch.bool))
    bad[0, 2, 7] = True
    bad[0, 3, 5] = True
    with pytest.raises(NotImplementedError):
        M.descriptors_from_dense(O.additive_from_allowed(bad))


def test_sampler_step_matches_reference_golden():
    z = FX.load("sampler.npz")
    for ci, case in enumerate(FX.sampler_cases()):
        c = FX.sampler_case(case, VOC)
        samp, masking, mask_len, temp = O.t2i_sample_step(c["logits"], c["ids_minus"], case["step"], case["T"], c["temp_in"],
                                                          VOC.mask_token_id, c["N"], c["expo"], c["unif"])
        assert np.array_equal(samp.numpy(), z[f"sampled_{ci}"]), ci
        assert np.array_equal(masking.numpy(), z[f"masking_{ci}"]), ci
        assert np.array_equal(mask_len.numpy().astype(np.int32), z[f"mask_len_{ci}"]), ci
        assert temp == float(z[f"temp_{ci}"][0])


def test_schedule_quirks():
    import showo_b200
    floors, temps = showo_b200.step_schedule(showo_b200.cosine_schedule, 18, 256, 1.0)
    assert floors[-1] == -1                      # cos(pi/2) = -4.37e-8 in fp32 -> floor(256 * .) = -1 -> clamps to 1
    assert floors[0] == 255 and floors[8] == 181 and floors[16] == 22
    assert abs(temps[0] - 17 / 18) < 1e-12 and temps[-1] == 0.0 and abs(temps[3] - 0.5441243713) < 1e-9
    assert float(O.cosine_schedule(torch.tensor(1.0))) < 0
    f8, t8 = showo_b200.step_schedule(showo_b200.get_mask_chedule("linear"), 8, 1024, 2.0)
    assert f8[0] == 896 and f8[-1] == 0 and abs(t8[0] - 1.75) < 1e-12
    with pytest.raises(ValueError):
        showo_b200.get_mask_chedule("nope")
    assert float(showo_b200.get_mask_chedule("pow2")(torch.tensor(0.5))) == 0.75


def _noise_reproducible(z):
    probe = torch.empty(8).exponential_(1, generator=torch.Generator().manual_seed(21))
    return np.array_equal(probe.numpy(), z["noise_probe"])


def test_tiny_model_matches_reference_golden():
    z = FX.load("tiny_t2i.npz")
    dims = O.PhiDims(**FX.TINY)
    W = O.make_showo_weights(dims, seed=3)
    assert np.array_equal(W["showo.model.layers.1.mlp.fc1.weight"][:4, :4].numpy(), z["weight_probe"])
    cond, uncond, mask = FX.tiny_t2i_inputs(VOC)
    with torch.no_grad():
        lg = O.showo_logits(W, dims, input_ids=torch.cat([cond, uncond]), add_mask=mask)[:, 130:386, VOC.image_offset:-1]
    assert np.abs(lg[:, ::16].numpy() - z["logits_slice"]).max() < 2e-5
    assert (lg.argmax(-1).numpy() == z["argmax"]).mean() > 0.999
    if not _noise_reproducible(z):
        pytest.skip("torch CPU exponential_ stream differs on this host; t2i replay needs identical noise")
    g = torch.Generator().manual_seed(21)
    c1 = cond.clone()
    with torch.no_grad():
        ids = O.t2i_generate(W, dims, VOC, c1, uncond.clone(), mask, guidance_scale=5.0, timesteps=6, generator=g)
    assert np.array_equal(ids.numpy(), z["t2i_ids"])
    assert np.array_equal(c1.numpy(), z["t2i_final_input_ids"])          # in-place mutation like the reference
    assert int(ids.min()) >= 0 and int(ids.max()) < 8192


def test_tiny_mmu_greedy_matches_reference_golden():
    z = FX.load("tiny_t2i.npz")
    dims = O.PhiDims(**FX.TINY)
    W = O.make_showo_weights(dims, seed=3)
    mm = FX.tiny_mmu_inputs(VOC)
    mk = O.create_attention_mask_for_mmu(mm[1:2])
    with torch.no_grad():
        r = O.mmu_generate(W, dims, mm[1:2], mk, max_new_tokens=8, top_k=1)
    assert np.array_equal(torch.stack(r).numpy(), z["mmu_tokens"][1])


def test_magvit_matches_reference_golden():
    z = FX.load("magvit.npz")
    W = MO.make_magvit_weights(1)
    assert np.array_equal(W["decoder.conv_in.weight"][:2, :2, 0, 0].numpy(), z["weight_probe"])
    assert sum(v.numel() for v in W.values()) == 55439171 + 39947833          # SURVEY.md section 6 [probed]
    codes_in, pixels_in = FX.magvit_inputs()
    with torch.no_grad():
        dec = MO.decode_code(codes_in, W)
        zz = MO.encoder_forward(pixels_in, W)
    assert np.abs(dec.numpy() - z["decode"].astype(np.float32)).max() < 2e-3        # fp16 storage
    assert np.abs(zz.numpy() - z["z"]).max() < 1e-4
    assert np.array_equal(MO.lfq_indices(zz).numpy(), z["codes"])


def test_lfq_bit_order_roundtrip():
    # channel 0 is the MSB (modeling_magvitv2.py:186-206); index -> +-1 vector -> index is the identity
    idx = torch.arange(8192).view(1, -1)
    e = MO.lfq_entry(idx, 64, 128)
    assert e.shape == (1, 13, 64, 128) and set(e.unique().tolist()) == {-1.0, 1.0}
    assert torch.equal(MO.lfq_indices(e), idx)
    assert torch.equal(e[0, :, 0, 1], torch.tensor([-1.0] * 12 + [1.0]))


def test_text_prefix_is_step_invariant():
    """Rows before soi never see image columns (a-3): swapping the image ids leaves their hidden states unchanged,
    which is what makes the engine's prefix reuse exact."""
    dims = O.PhiDims(**FX.TINY)
    W = O.make_showo_weights(dims, seed=3)
    cond, uncond, mask = FX.tiny_t2i_inputs(VOC)
    a = cond[:1].clone()
    b = cond[:1].clone()
    b[0, 130:386] = VOC.mask_token_id
    m = O.create_attention_mask_predict_next(a)
    with torch.no_grad():
        _, ha = O.showo_logits(W, dims, input_ids=a, add_mask=m, return_hidden=True)
        _, hb = O.showo_logits(W, dims, input_ids=b, add_mask=m, return_hidden=True)
    for x, y in zip(ha, hb):
        assert torch.equal(x[:, :129], y[:, :129])
    assert not torch.equal(ha[-1][:, 129:], hb[-1][:, 129:])


@pytest.mark.skipif(not os.environ.get("SHOWO_SLOW"), reason="full-size CPU forward (several minutes): set SHOWO_SLOW=1")
def test_full_size_slice_matches_reference_golden():
    z = FX.load("full_slice.npz")
    dims = O.PhiDims()
    W = O.make_showo_weights(dims, seed=0)
    ids, mask = FX.full_row_inputs(VOC)
    with torch.no_grad():
        lg = O.showo_logits(W, dims, input_ids=ids, add_mask=mask)[:, 130:386, VOC.image_offset:-1]
    assert np.abs(lg[:, ::16].numpy() - z["logits_slice"]).max() < 1e-4


def test_train_step_losses_and_gradients_match_reference_golden():
    """Showo.forward with labels (modeling_showo.py:81-100) and the gradients of the weighted loss on a mixed t2i / lm / mmu
    batch: the oracle's restated forward, differentiated by autograd, against what the unmodified reference produced
    (tests/golden/make_golden_train.py).  This is the pin the CUDA backward of the next round is tested against."""
    z = FX.load("train_step.npz")
    dims = O.PhiDims(**FX.TINY)
    W = {k: v.clone().requires_grad_(True) for k, v in O.make_showo_weights(dims, seed=3).items()}
    ids, mask, labels, (bt, bl, bm) = FX.train_batch(VOC)
    logits = O.showo_logits(W, dims, input_ids=ids, add_mask=mask)
    l1, l2, l3 = O.showo_losses(logits, labels, bt, bl, bm, 128)
    assert np.allclose([l1.item(), l2.item(), l3.item()], z["losses"], rtol=2e-6, atol=0)
    assert np.abs(logits[:, ::32, ::997].detach().numpy() - z["logits_slice"]).max() < 2e-5
    c = FX.TRAIN_COEFF
    (c[0] * l1 + c[1] * l2 + c[2] * l3).backward()
    names = [str(n) for n in z["grad_names"]]
    assert sorted(W.keys()) == names                                  # every parameter receives a gradient
    norms = np.array([W[k].grad.double().norm().item() for k in names])
    assert np.allclose(norms, z["grad_norms"], rtol=2e-4, atol=1e-9), np.abs(norms / z["grad_norms"] - 1).max()
    for k in FX.TRAIN_GRAD_PROBES:
        g = W[k].grad
        got = (g[:8, :8] if g.dim() == 2 else g[:64]).numpy()
        ref = z["grad:" + k]
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-9, k


def test_train_input_oracle_against_reference_golden():
    """oracle/train_inputs_oracle.py (mask_or_random_replace_tokens + t2i_prompt restated on numpy) == the unmodified reference's
    outputs on the same draws (tests/golden/make_golden_prep.py), bit for bit."""
    import importlib.util
    from oracle import train_inputs_oracle as TO
    spec = importlib.util.spec_from_file_location("make_golden_prep", os.path.join(os.path.dirname(__file__), "golden", "make_golden_prep.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_prep.npz"))
    for name in ("a256", "b1024", "c_trunc"):
        N, T, rate, drop, texts, codes, _ = mg.case_inputs(name)
        mp = TO.cosine_mask_prob(z[f"{name}_timesteps"])
        ids_img, lab_img, mpc = TO.mask_image_tokens(codes.numpy(), 58497, mp, z[f"{name}_rand"], rate)
        assert np.array_equal(ids_img, z[f"{name}_ids_img"]) and np.array_equal(lab_img, z[f"{name}_lab_img"])
        assert np.allclose(mpc, z[f"{name}_mask_prob"], rtol=0, atol=1e-6)
        ids, masks, labels = TO.t2i_prompt_rows(texts, ids_img, lab_img, z[f"{name}_probs"], max_text_len=T, pad=50295, bos=50256, eos=50256,
                                               task=50300, soi=50296, eoi=50297, cond_dropout_prob=drop)
        assert np.array_equal(ids, z[f"{name}_ids"]) and np.array_equal(labels, z[f"{name}_labels"]) and np.array_equal(masks, z[f"{name}_masks"])
        assert np.array_equal(TO.t2i_descriptors(ids, 50295, 50296, 50297), z[f"{name}_descs"])


# ------------------------------------------------------------------------------------------------ CLIP ViT tower oracle
@pytest.mark.parametrize("geo", [dict(image_size=56, patch_size=14, hidden=128, n_layers=3, n_heads=2, ffn=256),
                                 dict(image_size=70, patch_size=14, hidden=256, n_layers=4, n_heads=4, ffn=512)])
def test_clip_oracle_equals_the_live_transformers_model(geo):
    """oracle/clip_oracle.py restates transformers' CLIPVisionModel (the third-party network behind models/clip_encoder.py, pinned
    4.41.1 in the reference's requirements): pinned here to the LIVE library of this image on the same state_dict -- every entry of
    hidden_states to 1e-5, and the tower's feature selection (hidden_states[-2], CLS dropped) exactly as clip_encoder.py:29-37."""
    transformers = pytest.importorskip("transformers")
    from oracle import clip_oracle as CO
    d = CO.ClipDims(**geo)
    cfg = transformers.CLIPVisionConfig(hidden_size=d.hidden, intermediate_size=d.ffn, num_hidden_layers=d.n_layers, num_attention_heads=d.n_heads,
                                        image_size=d.image_size, patch_size=d.patch_size)
    assert cfg.hidden_act == "quick_gelu" and cfg.layer_norm_eps == d.ln_eps
    m = transformers.CLIPVisionModel(cfg).eval()
    W = CO.make_clip_weights(d, seed=2)
    assert set(k for k in m.state_dict() if "position_ids" not in k) == set(W)
    m.load_state_dict(W, strict=False)
    x = torch.randn(2, 3, d.image_size, d.image_size, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(x, output_hidden_states=True).hidden_states
        got = CO.hidden_states(x, W, d)
    assert len(ref) == len(got) == d.n_layers + 1
    for a, b in zip(ref, got):
        assert (a - b).abs().max().item() < 1e-5
    with torch.no_grad():
        f = CO.tower_features(x, W, d)
    assert f.shape == (2, d.n_tokens - 1, d.hidden) and torch.equal(f, got[-2][:, 1:])
    assert torch.equal(CO.tower_features(x, W, d, select_feature="cls_patch"), got[-2])
