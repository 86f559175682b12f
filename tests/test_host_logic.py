"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the product path fails
loudly without a GPU (no CPU fallback), the Python drop-in keeps the reference's surface, and the data-parallel
plumbing works at world_size 2 on gloo."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import showo_b200
from showo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_header_symbol():
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 26 and "showo_t2i_generate" in syms and "magvit_decode_code" in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(syms) == set(_lib._PROTOS.keys()), set(syms) ^ set(_lib._PROTOS.keys())
    assert lib.showo_abi_version() == 1
    # no link-time dependency on the driver library (must dlopen on a CPU-only box)
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_product_path_fails_loudly_without_gpu():
    lib = _lib.load()
    assert lib.showo_device_count() == 0
    with pytest.raises(_lib.ShowoError, match="no CPU fallback"):
        _lib.require_gpu()
    cfg = _lib.Config(58498, 2048, 24, 32, 8192, 32, 2048, 1e-5, 10000.0, 50295, 10, 8192)
    h = C.c_void_p()
    assert lib.showo_engine_create(C.byref(cfg), 0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.showo_last_error()
    hm = C.c_void_p()
    assert lib.magvit_engine_create(0, C.byref(hm)) != 0
    m = showo_b200.Showo(False, 58498, 50295, phi_dims=dict(hidden=256, n_layers=1, n_heads=4, ffn=256))
    with pytest.raises(_lib.ShowoError):
        m(torch.zeros(1, 8, dtype=torch.int64))
    with pytest.raises(_lib.ShowoError):
        showo_b200.MAGVITv2(materialize=False).decode_code(torch.zeros(1, 256, dtype=torch.int64))


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "show-o_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_showo_surface_matches_reference_names():
    import inspect
    m = showo_b200.Showo(False, 58498, 50295, phi_dims=dict(hidden=128, n_layers=2, n_heads=2, ffn=256))
    assert m.config.mask_token_id == 58497 and m.output_size == 58498 and m.vocab_size == 58498
    keys = set(m.state_dict().keys())
    for k in ("showo.model.embed_tokens.weight", "showo.model.layers.1.self_attn.q_proj.weight",
              "showo.model.layers.0.self_attn.k_layernorm.bias", "showo.model.layers.1.mlp.fc2.bias",
              "showo.model.layers.0.input_layernorm.weight", "showo.model.final_layernorm.bias", "showo.lm_head.bias"):
        assert k in keys, k
    assert m.state_dict()["showo.lm_head.weight"].shape == (58498, 128)
    fwd = list(inspect.signature(m.forward).parameters)
    assert fwd[:8] == ["input_ids", "input_embeddings", "attention_mask", "labels", "label_smoothing", "batch_size_t2i",
                       "batch_size_lm", "batch_size_mmu"]
    t2i = inspect.signature(m.t2i_generate).parameters
    assert list(t2i)[:9] == ["input_ids", "uncond_input_ids", "attention_mask", "temperature", "timesteps",
                             "guidance_scale", "noise_schedule", "generator", "config"]
    assert t2i["timesteps"].default == 18 and t2i["guidance_scale"].default == 0 and t2i["temperature"].default == 1.0
    mmu = inspect.signature(m.mmu_generate).parameters
    assert list(mmu)[:7] == ["idx", "input_embeddings", "attention_mask", "max_new_tokens", "temperature", "top_k", "eot_token"]
    e = m.showo.model.embed_tokens(torch.tensor([[1, 2]]))           # called from outside by inference_mmu.py:134
    assert e.shape == (1, 2, 128)
    m.showo.resize_token_embeddings(58500)
    assert m.showo.lm_head.weight.shape == (58500, 128)
    w = showo_b200.Showo(True, 58498, 50295, phi_dims=dict(hidden=128, n_layers=1, n_heads=2, ffn=256))
    assert "mm_projector.0.weight" in w.state_dict() and w.state_dict()["mm_projector.2.weight"].shape == (2048, 2048)


def test_magvit_surface():
    vq = showo_b200.MAGVITv2()
    sd = vq.state_dict()
    assert sum(v.numel() for v in sd.values()) == 55439171 + 39947833
    assert sd["decoder.up.3.block.0.nin_shortcut.weight"].shape == (256, 512, 1, 1)
    assert sd["encoder.down.0.downsample.conv.weight"].shape == (128, 128, 3, 3)
    assert vq._grid(torch.zeros(1, 1024), None) == (32, 32) and vq._grid(torch.zeros(1, 512), (16, 32)) == (16, 32)


def test_shard_rows_partitions_the_batch():
    from showo_b200.parallel import shard_rows
    for n, w in ((64, 8), (10, 4), (3, 8), (16, 1)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
import showo_b200
from showo_b200 import parallel as P
rank, world = P.init(backend="gloo")
b, e = P.shard_rows(8, rank, world)
local = torch.full((e - b, 4, 4, 3), rank + 1, dtype=torch.uint8)
local[:, 0, 0, 0] = torch.arange(b, e, dtype=torch.uint8)
allimg = P.gather_rows(local, world)
assert allimg.shape == (8, 4, 4, 3), allimg.shape
assert allimg[:, 0, 0, 0].tolist() == list(range(8))
assert allimg[:4, 1, 1, 1].eq(1).all() and allimg[4:, 1, 1, 1].eq(2).all()
t = P.max_over_ranks(float(rank + 1) * 1.5, torch.device("cpu"))
assert t == 3.0, t
codes = P.gather_rows(torch.full((e - b, 16), rank, dtype=torch.int64), world)
assert codes.shape == (8, 16) and codes[:, 0].tolist() == [0] * 4 + [1] * 4
torch.distributed.barrier()
torch.distributed.destroy_process_group()
os.write(1, ("rank %d ok\n" % rank).encode())      # one write(2) per rank: the two ranks share the pipe
"""


def test_data_parallel_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_bench_reference_arm_contract_offline():
    """`bench.py --impl reference` on non-zero ranks exits 0 without work (driver launches it under torchrun)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_from_pretrained_reads_single_sharded_and_bin_checkpoints(tmp_path):
    """Showo / MAGVITv2.from_pretrained on the layouts save_pretrained and the hub produce (inference_t2i.py:61-68): one
    safetensors file, a sharded safetensors set with an index, a torch .bin; a checkpoint that lacks a parameter is refused."""
    import json
    from safetensors.torch import save_file
    kw = dict(hidden=128, n_layers=2, n_heads=2, ffn=256)
    m = showo_b200.Showo(False, 58498, 50295, phi_dims=kw)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    cfg = {"_class_name": "Showo", "_diffusers_version": "0.30.1", "w_clip_vit": False, "vocab_size": 58498, "llm_vocab_size": 50295,
           "llm_model_path": "microsoft/phi-1_5", "codebook_size": 8192, "num_vq_tokens": 256, "load_from_showo": False}

    def check(d):
        m2 = showo_b200.Showo.from_pretrained(str(d), phi_dims=kw)
        sd2 = m2.state_dict()
        assert sd2.keys() == sd.keys() and all(torch.equal(sd[k], sd2[k]) for k in sd)
        assert m2.config.mask_token_id == 58497

    d1 = tmp_path / "single"; d1.mkdir()
    json.dump(cfg, open(d1 / "config.json", "w"))
    save_file({**sd, "showo.model.layers.0.self_attn.rotary_emb.inv_freq": torch.ones(16)}, str(d1 / "diffusion_pytorch_model.safetensors"))
    check(d1)
    d2 = tmp_path / "sharded"; d2.mkdir()
    json.dump(cfg, open(d2 / "config.json", "w"))
    names = sorted(sd)
    shards = {"diffusion_pytorch_model-00001-of-00002.safetensors": names[: len(names) // 2],
              "diffusion_pytorch_model-00002-of-00002.safetensors": names[len(names) // 2:]}
    for fn, ks in shards.items():
        save_file({k: sd[k] for k in ks}, str(d2 / fn))
    json.dump({"metadata": {}, "weight_map": {k: fn for fn, ks in shards.items() for k in ks}},
              open(d2 / "diffusion_pytorch_model.safetensors.index.json", "w"))
    check(d2)
    d3 = tmp_path / "bin"; d3.mkdir()
    json.dump(cfg, open(d3 / "config.json", "w"))
    torch.save(sd, str(d3 / "pytorch_model.bin"))
    check(d3)
    d4 = tmp_path / "broken"; d4.mkdir()
    json.dump(cfg, open(d4 / "config.json", "w"))
    save_file({k: v for k, v in sd.items() if k != "showo.lm_head.bias"}, str(d4 / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(KeyError, match="lm_head.bias"):
        showo_b200.Showo.from_pretrained(str(d4), phi_dims=kw)
    # MAGVIT-v2
    vq = showo_b200.MAGVITv2()
    vsd = {k: v.detach().clone() for k, v in vq.state_dict().items()}
    dv = tmp_path / "vq"; dv.mkdir()
    save_file(vsd, str(dv / "diffusion_pytorch_model.safetensors"))
    vq2 = showo_b200.MAGVITv2.from_pretrained(str(dv))
    assert all(torch.equal(vsd[k], v) for k, v in vq2.state_dict().items())
    first = sorted(vsd)[0]
    save_file({k: v for k, v in vsd.items() if k != first}, str(dv / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(KeyError):
        showo_b200.MAGVITv2.from_pretrained(str(dv))
    # save_pretrained -> from_pretrained round trip, single file and forced sharding
    for shard_bytes in (5 << 30, 200_000):
        dr = tmp_path / f"rt{shard_bytes}"
        m.save_pretrained(str(dr), max_shard_bytes=shard_bytes)
        assert (shard_bytes > 10 ** 6) == os.path.exists(dr / "diffusion_pytorch_model.safetensors")
        check(dr)


def test_clip_vision_tower_surface_without_gpu():
    """`from models import ... CLIPVisionTower` surface (models/clip_encoder.py:6-82) of the engine-backed tower: construction from a
    CLIPVisionConfig / dims, the reference's properties, weights kept under CLIPVisionModel.state_dict() names -- and no CPU fallback:
    the forward raises without an sm_100 device."""
    from transformers import CLIPVisionConfig
    from oracle import clip_oracle as CO
    from showo_b200 import CLIPVisionTower, ShowoError
    cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56, patch_size=14)
    tower = CLIPVisionTower(cfg)
    assert tower.num_patches == 16 and tower.num_patches_per_side == 4 and tower.hidden_size == 128 and tower.config is cfg
    assert tower.select_layer == -2 and tower.select_feature == "patch" and not tower.is_loaded
    assert tower.dummy_feature.shape == (1, 128) and tower.dtype == torch.float32
    assert all(not p.requires_grad for p in tower.parameters())
    tower.load_weights(CO.make_clip_weights(CO.ClipDims(image_size=56, patch_size=14, hidden=128, n_layers=3, n_heads=2, ffn=256)))
    assert tower.is_loaded
    if not torch.cuda.is_available():
        with pytest.raises(ShowoError):
            tower(torch.randn(1, 3, 56, 56))
    with pytest.raises(ValueError):                       # head_dim must be 64 on the engine
        CLIPVisionTower(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, image_size=28, patch_size=14))
    d = CLIPVisionTower(dict(image_size=336, patch_size=14, hidden=1024, n_layers=24, n_heads=16, ffn=4096))
    assert d.num_patches == 576 and d.config.hidden_size == 1024


def test_train_inputs_host_side_contract():
    """show-o_b200/train_inputs.py without a GPU: the schedule -> kernel code mapping, the refusal of what the device path does not
    provide (python-`random` branches of the reference), and no CPU fallback for the producer itself."""
    import torch
    from showo_b200 import ShowoError, get_mask_chedule, train_inputs as TI

    class Cfg(dict):
        __getattr__ = dict.get
    assert TI._schedule_code(get_mask_chedule("cosine")) == (0, 0.0)
    assert TI._schedule_code(get_mask_chedule("linear")) == (1, 0.0)
    assert TI._schedule_code(get_mask_chedule("pow2.5")) == (2, 2.5)
    assert TI._schedule_code(get_mask_chedule("sigmoid"))[0] == 3 and TI._schedule_code(lambda t: t)[0] == 3
    codes = torch.zeros(2, 16, dtype=torch.int64)
    with pytest.raises(NotImplementedError):
        TI.mask_or_random_replace_tokens(codes, 99, Cfg(training=Cfg(mask_contiguous_region_prob=0.3)), get_mask_chedule("cosine"))
    with pytest.raises(NotImplementedError):
        TI.mask_or_random_replace_tokens(codes, 99, Cfg(training=Cfg(predict_all_tokens=True)), get_mask_chedule("cosine"))
    with pytest.raises(NotImplementedError):
        TI.mask_or_random_replace_tokens(codes, 99, Cfg(training=Cfg(eval_mask_ratios=[0.5])), get_mask_chedule("cosine"), is_train=False)
    with pytest.raises(ShowoError):                       # CPU tensors: the producer only exists on the device
        TI.mask_or_random_replace_tokens(codes, 99, Cfg(training=Cfg()), get_mask_chedule("cosine"))


# ------------------------------------------------------------------------------------------------ f-1: inpainting / extrapolation
def _editing_case(name):
    import editing_stubs as ES
    from showo_b200 import editing, train_inputs as TI
    case = ES.CASES[name]
    cfg = ES.make_config(case)
    up = TI.UniversalPrompting(ES.FakeTokenizer(), max_text_len=128, ignore_id=-100, cond_dropout_prob=0.1)
    model, vq = ES.StubShowo(cfg.model.showo.num_vq_tokens), ES.StubVQ()
    img = ES.pixels(case["seed"], case["R"])
    if case["mode"] == "inpainting":
        editing.inpaint(model, vq, up, [case["prompt"]] * case["B"], img, ES.mask_pixels(case["seed"] + 100, case["R"]), cfg, mask_token_id=ES.MASK_ID)
    else:
        prompts = [p for p in case["prompt"].split(" *** ") if p]
        dirs = [d for d in case["direction"].split(" *** ") if d]
        editing.extrapolate(model, vq, up, prompts, dirs, img, cfg, offset=case["offset"], mask_token_id=ES.MASK_ID)
    return model, vq


@pytest.mark.parametrize("name", ["inpaint_cfg", "inpaint_nocfg", "extra_right_right", "extra_left_left", "extra_up", "extra_up_up"])
def test_editing_flows_equal_the_reference_script_blocks(name):
    """show-o_b200/editing.py (inpainting, extrapolation) against tests/golden/editing.npz = the reference's own script lines
    (inference_t2i.py:80-284) executed on the same stub models: every t2i_generate call receives the same input ids, unconditional ids
    and omni-mask descriptors, and decode_code the same token grid and shape, bit for bit."""
    import fixtures as FX
    z = FX.load("editing.npz")
    model, vq = _editing_case(name)
    assert len(model.calls) == int(z[f"{name}_n_calls"])
    for i, c in enumerate(model.calls):
        assert np.array_equal(c["input_ids"].numpy(), z[f"{name}_ids{i}"]), (name, i)
        if f"{name}_uncond{i}" in z.files:
            assert np.array_equal(c["uncond_input_ids"].numpy(), z[f"{name}_uncond{i}"]), (name, i)
        else:
            assert c["uncond_input_ids"] is None
        assert np.array_equal(np.asarray(c["attention_mask"], dtype=np.int32), z[f"{name}_descs{i}"]), (name, i)
        assert c["kw"]["seq_len"] == model.N and c["kw"]["timesteps"] == 4
    ids, shape = vq.decoded[-1]
    assert np.array_equal(ids.numpy(), z[f"{name}_decoded"])
    assert tuple(int(v) for v in z[f"{name}_shape"]) == (tuple(shape) if shape is not None else (-1, -1))


def test_editing_downward_extrapolation_runs_where_the_reference_script_raises():
    """inference_t2i.py:274-275 glues `image_left_part` along the row axis for the downward direction: the golden records that the
    script raises; editing.extrapolate implements the evident intent (mirror image of 'up') and is checked against it by symmetry."""
    import editing_stubs as ES
    import fixtures as FX
    from showo_b200 import editing
    assert int(FX.load("editing.npz")["down_raises"]) == 1
    g = torch.arange(2 * 8 * 8).reshape(2, 8, 8) % 8192
    mid = 777
    for off in (0, 1):
        up_c = editing.extrapolation_canvas(g, "up", off, mid, 8)
        dn_c = editing.extrapolation_canvas(g.flip(1), "down", off, mid, 8)
        assert torch.equal(up_c, dn_c.flip(1))
        gen = (torch.arange(2 * 64).reshape(2, 8, 8) * 3) % 8192
        assert torch.equal(editing.extrapolation_merge(g, gen, "up", off, 8), editing.extrapolation_merge(g.flip(1), gen.flip(1), "down", off, 8).flip(1))
        assert editing.extrapolation_merge(g, gen, "down", off, 8).shape == (2, 8 + 4 + off, 8)


def test_clip_vision_tower_loads_a_local_checkpoint_like_the_reference(tmp_path):
    """CLIPVisionTower(name) as inference_mmu.py:73-74 constructs it (`name` = hub id or local directory): `transformers` reads config,
    weights and the image processor; the weights land under CLIPVisionModel.state_dict() names ready for the engine, and the oracle on
    those weights equals the library model that wrote the checkpoint."""
    transformers = pytest.importorskip("transformers")
    from oracle import clip_oracle as CO
    from showo_b200 import CLIPVisionTower
    cfg = transformers.CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=28, patch_size=14)
    torch.manual_seed(4)
    ref = transformers.CLIPVisionModel(cfg).eval()
    ref.save_pretrained(str(tmp_path))
    transformers.CLIPImageProcessor(size={"shortest_edge": 28}, crop_size={"height": 28, "width": 28}).save_pretrained(str(tmp_path))
    tower = CLIPVisionTower(str(tmp_path))
    assert tower.is_loaded and tower.image_processor is not None and tower.vision_tower_name == str(tmp_path)
    assert tower.num_patches == 4 and tower.hidden_size == 128 and tower.config.num_hidden_layers == 2
    sd = {k: v for k, v in ref.state_dict().items() if "position_ids" not in k}
    assert set(tower._weights) == set(sd) and all(torch.equal(tower._weights[k], sd[k]) for k in sd)
    x = torch.randn(2, 3, 28, 28)
    d = CO.ClipDims(image_size=28, patch_size=14, hidden=128, n_layers=2, n_heads=2, ffn=256)
    with torch.no_grad():
        want = ref(x, output_hidden_states=True).hidden_states[-2][:, 1:]
        got = CO.tower_features(x, tower._weights, d)
    assert (got - want).abs().max().item() < 1e-5
    tower.load_model()          # "already loaded" like the reference, no reload
