"""-m gpu: the device-side producer of the t2i training rows (showo_t2i_train_prep through show-o_b200/train_inputs.py) against
the unmodified reference's outputs on the same uniform draws (tests/golden/train_prep.npz) -- integer outputs, bit-exact --
and, with the library's own Philox noise, against the properties the reference's code guarantees."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from showo_b200 import train_inputs as TI
from showo_b200.schedules import cosine_schedule, get_mask_chedule

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    spec = importlib.util.spec_from_file_location("make_golden_prep", os.path.join(HERE, "golden", "make_golden_prep.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg, np.load(os.path.join(HERE, "golden", "train_prep.npz"))


class Cfg(dict):
    __getattr__ = dict.get


@pytest.mark.parametrize("name", ["a256", "b1024", "c_trunc"])
def test_train_rows_bit_exact_against_reference_golden(name):
    mg, z = _golden()
    dev = torch.device("cuda", 0)
    N, T, rate, drop, texts, codes, _ = mg.case_inputs(name)
    cfg = Cfg(training=Cfg(min_masking_rate=rate, noise_type="mask"))
    up = TI.UniversalPrompting(mg.FakeTokenizer(), max_text_len=T, ignore_id=-100, cond_dropout_prob=drop)
    t = lambda k: torch.from_numpy(z[f"{name}_{k}"]).to(dev)      # noqa: E731
    # the two reference calls, separately (train.py:479-486)
    ids_img, lab_img, lw, mp = TI.mask_or_random_replace_tokens(codes.to(dev), 58497, cfg, cosine_schedule, noise=(t("timesteps"), t("rand")))
    assert lw is None
    assert torch.equal(ids_img, t("ids_img")) and torch.equal(lab_img, t("lab_img"))
    assert (mp - t("mask_prob")).abs().max().item() <= 1e-6
    ids, masks, labels = up.t2i_prompt(texts, ids_img, lab_img, probs=t("probs"))
    assert torch.equal(ids, t("ids")) and torch.equal(labels, t("labels")) and torch.equal(masks, t("masks"))
    # fused: one launch, plus the mask descriptors of the rows
    ids2, labels2, mp2, descs = up.t2i_train_rows(texts, codes.to(dev), 58497, cfg, cosine_schedule, noise=(t("timesteps"), t("rand"), t("probs")))
    assert torch.equal(ids2, t("ids")) and torch.equal(labels2, t("labels")) and torch.equal(mp2, mp)
    assert torch.equal(descs.cpu(), torch.from_numpy(z[f"{name}_descs"]))
    # an arbitrary python schedule evaluated by the caller gives the same rows when it is the cosine
    ids3, labels3, _, _ = up.t2i_train_rows(texts, codes.to(dev), 58497, cfg, lambda x: torch.cos(x * np.pi * 0.5),
                                           noise=(t("timesteps"), t("rand"), t("probs")))
    assert torch.equal(ids3, t("ids")) and torch.equal(labels3, t("labels"))


@pytest.mark.parametrize("sched", ["cosine", "linear", "pow2"])
def test_train_rows_philox_properties(sched):
    """library noise: every row masks exactly max(1, round(N * max(schedule(t), min_rate))) positions, labels carry the code exactly at
    the masked positions, unmasked codes pass through, text part / specials / descriptors are consistent."""
    mg, _ = _golden()
    dev = torch.device("cuda", 0)
    N, T, _, _, texts, codes, _ = mg.case_inputs("b1024")
    B = codes.shape[0]
    cfg = Cfg(training=Cfg(min_masking_rate=0.25, noise_type="mask"))
    up = TI.UniversalPrompting(mg.FakeTokenizer(), max_text_len=T, ignore_id=-100, cond_dropout_prob=0.5)
    torch.manual_seed(7)
    ids, labels, mp, descs = up.t2i_train_rows(texts, codes.to(dev), 58497, cfg, get_mask_chedule(sched))
    P, L = T + 1, T + 1 + N + 2
    assert ids.shape == (B, L) and labels.shape == (B, L)
    img, lab = ids[:, P + 1:L - 1].cpu(), labels[:, P + 1:L - 1].cpu()
    masked = img == 58497
    n = torch.clamp(torch.round(N * mp.cpu()), min=1).long()
    assert torch.equal(masked.sum(1), n) and bool((mp.cpu() >= 0.25).all()) and bool((mp.cpu() <= 1).all())
    assert torch.equal(lab[masked], codes[masked]) and bool((lab[~masked] == -100).all()) and torch.equal(img[~masked], codes[~masked])
    assert bool((ids[:, P] == 50296).all()) and bool((ids[:, L - 1] == 50297).all())
    d = descs.cpu()
    for b in range(B):
        row = ids[b].cpu()
        pads = int((row == 50295).sum())
        assert d[b].tolist() == [pads, P, L, 0, 0] and bool((row[:pads] == 50295).all()) and int(row[pads]) == 50300
        assert bool((labels[b, :pads] == -100).all()) and torch.equal(labels[b, pads:P].cpu(), row[pads:P])
    # a second call with another seed gives another mask
    torch.manual_seed(8)
    ids_b, _, _, _ = up.t2i_train_rows(texts, codes.to(dev), 58497, cfg, get_mask_chedule(sched))
    assert not torch.equal(ids_b, ids)


def test_train_rows_refuse_what_the_device_path_does_not_provide():
    mg, _ = _golden()
    dev = torch.device("cuda", 0)
    _, _, _, _, _, codes, _ = mg.case_inputs("c_trunc")
    with pytest.raises(NotImplementedError):
        TI.mask_or_random_replace_tokens(codes.to(dev), 58497, Cfg(training=Cfg(mask_contiguous_region_prob=0.5)), cosine_schedule)
    with pytest.raises(Exception):
        TI.mask_or_random_replace_tokens(codes, 58497, Cfg(training=Cfg()), cosine_schedule)       # CPU tensor: no fallback
