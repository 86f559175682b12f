"""Golden vectors for the TRAINING-side rows of the hot path (Showo.forward with labels, modeling_showo.py:81-100, and
its gradients): run the UNMODIFIED reference on the seeded mixed batch of fixtures.train_batch and store its outputs.

    python tests/golden/make_golden_train.py        (build container only; needs /root/reference)

train_step.npz: the three losses, a logits slice, per-parameter gradient norms of the weighted loss and small gradient
slices of the probe tensors (fixtures.TRAIN_GRAD_PROBES).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_loader as R  # noqa: E402
from oracle import showo_oracle as O  # noqa: E402
from fixtures import TINY, TRAIN_COEFF, TRAIN_GRAD_PROBES, train_batch  # noqa: E402

torch.set_num_threads(8)


def main():
    voc = O.ShowoVocab()
    dims = O.PhiDims(**TINY)
    W = O.make_showo_weights(dims, seed=3)
    model, _ = R.build_showo(dims, W)
    ids, mask, labels, (bt, bl, bm) = train_batch(voc)
    for p in model.parameters():
        p.requires_grad_(True)
    logits, l1, l2, l3 = model(ids, attention_mask=mask, labels=labels, batch_size_t2i=bt, batch_size_lm=bl,
                               batch_size_mmu=bm, max_seq_length=128)
    loss = TRAIN_COEFF[0] * l1 + TRAIN_COEFF[1] * l2 + TRAIN_COEFF[2] * l3
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    names = sorted(grads)
    out = dict(losses=np.array([l1.item(), l2.item(), l3.item()], dtype=np.float64),
               logits_slice=logits[:, ::32, ::997].detach().numpy().astype(np.float32),
               grad_names=np.array(names), grad_norms=np.array([grads[k].double().norm().item() for k in names]))
    for k in TRAIN_GRAD_PROBES:
        g = grads[k]
        out["grad:" + k] = (g[:8, :8] if g.dim() == 2 else g[:64]).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "train_step.npz"), **out)
    print("losses", out["losses"], "params with grad", len(names))


if __name__ == "__main__":
    main()
