"""Generate tests/golden/augment_dice.npz FROM THE REFERENCE (run here, where /root/reference exists):

    python -m oracle.make_golden_aux

dataset.ColorJitter (dataset.py:19-39) is run unmodified on seeded images with `random` seeded, recording the
scalars it drew; model.DiceLoss (model.py:5-43) is run unmodified with autograd for the logits gradient.
dataset.py imports skimage.color.rgb2yuv at module level, which is absent in this image and unused by
ColorJitter: a stub module stands in for that one import.
"""
from __future__ import annotations

import random
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(REF))


def main():
    if "skimage" not in sys.modules:
        try:
            import skimage.color  # noqa: F401
        except Exception:  # noqa: BLE001
            sk, skc = types.ModuleType("skimage"), types.ModuleType("skimage.color")
            skc.rgb2yuv = lambda x: x
            sk.color = skc
            sys.modules["skimage"], sys.modules["skimage.color"] = sk, skc
    import dataset as REFD
    import model as REFM
    import synth

    out = {}
    # ---- ColorJitter on normalised images; the reference draws b, c, s, h with random.uniform ----------
    g = torch.Generator().manual_seed(4242)
    imgs = torch.rand(6, 3, 12, 16, generator=g)                   # to_tensor range
    mean, std = torch.tensor([0.5, 0.0, 0.0]).view(3, 1, 1), torch.tensor([0.5, 0.5, 0.5]).view(3, 1, 1)
    labels = torch.randint(0, 5, (6, 12, 16), generator=g)
    flips = [False, True, False, True, True, False]
    random.seed(12345678)
    cj = REFD.ColorJitter()
    scal, outs, labs = [], [], []
    for i in range(6):
        state = random.getstate()
        b = random.uniform(-cj.b, cj.b); c = random.uniform(1 - cj.c, 1 + cj.c)
        s = random.uniform(1 - cj.s, 1 + cj.s); h = random.uniform(-cj.h, cj.h)
        random.setstate(state)                                     # the reference draws the same four
        img = (imgs[i] - mean) / std                               # transforms.Normalize (dataset.py:123)
        lab = labels[i]
        if flips[i]:                                               # dataset.py:126-128
            img = img.flip(2)
            lab = lab.flip(1)
        outs.append(cj(img.clone()))                               # dataset.py:129, the reference's code
        labs.append(lab)
        scal.append([b, c, s, h])
    out.update(aug_in=imgs.numpy(), aug_labels=labels.numpy(), aug_flip=np.array(flips), aug_scalars=np.array(scal),
               aug_out=torch.stack(outs).numpy(), aug_labels_out=torch.stack(labs).numpy())

    # ---- DiceLoss: value and gradient w.r.t. the logits --------------------------------------------------
    w = torch.tensor(synth.CLASS_WEIGHTS)
    logits = (torch.randn(3, 5, 24, 32, generator=g) * 3).requires_grad_(True)
    # (targets as the drivers pass them: int64 [B,H,W]; the docstring's [B,1,H,W] form indexes dim 4 of a 4-D
    # tensor at model.py:38-40 and raises IndexError under torch 2.11)
    true = torch.randint(0, 5, (3, 24, 32), generator=g)
    loss = REFM.DiceLoss(w)(logits, true)
    loss.backward()
    out.update(dice_logits=logits.detach().numpy(), dice_true=true.numpy(), dice_weights=w.numpy(),
               dice_loss=np.array(float(loss)), dice_grad=logits.grad.numpy())
    np.savez_compressed(OUT / "augment_dice.npz", **out)
    print("dice loss", float(loss), "jitter scalars", scal[0])


if __name__ == "__main__":
    main()
