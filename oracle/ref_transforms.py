"""TEST INFRASTRUCTURE (not product code): CPU restatement of the input-side augmentation and of the optional
Dice loss, the callers' steps either side of the hot path.  Pinned on
tests/golden/augment_dice.npz, which oracle/make_golden_aux.py produced by running the reference's own
dataset.ColorJitter and model.DiceLoss.

  normalize_flip_jitter   dataset.py:123-131 (SSYUVDataset.__getitem__, train split) + dataset.py:19-39
  dice_loss               model.py:5-43
"""
from __future__ import annotations

import math

import torch


def jitter_matrix(s_val: float, h_val: float) -> torch.Tensor:
    """dataset.py:32: torch.FloatTensor([[s cos h, -sin h], [sin h, s cos h]])."""
    return torch.tensor([[s_val * math.cos(h_val), -math.sin(h_val)], [math.sin(h_val), s_val * math.cos(h_val)]],
                        dtype=torch.float32)


def normalize_flip_jitter(img: torch.Tensor, label, flip: bool, b_val, c_val, s_val, h_val,
                          mean=(0.5, 0.0, 0.0), std=(0.5, 0.5, 0.5)):
    """One image [3,H,W] in [0,1] (to_tensor output) -> (image, label) as the train split returns them."""
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)
    img = (img - m) / s                                  # dataset.py:123 self.normalize(img)
    if flip:                                             # dataset.py:126-128
        img = img.flip(2)
        label = label.flip(1) if label is not None else None
    out = img.clone()
    out[0] = (img[0] + b_val) * c_val                    # dataset.py:34
    out[1:] = torch.einsum("nm,mbc->nbc", jitter_matrix(s_val, h_val), img[1:])  # dataset.py:36
    return out, label


def dice_loss(logits: torch.Tensor, true: torch.Tensor, weights: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """model.py:34-43 for num_classes > 1; `true` int64 [B,H,W]; weights as given to the constructor."""
    c = logits.shape[1]
    w = weights / weights.sum().item() * weights.shape[0]            # model.py:8
    onehot = torch.eye(c)[true.long()].permute(0, 3, 1, 2).float()   # model.py:34-35
    probas = torch.softmax(logits, dim=1)                            # model.py:36
    dims = (0, 2, 3)
    inter = torch.sum(probas * onehot, dims)                         # model.py:39
    card = torch.sum(probas + onehot, dims)                          # model.py:40
    return 1 - (2.0 * w * inter / (card + eps)).mean()               # model.py:41-42
