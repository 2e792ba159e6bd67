"""Deterministic synthetic inputs shared by the oracle, the golden generator and the tests
(SURVEY.md section 8d): CPU torch generators, so every box sees the same tensors."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def images(n, c, h, w, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, c, h, w, generator=g)


def labels_random(n, h, w, num_classes=5, seed=4321):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, num_classes, (n, h, w), generator=g, dtype=torch.int64)


def labels_learnable(x, num_classes=5):
    """clamp(avg_pool5x5(x[:,0])*3 + 2, 0, C-1): a rule the nets can learn (loss-curve runs)."""
    s = F.avg_pool2d(x[:, :1], 5, 1, 2)[:, 0]
    return (s * 3 + 2).clamp(0, num_classes - 1).long()


CLASS_WEIGHTS = [1.0, 10.0, 30.0, 10.0, 2.0]      # train.py:309
LP_CLASS_WEIGHTS = [1.0, 6.0, 1.0, 3.0, 2.0]      # labelPropTrain.py:94
