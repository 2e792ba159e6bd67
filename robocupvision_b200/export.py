"""Export wire format of the reference (paramSave.py:5-17): the state_dict flattened in key order
into one raw float64 file that an external engine reads next to a hand-written net.cfg.

Host-side code (the reference calls ``.numpy()`` on CPU tensors); CUDA tensors are copied back first.
``weightsLP/weights.dat`` in the reference tree is exactly this flatten of
``pth/bestModelLPFinetunedPruned.pth`` (tests/golden/weightsLP_head.npz pins head, tail, length, sha256).
"""
from __future__ import annotations

import os
from typing import Mapping

import numpy as np
import torch


def flatten_state_dict(state_dict: Mapping[str, torch.Tensor], skipClassifier: bool = False) -> np.ndarray:
    """float64 concatenation of every entry in key order (paramSave.py:9-16)."""
    parts = [np.empty(0)]
    for name, t in state_dict.items():
        if "classifier" in name and skipClassifier:
            print("Classifier module skipped")
            continue
        a = t.detach().cpu().numpy()
        parts.append(a.reshape(a.size).astype(np.float64, copy=False))
    return np.concatenate(parts)


def saveParams(path, model, fName: str = "weights.dat", skipClassifier: bool = False):
    """Drop-in for paramSave.saveParams (same arguments, same file)."""
    if not os.path.exists(path):
        os.makedirs(path)
    flatten_state_dict(model.state_dict(), skipClassifier).tofile(os.path.join(path, fName))
