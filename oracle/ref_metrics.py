"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Restatement (numpy) of the argmax / pixel-accuracy / per-image confusion / IoU loop the
reference copy-pastes into every driver (train.py:128-163, test.py:125-169, tester.py:146-173,
labelPropTrain.py:266-288, validLabelProp.py:135-164).  Integer counts are exact.
"""
from __future__ import annotations

import numpy as np


def argmax_first(logits: np.ndarray) -> np.ndarray:
    """torch.max(pred, 1)[1]: lowest class index among equal maxima (train.py:128)."""
    return np.argmax(logits, axis=1).astype(np.int64)


def confusion_per_image(pred: np.ndarray, target: np.ndarray, num_classes: int) -> np.ndarray:
    """conf[n, p, l] = #pixels of image n with pred == p and target == l (train.py:142-147)."""
    n = pred.shape[0]
    out = np.zeros((n, num_classes, num_classes), dtype=np.int64)
    for i in range(n):
        p = pred[i].reshape(-1).astype(np.int64)
        t = target[i].reshape(-1).astype(np.int64)
        ok = (p >= 0) & (p < num_classes) & (t >= 0) & (t < num_classes)
        out[i] = np.bincount(p[ok] * num_classes + t[ok], minlength=num_classes * num_classes).reshape(
            num_classes, num_classes)
    return out


def iou_sums(conf: np.ndarray) -> np.ndarray:
    """Sum over images of per-class IoU with the reference's union==0 -> 1 rule (train.py:148-153).
    union_c = rowsum_c + colsum_c - conf[c, c]."""
    n, c, _ = conf.shape
    out = np.zeros(c, dtype=np.float64)
    for i in range(n):
        for k in range(c):
            inter = conf[i, k, k]
            union = conf[i, k, :].sum() + conf[i, :, k].sum() - inter
            out[k] += 1.0 if union == 0 else inter / union
    return out


def epoch_summary(conf_total: np.ndarray, iou_sum: np.ndarray, img_cnt: int):
    """meanClassAcc, meanIoU, score exactly as train.py:157-164 computes them."""
    c = conf_total.shape[0]
    lab_cnts = conf_total.sum(axis=0).astype(np.float64)  # per label
    conf_pct = conf_total.astype(np.float64) / (lab_cnts[None, :] / 100.0)
    mean_class_acc = sum(conf_pct[j, j] for j in range(c)) / c
    mean_iou = float((iou_sum / img_cnt).sum()) / c * 100
    return mean_class_acc, mean_iou, (mean_class_acc + mean_iou) / 2
