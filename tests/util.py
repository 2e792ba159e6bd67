"""Shared helpers for the parity tests."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def assert_close(name, got, ref, tol=1e-5, atol=0.0):
    """max |got-ref| <= tol * max(1, max|ref|) + atol  (SURVEY.md section 8c gate shape)."""
    got, ref = got.detach().cpu(), ref.detach().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = max_err(got, ref)
    scale = max(1.0, float(ref.abs().max())) if ref.numel() else 1.0
    assert err <= tol * scale + atol, f"{name}: max err {err:.3e} > {tol:.1e} * {scale:.3e}"
    return err


def load_ckpt(name):
    z = np.load(GOLDEN / "ckpt" / (name + ".npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def load_golden(name):
    z = np.load(GOLDEN / (name + ".npz"))
    return {k: z[k] for k in z.files}


def with_nbt(sd):
    """Add the num_batches_tracked entries pre-0.4.1 checkpoints lack."""
    out = dict(sd)
    for k in list(sd):
        if k.endswith("running_mean"):
            out.setdefault(k[: -len("running_mean")] + "num_batches_tracked", torch.zeros((), dtype=torch.long))
    return out
