#!/usr/bin/env python
"""Time the BatchNorm backward alone (CUDA graph of 20 calls): RCV_BN_BWD_FUSED=0|1 python tools/bn_probe.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import torch
from robocupvision_b200 import ops
from umma_probe import timeit
for n, c, h, w in [(64, 128, 15, 20), (64, 64, 15, 20), (64, 32, 30, 40), (64, 16, 60, 80), (64, 8, 120, 160), (64, 64, 30, 40)]:
    dy = torch.randn(n, c, h, w, device="cuda"); z = torch.randn(n, c, h, w, device="cuda")
    sc = torch.rand(c, device="cuda") + 0.5; sh = torch.randn(c, device="cuda"); mean = torch.randn(c, device="cuda") * 0.1
    istd = torch.rand(c, device="cuda") + 0.5
    dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda"); dbias = torch.zeros(c, device="cuda")
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    def fn():
        sums.zero_()
        ops.bn_bwd(ops.EPI_RELU_AFFINE, dy, z, sc, sh, mean, istd, dgamma=dg, dbeta=db, dbias=dbias, sums=sums)
    t = timeit(fn)
    print(f"bn_bwd {n}x{c}x{h}x{w}: {t:6.1f} us  ({3 * dy.numel() * 4 / t / 1e3:6.0f} GB/s algorithmic)", flush=True)
