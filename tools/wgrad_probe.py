#!/usr/bin/env python
"""Time the weight-gradient kernels alone (CUDA graph of 20 launches, L2 warm): python tools/wgrad_probe.py [math]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import torch
from robocupvision_b200 import ops
from umma_probe import timeit
math = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for tr, s, cin, cout, dil, n, h, w in [(False, 1, 128, 128, 1, 64, 15, 20), (False, 1, 64, 128, 1, 64, 15, 20),
                                       (False, 1, 128, 64, 1, 64, 15, 20), (False, 1, 64, 64, 1, 64, 15, 20),
                                       (False, 1, 32, 32, 1, 64, 30, 40), (False, 2, 32, 64, 1, 64, 30, 40),
                                       (False, 2, 16, 32, 1, 64, 60, 80), (True, 2, 64, 32, 1, 64, 15, 20),
                                       (False, 1, 128, 128, 2, 64, 15, 20)]:
    g = ops.ConvGeom(cin, cout, 3, s, dil, dil, tr)
    x = torch.randn(n, cin, h, w, device="cuda")
    ho, wo = g.out_hw(h, w)
    dy = torch.randn(n, cout, ho, wo, device="cuda")
    dw = torch.zeros(g.weight_shape(), device="cuda")
    eng = ops.conv_engine(g, n, h, w, 2, math)
    t = timeit(lambda: ops.conv_wgrad(g, x, dy, dw=dw, math=math))
    fl = 2.0 * cin * cout * 9 * (h * w if tr else ho * wo) * n
    print(f"wgrad {'convT' if tr else 'conv'} s{s} d{dil} {cin:3d}->{cout:3d} {n}x{h}x{w}: engine {eng} {t:6.1f} us  {fl / (t * 1e-6) / 1e12:6.1f} TFLOP/s", flush=True)
