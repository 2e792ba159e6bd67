#!/usr/bin/env python
"""Narrow-layer engine (fp32 FFMA2) vs the tensor-core engine forced (RCV_MATH_TF32X3) on the <= 16-channel layers."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import torch
from robocupvision_b200 import ops
from umma_probe import timeit
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for tr, k, s, cin, cout, h, w in [(False, 3, 1, 16, 16, 60, 80), (False, 3, 2, 8, 16, 120, 160), (True, 3, 2, 32, 16, 30, 40),
                                  (True, 3, 2, 16, 8, 60, 80), (False, 3, 1, 3, 8, 120, 160), (False, 3, 1, 8, 8, 120, 160),
                                  (False, 3, 2, 16, 32, 60, 80), (False, 3, 1, 32, 32, 30, 40)]:
    g = ops.ConvGeom(cin, cout, k, s, 1 if k == 3 else 0, 1, tr)
    x = torch.randn(B, cin, h, w, device="cuda")
    wt = torch.randn((cin, cout, 3, 3) if tr else (cout, cin, k, k), device="cuda") / (cin * k * k) ** 0.5
    ho, wo = g.out_hw(h, w)
    y = torch.empty(B, cout, ho, wo, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x)
    dw = torch.zeros_like(wt)
    out = []
    for math in (ops.MATH_AUTO, ops.MATH_TF32X3):
        res = []
        for d in (0, 1, 2):
            try:
                eng = ops.conv_engine(g, B, h, w, d, math)
                wp = ops.conv_pack(g, wt, d, math=math, nhw=(B, h, w)) if (d < 2 and eng == ops.ENGINE_UMMA) else None
                fn = [lambda: ops.conv_fwd(g, x, wt, None, epilogue=ops.EPI_RELU, out=y, math=math, wpacked=wp),
                      lambda: ops.conv_dgrad(g, dy, wt, (h, w), math=math, out=dx, wpacked=wp),
                      lambda: ops.conv_wgrad(g, x, dy, dw=dw, math=math)][d]
                res.append(f"{eng}:{timeit(fn):6.1f}")
            except Exception as e:  # noqa: BLE001
                res.append(f"x:{str(e)[:20]}")
        out.append(" ".join(res))
    print(f"{'convT' if tr else 'conv'} k{k} s{s} {cin:2d}->{cout:2d} @{h}x{w} x{B}: auto [fwd dgrad wgrad engine:us] {out[0]} | forced TC {out[1]}", flush=True)
