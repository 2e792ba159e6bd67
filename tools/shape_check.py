#!/usr/bin/env python
"""GPU check: every conv geometry of ROBO-UNet at a given input size, engine chosen by the dispatcher, against
the ATen CPU op (forward with a folded-BN epilogue + residual, input gradient, weight gradient) -- not a test."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.nn.functional as F
from robocupvision_b200 import ops
from test_gpu_ops import GEOMS, _ref_conv

H, W, N = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (24, 32, 2)))
layers = [("k3s1d1", 3, 8, 1), ("k3s2", 8, 16, 1), ("k3s1d1", 16, 16, 2), ("k3s2", 16, 32, 2), ("k3s1d1", 32, 32, 4),
          ("k3s2", 32, 64, 4), ("k3s1d1", 64, 64, 8), ("k3s1d1", 64, 128, 8), ("k3s1d1", 128, 128, 8),
          ("k3s1d1", 128, 64, 8), ("convT", 64, 32, 8), ("convT", 32, 16, 4), ("convT", 16, 8, 2), ("k1", 8, 5, 1)]
names = {0: "simt", 1: "direct", 2: "umma", 3: "narrow"}
for geom, cin, cout, div in layers:
    h, w = H // div, W // div
    k, s, p, d, tr = GEOMS[geom]
    g = ops.ConvGeom(cin, cout, k, s, p, d, tr)
    gen = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(N, cin, h, w, generator=gen)
    wt = torch.randn((cin, cout, 3, 3) if tr else (cout, cin, k, k), generator=gen) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=gen)
    sc, sh = torch.randn(cout, generator=gen), torch.randn(cout, generator=gen)
    xr = x.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True)
    v = _ref_conv(geom, xr, wr, b)
    res = torch.randn(v.shape, generator=gen)
    ref = sc.view(1, -1, 1, 1) * F.relu(v) + sh.view(1, -1, 1, 1) + res
    dy = torch.randn(v.shape, generator=gen)
    v.backward(dy)
    eng = ops.conv_engine(g, N, h, w, ops.PACK_FWD, ops.MATH_AUTO)
    wp = ops.conv_pack(g, wt.cuda(), ops.PACK_FWD, nhw=(N, h, w)) if eng == 2 else None
    got = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), epilogue=ops.EPI_RELU_AFFINE, scale=sc.cuda(), shift=sh.cuda(),
                       residual=res.cuda(), math=ops.MATH_AUTO, wpacked=wp)
    ef = float((got.cpu() - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    engd = ops.conv_engine(g, N, h, w, ops.PACK_DGRAD, ops.MATH_AUTO)
    wpd = ops.conv_pack(g, wt.cuda(), ops.PACK_DGRAD, nhw=(N, h, w)) if engd == 2 else None
    dx = ops.conv_dgrad(g, dy.cuda(), wt.cuda(), (h, w), math=ops.MATH_AUTO, wpacked=wpd)
    ed = float((dx.cpu() - xr.grad).abs().max()) / max(1.0, float(xr.grad.abs().max()))
    engw = ops.conv_engine(g, N, h, w, 2, ops.MATH_AUTO)
    dw, db = ops.conv_wgrad(g, x.cuda(), dy.cuda(), want_bias=True, math=ops.MATH_AUTO)
    ew = float((dw.cpu() - wr.grad).abs().max()) / max(1.0, float(wr.grad.abs().max()))
    flag = "  <<<<" if max(ef, ed, ew) > 2e-5 else ""
    print(f"{geom:7s} {cin:3d}->{cout:3d} {N}x{h}x{w}: fwd[{names[eng]}] {ef:.2e}  dgrad[{names[engd]}] {ed:.2e}  "
          f"wgrad[{names[engw]}] {ew:.2e}{flag}", flush=True)
