#!/usr/bin/env python
"""GPU check: the narrow-layer engine is bitwise repeatable (static tile -> CTA assignment) -- not a test."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from robocupvision_b200 import ops
from test_gpu_ops import GEOMS

for geom, cin, cout, n, h, w in [("k3s1d1", 3, 8, 4, 48, 64), ("k3s2", 8, 16, 4, 48, 64), ("k3s1d1", 16, 16, 4, 24, 32),
                                  ("convT", 32, 16, 4, 12, 16), ("convT", 16, 8, 4, 24, 32), ("k1", 8, 5, 4, 48, 64),
                                  ("k3s1d1", 16, 16, 64, 60, 80)]:
    k, s, p, d, tr = GEOMS[geom]
    g = ops.ConvGeom(cin, cout, k, s, p, d, tr)
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn((cin, cout, 3, 3) if tr else (cout, cin, k, k), device="cuda")
    b = torch.randn(cout, device="cuda")
    outs, sts = [], []
    for rep in range(6):
        st = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        y = ops.conv_fwd(g, x, wt, b, epilogue=ops.EPI_RELU, stats=st, math=ops.MATH_AUTO)
        outs.append(y.clone()); sts.append(st.clone())
    torch.cuda.synchronize()
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    dst = max(float((sts[0] - s_).abs().max() / sts[0].abs().max()) for s_ in sts[1:])
    print(f"{geom} {cin}->{cout} {n}x{h}x{w}: engine {ops.conv_engine(g, n, h, w)} outputs identical {same}, stats rel diff {dst:.2e}", flush=True)
