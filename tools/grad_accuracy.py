#!/usr/bin/env python
"""Per-tensor gradient accuracy at full size: |g_gpu - g_f64| / |g_f64| next to the fp32 reference's own distance
from the float64 oracle, for ROBO_UNet(noScale) 2x3x240x320 and PB_FCN(noScale, bestModelSegVGA) 2x3x480x640.
    RCV_B200_MATH=parity|fp32|tf32|bf16 python tools/grad_accuracy.py [noscale|vga]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import synth
from nets import pb_fcn_state, robo_state
from oracle import ref_model as R

which = sys.argv[1] if len(sys.argv) > 1 else "noscale"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if which == "noscale":
    from robocupvision_b200.model import ROBO_UNet
    sd, kw, okw = robo_state("robo_noscale")
    fwd = lambda s, xx: R.robo_unet_forward(s, xx, training=True, **okw)
    m = ROBO_UNet(**kw); m.load_state_dict(sd)
    x = synth.images(2, 3, 240, 320, seed=6 + seed); y = synth.labels_learnable(x)
else:
    from robocupvision_b200.model import PB_FCN, load_legacy_state_dict
    osd_, raw = pb_fcn_state("bestModelSegVGA")
    fwd = lambda s, xx: R.pb_fcn_forward(s, xx, True, training=True)
    m = PB_FCN(32, 5, 1, True, 0); load_legacy_state_dict(m, raw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = synth.images(2, 3, 480, 640, seed=11 + seed); y = synth.labels_random(2, 480, 640, seed=12 + seed)
w = synth.CLASS_WEIGHTS

def oracle(dtype):
    o = R.leaf_state_dict({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()})
    loss = R.cross_entropy_2d(fwd(o, x.to(dtype)), y, torch.tensor(w, dtype=dtype))
    loss.backward()
    return {k: v.grad.double() for k, v in o.items() if v.grad is not None}
g64, g32 = oracle(torch.float64), oracle(torch.float32)
from robocupvision_b200.model import CrossEntropyLoss2d
import os
if os.environ.get("GA_BN64"):
    # diagnostic: BatchNorm backward in float64 torch ops ("A": statistics recomputed in float64 from z; "B": the
    # kernel's own fp32 mean / invstd) in place of rcv_bn_bwd_*: which rounding carries the full-size error?
    from robocupvision_b200 import engine, ops as _ops
    variant = os.environ["GA_BN64"]

    def bn_bwd64(order, dy, z, scale, shift, mean, invstd, dgamma=None, dbeta=None, dbias=None, want_dbias=False,
                 sums=None):
        g, zz = dy.double(), z.double()
        C = z.shape[1]
        v = lambda t: t.view(1, C, 1, 1)
        if variant == "A":
            mu = zz.mean((0, 2, 3)); var = zz.var((0, 2, 3), unbiased=False); istd = 1.0 / torch.sqrt(var + 1e-5)
        else:
            mu, istd = mean.double(), invstd.double()
        gamma = scale.double() / invstd.double()
        if order == _ops.EPI_AFFINE_RELU:
            g = g * (torch.addcmul(v(shift), v(scale), z) > 0)
        xh = (zz - v(mu)) * v(istd)
        s1, s2 = g.sum((0, 2, 3)), (g * xh).sum((0, 2, 3))
        n = g.numel() / C
        d = v(gamma * istd) * (g - v(s1 / n) - xh * v(s2 / n))
        if order == _ops.EPI_RELU_AFFINE:
            d = d * (z > 0)
        dgamma += s2.float(); dbeta += s1.float()
        if dbias is not None:
            dbias += d.sum((0, 2, 3)).float()
        return d.float().contiguous(), dgamma, dbeta, dbias
    engine.ops.bn_bwd = bn_bwd64
m.cuda().train()
loss = CrossEntropyLoss2d(torch.tensor(w)).cuda()(m(x.cuda()), y.cuda())
loss.backward()
gmax = max(float(v.norm()) for v in g64.values())
tot = [0.0, 0.0, 0.0]
print(f"{'tensor':48s} {'|g64|/gmax':>10s} {'gpu':>9s} {'fp32 ref':>9s}")
for k, p in m.named_parameters():
    if k not in g64:
        continue
    a = p.grad.cpu().double(); den = float(g64[k].norm())
    e, er = float((a - g64[k]).norm()), float((g32[k] - g64[k]).norm())
    tot[0] += e * e; tot[1] += er * er; tot[2] += den * den
    print(f"{k:48s} {den / gmax:10.2e} {e / max(den, 1e-30):9.2e} {er / max(den, 1e-30):9.2e}")
print(f"ALL {which}: gpu {(tot[0] / tot[2]) ** 0.5:.2e}  fp32 ref {(tot[1] / tot[2]) ** 0.5:.2e}")
