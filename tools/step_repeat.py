#!/usr/bin/env python
"""GPU check: run-to-run spread of 5 training steps (same seeds) -- separates atomic-order rounding noise from a race."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import synth
from robocupvision_b200.model import ROBO_UNet
from robocupvision_b200.train import TrainStep

xs = [synth.images(4, 3, 48, 64, seed=200 + s) for s in range(5)]
ys = [synth.labels_learnable(x) for x in xs]
runs = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    torch.manual_seed(12345678)
    m = ROBO_UNet().cuda()
    eps = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
    st = TrainStep(m, synth.CLASS_WEIGHTS, lr=1e-3, l1_decay=1e-6, eps=eps, use_graph=(rep % 2 == 0))
    losses = []
    for x, y in zip(xs, ys):
        st.step(x.cuda(), y.cuda())
        losses.append(st.loss_value())
    runs.append((losses, {k: v.detach().clone() for k, v in m.state_dict().items() if v.is_floating_point()}))
l0, p0 = runs[0]
for i, (l, p) in enumerate(runs[1:], 1):
    dl = [abs(a - b) / abs(a) for a, b in zip(l0, l)]
    worst = max(((float((p[k] - p0[k]).abs().max()) / max(1.0, float(p0[k].abs().max()))), k) for k in p0)
    print(f"run {i}: loss rel diff per step {['%.1e' % d for d in dl]}  worst param diff {worst[0]:.2e} ({worst[1]})", flush=True)
