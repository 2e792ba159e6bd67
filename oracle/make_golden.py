"""Generate tests/golden/* FROM THE REFERENCE (run here, where /root/reference exists):

    python -m oracle.make_golden

  ckpt/<name>.npz        released checkpoints (weights are data, not source) as fp32 arrays
  <name>_eval.npz        reference-class outputs on seeded synthetic frames (eval mode)
  robo_train.npz         3 reference training steps (train.py:43-74 restated with the reference's
                         own ROBO_UNet / CrossEntropyLoss2d / torch.optim.Adam): losses, grad norms
  weightsLP_head.npz     first/last values + sha256 of weightsLP/weights.dat (paramSave.py format)
  robo_curve200.npz      200 reference training steps at two sizes (`... curve` alone regenerates it)
  bestModelSegFinetunedPruned_bu*   the channel-pruned legacy PB_FCN (reference blocks + PB_FCN forward; `... bu` alone
                         regenerates just this one)
Everything is produced by importing /root/reference/model.py unmodified (plus the LabelProp
constructor shim of SURVEY.md section 8c, since the shipped constructor raises TypeError).
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(REF))

import model as REFM  # noqa: E402  (the reference)
import synth  # noqa: E402


def load_pth(name):
    return torch.load(REF / "pth" / name, map_location="cpu", weights_only=False)


def save_ckpt(name, sd):
    np.savez_compressed(OUT / "ckpt" / (name + ".npz"), **{k: v.numpy() for k, v in sd.items()})


def ref_labelprop():
    """Reference LabelProp with the constructor's surplus 8th argument dropped."""
    orig = REFM.ConvPoolSimple.__init__

    def shim(self, inplanes, planes, size, stride, padding, dilation, bias, dropout=None):
        orig(self, inplanes, planes, size, stride, padding, dilation, bias)

    REFM.ConvPoolSimple.__init__ = shim
    try:
        return REFM.LabelProp(5, 32, 0)
    finally:
        REFM.ConvPoolSimple.__init__ = orig


def labelprop_ref_forward(m, x):
    """model.py:556-567 with line 565 written out of place (identical arithmetic)."""
    top = m.pre(x)
    middle = m.down1(top)
    bottom = m.down2(middle)
    a = m.conv3(m.conv2(m.conv1(m.down3(bottom))))
    a = bottom + m.upConv1(a)
    a = middle + m.upConv2(a)
    a = m.upConv3(a)
    a = torch.cat([a[:, 0:8] + top, a[:, 8:]], 1)
    return m.classifier(a)


def eval_golden(tag, model, fwd, shapes, num_classes=5, weights=synth.CLASS_WEIGHTS):
    out = {}
    crit = REFM.CrossEntropyLoss2d(torch.tensor(weights))
    model.eval()
    for i, (n, c, h, w) in enumerate(shapes):
        x = synth.images(n, c, h, w, seed=1234 + i)
        y = synth.labels_random(n, h, w, num_classes, seed=4321 + i)
        with torch.no_grad():
            logits = fwd(model, x)
            loss = crit(logits, y)
        pred = logits.argmax(1)
        conf = np.zeros((n, num_classes, num_classes), dtype=np.int64)
        for b in range(n):  # train.py:142-147
            for l in range(num_classes):
                for p in range(num_classes):
                    conf[b, p, l] = int(((pred[b] == p) & (y[b] == l)).sum())
        top2 = logits.topk(2, dim=1).values
        out[f"shape{i}"] = np.array([n, c, h, w])
        out[f"loss{i}"] = np.array(float(loss))
        out[f"argmax{i}"] = pred.numpy().astype(np.uint8)
        out[f"margin_min{i}"] = np.array(float((top2[:, 0] - top2[:, 1]).min()))
        out[f"conf{i}"] = conf
        lf = logits.reshape(-1)
        out[f"logits_sub{i}"] = lf[::13].numpy().copy() if lf.numel() > 50000 else lf.numpy().copy()
        out[f"logits_absmax{i}"] = np.array(float(logits.abs().max()))
    np.savez_compressed(OUT / (tag + "_eval.npz"), **out)
    print(tag, {k: v.shape for k, v in out.items() if k.startswith("logits_sub")})


def main():
    (OUT / "ckpt").mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)  # fixed accumulation order

    # ---- PB_FCN checkpoints (legacy head name -> segmenter) ------------------------------
    for name, no_scale, shapes in [
        ("bestModelSeg", False, [(2, 3, 24, 32), (1, 3, 120, 160)]),
        ("bestModelSegFinetunedPruned", False, [(2, 3, 24, 32), (4, 3, 120, 160)]),
        ("bestModelSegVGA", True, [(1, 3, 48, 64), (1, 3, 480, 640)]),
    ]:
        sd = load_pth(name + ".pth")
        save_ckpt(name, sd)
        m = REFM.PB_FCN(32, 5, 1, no_scale, 0)
        remapped = {("segmenter." + k[len("classifier."):] if k.startswith("classifier.classifier.") else k): v
                    for k, v in sd.items()}
        res = m.load_state_dict(remapped, strict=False)
        assert not res.unexpected_keys, res
        eval_golden(name, m, lambda mm, x: mm(x), shapes)

    # ---- LabelProp ------------------------------------------------------------------------
    sd = load_pth("bestModelLPFinetunedPruned.pth")
    save_ckpt("bestModelLPFinetunedPruned", sd)
    m = ref_labelprop()
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith("num_batches_tracked") for k in res.missing_keys), res
    eval_golden("bestModelLPFinetunedPruned", m, labelprop_ref_forward, [(2, 8, 24, 32), (2, 8, 120, 160)],
                weights=synth.LP_CLASS_WEIGHTS)
    # paramSave.py:5-17 wire format: float64 flatten of the state dict in key order
    dat = np.fromfile(REF / "weightsLP" / "weights.dat", dtype=np.float64)
    np.savez_compressed(OUT / "weightsLP_head.npz", n=np.array(dat.size), head=dat[:64], tail=dat[-64:],
                        sha256=np.frombuffer(hashlib.sha256(dat.tobytes()).digest(), dtype=np.uint8))

    # ---- ROBO_UNet variants, random init (train.py:332-337 seeds) -------------------------
    for tag, kw in [("robo_default", {}), ("robo_unet_pool", dict(pool=True, levels=3, bellySize=0)),
                    ("robo_noscale", dict(noScale=True))]:
        torch.manual_seed(12345678)
        m = REFM.ROBO_UNet(**kw)
        # make eval-mode BN non-trivial: a few training forwards move the running stats
        m.train()
        with torch.no_grad():
            for s in range(3):
                m(synth.images(4, 3, 48, 64, seed=77 + s))
        # (not saved as a checkpoint: tests rebuild it from the seed with the same three forwards)
        full = (1, 3, 240, 320) if kw.get("noScale") else (1, 3, 120, 160)
        eval_golden(tag, m, lambda mm, x: mm(x), [(2, 3, 48, 64), full])

    # ---- 3 reference training steps ---------------------------------------------------------
    torch.manual_seed(12345678)
    m = REFM.ROBO_UNet()
    crit = REFM.CrossEntropyLoss2d(torch.tensor(synth.CLASS_WEIGHTS))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    m.train()
    losses, regs, corrects, gnorms = [], [], [], []
    for s in range(3):
        x = synth.images(8, 3, 48, 64, seed=100 + s)
        y = synth.labels_learnable(x)
        opt.zero_grad()
        pred = m(x)
        loss = crit(pred, y)
        reg = 1e-6 * sum(p.abs().sum() for p in m.parameters())  # train.py:23-27, 52-55
        loss = loss + reg
        loss.backward()
        gnorms.append([float(p.grad.norm()) for p in m.parameters()])
        opt.step()
        losses.append(float(loss)); regs.append(float(reg))
        corrects.append(int((pred.argmax(1) == y).sum()))
    np.savez_compressed(OUT / "robo_train.npz", losses=np.array(losses), regs=np.array(regs),
                        corrects=np.array(corrects), gnorms=np.array(gnorms),
                        final_w0=m.state_dict()["downPart.Level0.layers.Conv0.conv.weight"].numpy(),
                        final_rv=m.state_dict()["PB.PB_1.layers.Conv1.bn.running_var"].numpy())
    print("losses", losses)


class RefChannelPBFCN(torch.nn.Module):
    """The legacy 16-plane PB_FCN that pth/bestModelSegFinetunedPruned_bu.pth belongs to, assembled from the
    REFERENCE's own blocks (model.py:126-199, 256-267) with PB_FCN's forward (model.py:291-309, noScale=False).  No
    class of the current model.py has these widths; the checkpoint strict-loads into this one (SURVEY 8c vi)."""

    def __init__(self, enc, ups, num_classes):
        super().__init__()
        f = torch.nn.Module()
        f.conv0 = REFM.ConvPoolSimple(3, enc[0], 3, 1, 2, 2, False)
        f.conv1 = REFM.ConvPoolSimple(enc[0], enc[1], 3, 2, 1, 1, False)
        f.conv2 = REFM.ConvPool(enc[1], enc[2])
        f.conv3 = REFM.ConvPool(enc[2], enc[3])
        for i in range(4, 9):
            setattr(f, "conv%d" % i, REFM.ConvPoolSimple(enc[i - 1], enc[i], 3, 1, 2, 2, False))
        self.FCN = f
        self.up1 = REFM.upSampleTransposeConv(enc[8], ups[0])
        self.up2 = REFM.upSampleTransposeConv(ups[0], ups[1])
        self.up3 = REFM.upSampleTransposeConv(ups[1], ups[2])
        self.classifier = REFM.Classifier(ups[2], num_classes)

    def forward(self, x):
        f = self.FCN
        x0 = f.conv0(x)
        x1 = f.conv1(x0)
        x2 = f.conv2(x1)
        x3 = f.conv8(f.conv7(f.conv6(f.conv5(f.conv4(f.conv3(x2))))))
        x = self.up1(x3) + x2
        x = self.up2(x) + x1
        x = self.up3(x) + x0
        return self.classifier(x)


def channel_pruned():
    """`python -m oracle.make_golden bu`: the irregular-channel checkpoint (BASELINE configs[2])."""
    torch.set_num_threads(1)
    name = "bestModelSegFinetunedPruned_bu"
    sd = load_pth(name + ".pth")
    save_ckpt(name, sd)
    enc = [sd[f"FCN.conv{i}.{'pool' if i in (2, 3) else 'conv'}.weight"].shape[0] for i in range(9)]
    ups = [sd[f"up{i}.conv.weight"].shape[1] for i in (1, 2, 3)]
    m = RefChannelPBFCN(enc, ups, sd["classifier.classifier.weight"].shape[0])
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith("num_batches_tracked") for k in res.missing_keys), res
    eval_golden(name, m, lambda mm, x: mm(x), [(2, 3, 24, 32), (4, 3, 120, 160)])
    print(name, "enc", enc, "ups", ups)


def fcn_checkpoint():
    """`python -m oracle.make_golden fcn`: the released FCN checkpoint (model.py:311-331, pth/bestModelSeg1.pth)."""
    torch.set_num_threads(1)
    name = "bestModelSeg1"
    sd = load_pth(name + ".pth")
    save_ckpt(name, sd)
    m = REFM.FCN()
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith("num_batches_tracked") for k in res.missing_keys), res
    eval_golden(name, m, lambda mm, x: mm(x), [(2, 3, 24, 32), (2, 3, 120, 160)])


def loss_curves():
    """`python -m oracle.make_golden curve`: 200 REFERENCE training steps (train.py:43-74: zero_grad -> model ->
    CrossEntropyLoss2d -> + decay * l1reg -> backward -> Adam.step) of the reference's own ROBO_UNet, seed 12345678
    (train.py:332), Adam 1e-3, L1 1e-6, class weights train.py:309, a fresh synthetic batch with learnable labels
    every step.  Two sizes: 8x3x48x64 (also re-run by the CPU oracle test) and 8x3x120x160 (the real frame size)."""
    torch.set_num_threads(1)
    out = {}
    for tag, (b, h, w) in (("small", (8, 48, 64)), ("full", (8, 120, 160))):
        torch.manual_seed(12345678)
        m = REFM.ROBO_UNet()
        crit = REFM.CrossEntropyLoss2d(torch.tensor(synth.CLASS_WEIGHTS))
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        m.train()
        losses, corrects = [], []
        for s in range(200):
            x = synth.images(b, 3, h, w, seed=5000 + s)
            y = synth.labels_learnable(x)
            opt.zero_grad()
            pred = m(x)
            loss = crit(pred, y) + 1e-6 * sum(p.abs().sum() for p in m.parameters())
            loss.backward()
            opt.step()
            losses.append(float(loss))
            corrects.append(int((pred.argmax(1) == y).sum()))
        out[f"losses_{tag}"] = np.array(losses)
        out[f"corrects_{tag}"] = np.array(corrects)
        out[f"shape_{tag}"] = np.array([b, 3, h, w])
        print(tag, "first", losses[:3], "last", losses[-3:])
    np.savez_compressed(OUT / "robo_curve200.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bu":
        channel_pruned()
    elif len(sys.argv) > 1 and sys.argv[1] == "curve":
        loss_curves()
    elif len(sys.argv) > 1 and sys.argv[1] == "fcn":
        fcn_checkpoint()
    else:
        main()
        channel_pruned()
        fcn_checkpoint()
        loss_curves()
