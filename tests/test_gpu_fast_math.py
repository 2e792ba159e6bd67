"""The FAST math modes (rcv_math RCV_MATH_TF32 / RCV_MATH_BF16), stated separately from the fp32-parity mode
(north_star: "logits within 1e-4 relative in fp32 (bf16 variant stated separately)").

What the modes are, and what is asserted here:
  * kernel level: a fast-mode conv equals the fp32 conv of operands ROUNDED to the mode's operand format (tf32: 10
    mantissa bits, round-to-nearest-even on the low 13 bits as the split helper does; bf16: 8 mantissa bits) to
    accumulation-order accuracy (1e-5 of the output range): the only difference from the parity mode is that rounding;
  * net level, released checkpoints (BN gains up to 50x): logit error and argmax-flip rate against the fp32 oracle,
    with the tolerance of each mode written below; confusion counts are compared on the label map the mode produced;
  * training: gradients of one step within the mode's tolerance of CPU autograd, and the 200-step loss curve within 5 %.
The CUDA-core layers (<= 16 output channels) stay exact fp32 in every mode."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
from nets import pb_fcn_state, robo_state
from oracle import ref_model as R
from oracle.ref_train import OracleTrainer
from util import assert_close

pytestmark = pytest.mark.gpu

# mode -> (logit tolerance relative to max|ref|, allowed argmax flip fraction, gradient tolerance)
TOL = {"tf32": (2e-3, 1e-3, 8e-2), "bf16": (6e-3, 2e-3, 1.5e-1)}


def _round_tf32(t):
    """fp32 -> tf32 as the staging code rounds it (add half an ulp of the 13 dropped bits, truncate)."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


ROUND = {"tf32": _round_tf32, "bf16": _round_bf16}


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("cin,cout,dil,nhw", [(128, 128, 1, (4, 15, 20)), (64, 128, 2, (3, 15, 20)), (128, 64, 1, (2, 9, 7)),
                                              (32, 64, 1, (2, 30, 40)), (64, 64, 1, (2, 12, 16))])
def test_fast_conv_is_the_conv_of_rounded_operands(mode, cin, cout, dil, nhw):
    from robocupvision_b200 import ops
    math = {"tf32": ops.MATH_TF32, "bf16": ops.MATH_BF16}[mode]
    n, h, w_ = nhw
    gen = torch.Generator().manual_seed(cin + cout + dil)
    x = torch.randn(n, cin, h, w_, generator=gen)
    wt = torch.randn(cout, cin, 3, 3, generator=gen) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=gen)
    g = ops.ConvGeom(cin, cout, 3, 1, dil, dil, False)
    assert ops.conv_engine(g, n, h, w_, ops.PACK_FWD, math) == ops.ENGINE_UMMA
    wp = ops.conv_pack(g, wt.cuda(), ops.PACK_FWD, math=math, nhw=(n, h, w_))
    got = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), epilogue=ops.EPI_RELU, math=math, wpacked=wp)
    # the bf16 operand format applies where the reduced channel count is a multiple of 64 (halo-staged kernel);
    # other tensor-core layers of that mode run single-pass tf32
    rnd = ROUND["bf16" if (mode == "bf16" and cin % 64 == 0) else "tf32"]
    ref = F.relu(F.conv2d(rnd(x).double(), rnd(wt).double(), b.double(), 1, dil, dil)).float()
    assert_close(f"{mode} conv {cin}->{cout} d{dil} vs rounded operands", got, ref, 1e-5)
    exact = F.relu(F.conv2d(x, wt, b, 1, dil, dil))
    err = float((got.cpu() - exact).abs().max()) / float(exact.abs().max())
    print(f"{mode} conv {cin}->{cout}: {err:.2e} of the output range away from fp32")
    assert err <= {"tf32": 2e-3, "bf16": 8e-3}[mode]

    # input gradient and weight gradient in the same mode
    dy = torch.randn(n, cout, h, w_, generator=gen)
    wpd = ops.conv_pack(g, wt.cuda(), ops.PACK_DGRAD, math=math, nhw=(n, h, w_))
    dx = ops.conv_dgrad(g, dy.cuda(), wt.cuda(), (h, w_), math=math, wpacked=wpd)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, 1, dil, dil)
    dw, _ = ops.conv_wgrad(g, x.cuda(), dy.cuda(), math=math)
    dw_ref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, 1, dil, dil)
    gt = {"tf32": 2e-3, "bf16": 8e-3}[mode]
    assert_close(f"{mode} dgrad", dx, dx_ref, gt)
    assert_close(f"{mode} wgrad", dw, dw_ref, gt)


def _eval_fast(tag, mode, model, oracle_fwd, shapes, weights=synth.CLASS_WEIGHTS):
    from robocupvision_b200.train import EvalStep
    ltol, ftol, _ = TOL[mode]
    model.cuda().eval()
    par = EvalStep(model, weights)                     # parity mode first: its own logits are the near reference
    outs_par = []
    xs = []
    for i, (n, c, h, w) in enumerate(shapes):
        x = synth.images(n, c, h, w, seed=1234 + i)
        y = synth.labels_random(n, h, w, 5, seed=4321 + i)
        xs.append((x, y))
        o = par(x.cuda(), y.cuda())
        outs_par.append({k: v.clone() for k, v in o.items()})
    model.set_math(mode)
    fast = EvalStep(model, weights)
    for (x, y), op in zip(xs, outs_par):
        o = fast(x.cuda(), y.cuda())
        with torch.no_grad():
            ref = oracle_fwd(x)
        logits = o["logits"].cpu()
        scale = float(ref.abs().max())
        err = float((logits - ref).abs().max()) / scale
        perr = float((op["logits"].cpu() - ref).abs().max()) / scale
        am, am_ref = o["argmax"].cpu(), ref.argmax(1)
        flips = float((am != am_ref).float().mean())
        top2 = ref.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]) / scale
        # a flipped pixel is one whose fp32 top-2 margin is inside the mode's logit error
        worst_margin = float(margin[am != am_ref].max()) if bool((am != am_ref).any()) else 0.0
        lref = float(R.cross_entropy_2d(ref, y, torch.tensor(weights)))
        lrel = abs(float(o["loss"]) - lref) / max(1.0, abs(lref))
        print(f"{tag} {mode} {tuple(x.shape)}: logits {err:.2e} of range (parity mode {perr:.1e}), argmax flips "
              f"{flips * 100:.3f} % (largest flipped margin {worst_margin:.2e}), loss rel {lrel:.2e}")
        assert perr <= 1e-4
        assert err <= ltol, f"{tag} {mode}: logit error {err:.2e} > {ltol:.1e}"
        assert flips <= ftol, f"{tag} {mode}: {flips * 100:.3f} % of label-map pixels flipped"
        assert worst_margin <= 2 * ltol, f"{tag} {mode}: a pixel with margin {worst_margin:.2e} flipped"
        assert lrel <= ltol
        assert int(o["correct"]) == int((am == y).sum())
        assert int(o["conf"].sum()) == y.numel()
    model.set_math("parity")


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("name,no_scale,shapes", [
    ("bestModelSeg", False, [(8, 3, 120, 160), (1, 3, 120, 160)]),
    ("bestModelSegFinetunedPruned", False, [(8, 3, 120, 160)]),
    ("bestModelSegVGA", True, [(1, 3, 480, 640)])])
def test_released_checkpoints_fast_modes(mode, name, no_scale, shapes):
    from robocupvision_b200.model import PB_FCN, load_legacy_state_dict
    osd, raw = pb_fcn_state(name)
    m = PB_FCN(32, 5, 1, no_scale, 0)
    load_legacy_state_dict(m, raw)
    _eval_fast(name, mode, m, lambda x: R.pb_fcn_forward(osd, x, no_scale), shapes)


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_robo_unet_fast_modes_eval(mode):
    from robocupvision_b200.model import ROBO_UNet
    sd, kw, okw = robo_state("robo_default")
    m = ROBO_UNet(**kw)
    m.load_state_dict(sd)
    _eval_fast("robo_default", mode, m, lambda x: R.robo_unet_forward(sd, x, **okw), [(8, 3, 120, 160)])


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_robo_unet_fast_modes_gradients(mode):
    """One training step's gradients in a fast mode against CPU autograd over the oracle: the whole gradient in relative
    L2 within the mode's tolerance, every tensor within four times that (the operand rounding flips ReLU decisions
    on pre-activations within ~1e-3 of zero, so the deep encoder's small gradients carry the largest share)."""
    from robocupvision_b200.model import CrossEntropyLoss2d, ROBO_UNet
    sd, kw, okw = robo_state("robo_default")
    m = ROBO_UNet(**kw)
    m.load_state_dict(sd)
    x = synth.images(4, 3, 48, 64, seed=5)
    y = synth.labels_learnable(x)
    osd = R.leaf_state_dict(sd)
    pred_ref = R.robo_unet_forward(osd, x, training=True, **okw)
    loss_ref = R.cross_entropy_2d(pred_ref, y, torch.tensor(synth.CLASS_WEIGHTS))
    loss_ref.backward()
    m.cuda().train()
    m.set_math(mode)
    crit = CrossEntropyLoss2d(torch.tensor(synth.CLASS_WEIGHTS)).cuda()
    pred = m(x.cuda())
    loss = crit(pred, y.cuda())
    loss.backward()
    ltol, _, gtol = TOL[mode]
    assert_close(f"{mode} train logits", pred, pred_ref, ltol)
    assert abs(float(loss.detach()) - float(loss_ref)) <= ltol * max(1.0, abs(float(loss_ref)))
    gmax = max(float(v.grad.norm()) for v in osd.values() if v.grad is not None)
    worst, tot = (0.0, ""), [0.0, 0.0]
    for k, p in m.named_parameters():
        gref = osd[k].grad
        if gref is None:
            continue
        e = float((p.grad.cpu() - gref).norm())
        tot = [tot[0] + e * e, tot[1] + float(gref.norm()) ** 2]
        err = e / max(float(gref.norm()), 1e-3 * gmax)
        worst = max(worst, (err, k))
        assert err <= 4 * gtol, f"{mode}: grad {k} relative L2 {err:.3e}"
    whole = (tot[0] / tot[1]) ** 0.5
    print(f"{mode}: gradient relative L2: whole {whole:.2e}, worst tensor {worst[0]:.2e} ({worst[1]})")
    assert whole <= gtol, f"{mode}: whole gradient relative L2 {whole:.3e}"


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_fast_modes_loss_curve(mode):
    """The 200-step protocol of tests/test_gpu_train.py::test_loss_curve_200_steps (train.py:43-74: Adam 1e-3, L1 1e-6,
    learnable labels) in a fast mode against the curve the fp32 REFERENCE produced (tests/golden/robo_curve200.npz,
    small size): step 1 within the mode's logit tolerance, every step within 5 %, mean of the last 10 within 3 %."""
    from robocupvision_b200.model import ROBO_UNet
    from robocupvision_b200.train import TrainStep
    from util import load_golden
    gold = load_golden("robo_curve200")
    ref = np.asarray(gold["losses_small"], dtype=np.float64)
    b, c, h, w = (int(v) for v in gold["shape_small"])
    torch.manual_seed(12345678)
    m = ROBO_UNet().cuda()
    m.set_math(mode)
    ts = TrainStep(m, synth.CLASS_WEIGHTS, lr=1e-3, l1_decay=1e-6, use_graph=True)
    got = []
    for s in range(200):
        x = synth.images(b, c, h, w, seed=5000 + s)
        y = synth.labels_learnable(x)
        ts.step(x.cuda(), y.cuda())
        got.append(ts.loss_value())
    got = np.asarray(got)
    rel = np.abs(got - ref) / np.abs(ref)
    print(f"{mode}: loss-curve step-1 rel {rel[0]:.2e}, max rel {rel.max():.2e} at step {int(rel.argmax())}, "
          f"final {got[-10:].mean():.5f} vs reference {ref[-10:].mean():.5f}")
    assert rel[0] <= TOL[mode][0]
    assert rel.max() <= 5e-2
    assert abs(got[-10:].mean() - ref[-10:].mean()) <= 3e-2 * ref[-10:].mean()
