"""GPU parity tests whose first run is still pending (they sort after every other GPU test file on purpose: a fault in
an unproven path cannot disturb the verified suite).  Same gates as tests/test_gpu_models.py."""
import pytest
import torch

import synth
from oracle import ref_model as R
from nets import pb_fcn_state
from test_gpu_models import LOGIT_TOL, _check_eval, _grad_check
from util import load_ckpt, load_golden, with_nbt
from util import assert_close

pytestmark = pytest.mark.gpu

# The two model families below are SURVEY.md section 8(f) N4 rows ("FCN / --v2 variants: cat-skip, 3x3 head").  The
# tests were written after this round's GPU budget was spent: their first run is the driver's, hence non-strict
# xfail (a pass shows as XPASS, a failure does not hide the verified suite).  Drop the marker once seen green.
_first_run = pytest.mark.xfail(strict=False, reason="first GPU run pending (added after the round's GPU budget)")


def _seeded_state(model, oracle_fwd_train, cin=3):
    """Seeded init + three oracle training forwards (running statistics away from the identity)."""
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for s in range(3):
            oracle_fwd_train(sd, synth.images(4, cin, 48, 64, seed=77 + s))
    return sd


@_first_run
def test_robo_unet_v2_cat_skips_eval_and_backward():
    """`--v2` (model.py:462-511 with v2=True): decoder concatenates the skip tensors, 3x3 head on 16 channels."""
    from robocupvision_b200.model import ROBO_UNet
    kw = dict(v2=True, classSize=3)
    okw = dict(v2=True, class_size=3)
    torch.manual_seed(12345678)
    m = ROBO_UNet(**kw)
    sd = _seeded_state(m, lambda s, xx: R.robo_unet_forward(s, xx, training=True, **okw))
    m.load_state_dict(sd)
    m.cuda().eval()
    x = synth.images(3, 3, 120, 160, seed=31)
    with torch.no_grad():
        assert_close("v2 eval logits", m(x.cuda()), R.robo_unet_forward(sd, x, **okw), LOGIT_TOL)
    xb = synth.images(4, 3, 48, 64, seed=5)
    _grad_check("robo_v2", m, lambda s, xx: R.robo_unet_forward(s, xx, training=True, **okw), sd, xb,
                synth.labels_learnable(xb), synth.CLASS_WEIGHTS)


@_first_run
def test_fcn_eval_and_backward():
    """`FCN` (model.py:311-331): DownSamplerThick encoder (ConvPoolDouble blocks), three up blocks, 1x1 head."""
    from robocupvision_b200.model import FCN
    torch.manual_seed(12345678)
    m = FCN()
    sd = _seeded_state(m, lambda s, xx: R.fcn_forward(s, xx, training=True))
    m.load_state_dict(sd)
    m.cuda().eval()
    x = synth.images(3, 3, 120, 160, seed=32)
    with torch.no_grad():
        assert_close("FCN eval logits", m(x.cuda()), R.fcn_forward(sd, x), LOGIT_TOL)
    xb = synth.images(4, 3, 48, 64, seed=6)
    _grad_check("fcn", m, lambda s, xx: R.fcn_forward(s, xx, training=True), sd, xb, synth.labels_learnable(xb),
                synth.CLASS_WEIGHTS)


@_first_run
def test_pb_fcn_channel_pruned_checkpoint():
    """BASELINE configs[2] "irregular channel counts": pth/bestModelSegFinetunedPruned_bu.pth (encoder
    16-16-16-32-64-64-128-64-32, decoder 16-16-16; 70.8 % zero weights) through PB_FCN_Channels, against the golden
    outputs of the reference's own blocks (oracle/make_golden.py bu) and the oracle."""
    from robocupvision_b200.model import PB_FCN_Channels
    name = "bestModelSegFinetunedPruned_bu"
    osd, raw = pb_fcn_state(name)
    m = PB_FCN_Channels.from_state_dict(raw)
    m.cuda().eval()
    _check_eval(name, m, lambda x: R.pb_fcn_forward(osd, x, False), load_golden(name + "_eval"))


@_first_run
def test_fcn_released_checkpoint():
    """pth/bestModelSeg1.pth through the drop-in FCN on the GPU, against the reference's golden outputs."""
    from robocupvision_b200.model import FCN, load_legacy_state_dict
    raw = load_ckpt("bestModelSeg1")
    osd = with_nbt(raw)
    m = FCN()
    load_legacy_state_dict(m, raw)
    m.cuda().eval()
    _check_eval("bestModelSeg1", m, lambda x: R.fcn_forward(osd, x), load_golden("bestModelSeg1_eval"))


# Per-kernel parity for the channel combinations the three families above add to the verified matrix (same checks as
# tests/test_gpu_ops.py::test_conv_fwd / test_conv_dgrad_wgrad, automatic engine choice).
@_first_run
@pytest.mark.parametrize("geom,cin,cout", [("k3s1d2", 3, 16), ("k3s1d2", 16, 32), ("k3s2", 16, 16), ("k3s2", 32, 32),
                                           ("convT", 64, 16), ("convT", 32, 8), ("convT", 32, 16), ("k3s1d1", 16, 5),
                                           ("k3s1d2", 64, 32), ("k3s1d2", 32, 64)])
def test_conv_kernels_for_the_new_families(geom, cin, cout):
    import torch.nn.functional as F  # noqa: F401
    from robocupvision_b200 import ops
    from test_gpu_ops import _mk, _ref_conv
    g, x, w, b = _mk(geom, cin, cout, 3, 12, 20, seed=5)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = _ref_conv(geom, x, w, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
    y.backward(dy)
    wc = w.detach().cuda()
    tc_f, tc_d = (ops.conv_uses_tensor_cores(g, d, ops.MATH_AUTO) for d in (ops.PACK_FWD, ops.PACK_DGRAD))
    got = ops.conv_fwd(g, x.detach().cuda(), wc, b.detach().cuda(), math=ops.MATH_AUTO,
                       wpacked=ops.conv_pack(g, wc, ops.PACK_FWD) if tc_f else None)
    assert_close(f"fwd {geom} {cin}->{cout}", got, y.detach(), 8e-6)
    dx = ops.conv_dgrad(g, dy.cuda(), wc, (12, 20), math=ops.MATH_AUTO,
                        wpacked=ops.conv_pack(g, wc, ops.PACK_DGRAD) if tc_d else None)
    assert_close(f"dgrad {geom} {cin}->{cout}", dx, x.grad, 1.2e-5)
    dw, db = ops.conv_wgrad(g, x.detach().cuda(), dy.cuda(), want_bias=True, math=ops.MATH_AUTO)
    assert_close(f"wgrad {geom} {cin}->{cout}", dw, w.grad, 1e-5)
    assert_close(f"bgrad {geom}", db, b.grad, 1e-5)


# Cooperative single-launch BatchNorm backward (csrc/rcv_bn_fused.cu): never run on a GPU yet, and a grid-wide barrier
# is the kind of code whose first run belongs in an interactive session, not in an unattended suite -- opt in with
# RCV_TEST_EXPERIMENTAL=1 (tools/ab_queue.sh does).
import os  # noqa: E402

_experimental = pytest.mark.skipif(os.environ.get("RCV_TEST_EXPERIMENTAL", "0") == "0",
                                   reason="experimental kernel: set RCV_TEST_EXPERIMENTAL=1")


@_experimental
@pytest.mark.parametrize("order", ["relu_affine", "affine_relu"])
@pytest.mark.parametrize("shape", [(64, 128, 15, 20), (64, 64, 15, 20), (64, 32, 30, 40), (4, 8, 12, 20), (3, 128, 5, 4)])
def test_bn_bwd_fused_matches_two_pass(order, shape):
    from robocupvision_b200 import _lib, ops
    n, c, h, w = shape
    if not _lib.load().rcv_bn_bwd_fused_supported(n, c, h * w):
        pytest.skip("tensor does not fit one co-resident grid")
    gen = torch.Generator().manual_seed(3)
    z = torch.randn(shape, generator=gen).cuda()
    if order == "relu_affine":
        z = torch.relu(z)
    dy = torch.randn(shape, generator=gen).cuda()
    gamma, beta = torch.randn(c, generator=gen).cuda(), torch.randn(c, generator=gen).cuda()
    stats = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    d = z.double()
    stats[:c], stats[c:] = d.sum((0, 2, 3)), (d * d).sum((0, 2, 3))
    scale, shift, mean, invstd = ops.bn_finalize(stats, n * h * w, gamma, beta, None, None, 0.1, 1e-5)
    code = ops.EPI_RELU_AFFINE if order == "relu_affine" else ops.EPI_AFFINE_RELU
    old = ops.BN_BWD_FUSED
    try:
        ops.BN_BWD_FUSED = False
        ref = ops.bn_bwd(code, dy, z, scale, shift, mean, invstd, want_dbias=True)
        ops.BN_BWD_FUSED = True
        got = ops.bn_bwd(code, dy, z, scale, shift, mean, invstd, want_dbias=True)
    finally:
        ops.BN_BWD_FUSED = old
    torch.cuda.synchronize()
    for name, a, b in zip(("dconv", "dgamma", "dbeta", "dbias"), got, ref):
        assert_close(f"bn_bwd_fused {name} {order} {shape}", a, b.cpu(), 2e-6, atol=1e-6)
