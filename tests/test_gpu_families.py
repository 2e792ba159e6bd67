"""GPU parity of the remaining model families (`--v2` cat skips + 3x3 head, FCN, the channel-pruned PB_FCN checkpoint)
and of the conv geometries only they use.  Same gates as tests/test_gpu_models.py."""
import pytest
import torch

import synth
from oracle import ref_model as R
from nets import pb_fcn_state
from test_gpu_models import LOGIT_TOL, _check_eval, _grad_check
from util import load_ckpt, load_golden, with_nbt
from util import assert_close

pytestmark = pytest.mark.gpu

# SURVEY.md section 8(f) N4 rows ("FCN / --v2 variants: cat-skip, 3x3 head") and BASELINE configs[2]
# "irregular channel counts"; first seen green on a B200 in round 2 (gpurun_out/ab_queue.log).


def _seeded_state(model, oracle_fwd_train, cin=3):
    """Seeded init + three oracle training forwards (running statistics away from the identity)."""
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for s in range(3):
            oracle_fwd_train(sd, synth.images(4, cin, 48, 64, seed=77 + s))
    return sd


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
                       wpacked=ops.conv_pack(g, wc, ops.PACK_FWD, nhw=(3, 12, 20)) if tc_f else None)
    assert_close(f"fwd {geom} {cin}->{cout}", got, y.detach(), 8e-6)
    dx = ops.conv_dgrad(g, dy.cuda(), wc, (12, 20), math=ops.MATH_AUTO,
                        wpacked=ops.conv_pack(g, wc, ops.PACK_DGRAD, nhw=(3, 12, 20)) if tc_d else None)
    assert_close(f"dgrad {geom} {cin}->{cout}", dx, x.grad, 1.2e-5)
    dw, db = ops.conv_wgrad(g, x.detach().cuda(), dy.cuda(), want_bias=True, math=ops.MATH_AUTO)
    assert_close(f"wgrad {geom} {cin}->{cout}", dw, w.grad, 1e-5)
    assert_close(f"bgrad {geom}", db, b.grad, 1e-5)
