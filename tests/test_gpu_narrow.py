"""Parity of the narrow-layer engine (csrc/rcv_narrow.cu: TMA halo staging + FFMA2 direct conv,
<= 16 output channels) against the ATen CPU op, through the C ABI.  Every case first checks that the
dispatcher really picks that engine (rcv_conv_engine), so a silent detour through another kernel
cannot pass.  Geometries are the ones model.py uses (model.py:112,130-133,170,186-187,259,411,554)."""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_ops import GEOMS, _mk, _ref_conv
from util import assert_close

pytestmark = pytest.mark.gpu

FWD_CASES = [
    # geom, cin, cout, n, h, w
    ("k3s1d1", 3, 8, 2, 12, 20),
    ("k3s1d1", 8, 8, 2, 7, 8),       # odd height: slot-1 row of the last thread row is outside
    ("k3s1d1", 16, 16, 2, 30, 40),   # several channel chunks, double-buffered
    ("k3s1d1", 32, 16, 1, 12, 20),
    ("k3s1d1", 5, 7, 3, 9, 12),      # channel counts that are not multiples of anything
    ("k3s1d1", 3, 8, 1, 120, 160),   # ROBO_UNet Level0 at full size
    ("k3s1d1", 3, 8, 1, 5, 640),     # three column tiles (VGA width)
    ("k3s1d1", 16, 16, 1, 6, 320),
    ("k3s1d2", 3, 8, 2, 12, 20),     # PB_FCN conv0 (dilated)
    ("k3s1d2", 16, 16, 2, 13, 24),
    ("k3s2", 8, 16, 2, 24, 40),      # stride 2: grid stride 2, 9-wide register window
    ("k3s2", 3, 8, 2, 26, 24),       # odd output height (13)
    ("k3s2", 16, 16, 1, 120, 160),   # LabelProp down2 shape
    ("convT", 16, 8, 2, 12, 20),     # four parity classes, two per thread
    ("convT", 32, 16, 2, 15, 20),
    ("convT", 16, 16, 1, 60, 80),
    ("k1", 8, 5, 2, 12, 20),         # class head
    ("k1", 16, 5, 2, 120, 160),
]


def _engine(g, n, h, w, direction=0, math=2):
    """The engine rcv_conv_engine names for the layer GIVEN a packed panel.  These tests call the ops without a panel,
    which keeps a <= 16-channel layer on the narrow-layer kernels they are about; with a panel, the 16 -> <= 16
    stride-1 3x3 layers go to the persistent tensor-core kernel instead (tests/test_gpu_ops.py::test_conv16_*), so
    for those the query's answer is mapped back."""
    from robocupvision_b200 import _lib, ops
    e = ops.conv_engine(g, n, h, w, direction, math)
    reduced = g.cout if direction == 1 else g.cin
    if (e == _lib.ENGINE_UMMA and direction < 2 and reduced == 16 and g.k == 3 and g.stride == 1 and g.dil == 1
            and not g.transposed and max(g.cin, g.cout) <= 16):
        return _lib.ENGINE_NARROW
    return e


@pytest.mark.parametrize("math", [0, 2])
@pytest.mark.parametrize("geom,cin,cout,n,h,w", FWD_CASES)
def test_narrow_fwd(geom, cin, cout, n, h, w, math):
    from robocupvision_b200 import _lib, ops
    g, x, wt, b = _mk(geom, cin, cout, n, h, w, seed=31)
    assert _engine(g, n, h, w, 0, math) == _lib.ENGINE_NARROW
    ref = _ref_conv(geom, x, wt, b)
    got = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), math=math)
    assert_close(f"narrow fwd {geom} {cin}->{cout} {n}x{h}x{w}", got, ref, 2e-6)


@pytest.mark.parametrize("epi", ["none", "relu", "relu_affine", "affine_relu", "affine"])
@pytest.mark.parametrize("geom,cin,cout,h,w", [("k3s1d1", 8, 16, 11, 16), ("convT", 32, 16, 9, 12), ("k3s2", 8, 16, 14, 24),
                                              ("k1", 8, 5, 10, 16), ("k3s1d2", 3, 8, 12, 16)])
def test_narrow_epilogues_residual_stats(epi, geom, cin, cout, h, w):
    from robocupvision_b200 import _lib, ops
    g, x, wt, b = _mk(geom, cin, cout, 3, h, w, seed=2)
    assert _engine(g, 3, h, w) == _lib.ENGINE_NARROW
    gen = torch.Generator().manual_seed(9)
    sc, sh = torch.randn(cout, generator=gen), torch.randn(cout, generator=gen)
    v = _ref_conv(geom, x, wt, b)
    res = torch.randn(v.shape, generator=gen)
    A, B = sc.view(1, -1, 1, 1), sh.view(1, -1, 1, 1)
    ref = {"none": v, "relu": F.relu(v), "relu_affine": A * F.relu(v) + B, "affine_relu": F.relu(A * v + B),
           "affine": A * v + B}[epi] + res
    code = {"none": ops.EPI_NONE, "relu": ops.EPI_RELU, "relu_affine": ops.EPI_RELU_AFFINE,
            "affine_relu": ops.EPI_AFFINE_RELU, "affine": ops.EPI_AFFINE}[epi]
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    got = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), epilogue=code, scale=sc.cuda(), shift=sh.cuda(),
                       residual=res.cuda(), stats=stats, math=ops.MATH_AUTO)
    assert_close(f"narrow epilogue {epi} {geom}", got, ref, 3e-6)
    rd = ref.double()
    assert_close("stats sum", stats[:cout], rd.sum((0, 2, 3)), 1e-6, atol=1e-3)
    assert_close("stats sumsq", stats[cout:], (rd * rd).sum((0, 2, 3)), 1e-6, atol=1e-3)


DGRAD_CASES = [
    # the gradient w.r.t. the input of a cin->cout layer reduces over cout and produces cin (<= 16) channels
    ("k3s1d1", 16, 16, 2, 12, 20),
    ("k3s1d1", 8, 64, 2, 9, 12),     # long reduction (64 channels in 8 chunks)
    ("k3s1d2", 8, 16, 2, 12, 20),
    ("k3s2", 8, 16, 2, 24, 40),      # stride-2 conv: parity classes over the coarse grid
    ("k3s2", 16, 32, 2, 12, 24),
    ("convT", 16, 8, 2, 12, 20),     # transposed conv: a stride-2 conv over dy
    ("convT", 16, 16, 1, 30, 40),
    ("k1", 8, 5, 2, 12, 20),
]


@pytest.mark.parametrize("geom,cin,cout,n,h,w", DGRAD_CASES)
def test_narrow_dgrad(geom, cin, cout, n, h, w):
    from robocupvision_b200 import _lib, ops
    g, x, wt, b = _mk(geom, cin, cout, n, h, w, seed=5)
    assert _engine(g, n, h, w, 1) == _lib.ENGINE_NARROW
    x.requires_grad_(True)
    y = _ref_conv(geom, x, wt, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
    y.backward(dy)
    dx = ops.conv_dgrad(g, dy.cuda(), wt.cuda(), (h, w), math=ops.MATH_AUTO)
    assert_close(f"narrow dgrad {geom} {cin}->{cout}", dx, x.grad, 3e-6)
    other = torch.randn(x.shape, generator=torch.Generator().manual_seed(7))
    buf = other.cuda()
    dx2 = ops.conv_dgrad(g, dy.cuda(), wt.cuda(), (h, w), residual=buf, math=ops.MATH_AUTO, out=buf)
    assert_close("narrow dgrad + residual in place", dx2, x.grad + other, 3e-6)


def test_narrow_falls_back_on_odd_widths():
    """Widths that are not multiples of 4 are outside the TMA box alignment: another engine takes them."""
    from robocupvision_b200 import _lib, ops
    g, x, wt, b = _mk("k3s1d1", 3, 8, 2, 9, 7)
    assert _engine(g, 2, 9, 7) != _lib.ENGINE_NARROW
    assert_close("odd width", ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), math=ops.MATH_AUTO),
                 _ref_conv("k3s1d1", x, wt, b), 2e-6)


WGRAD_CASES = [
    # geom, cin, cout, n, h, w -- the dense ("row") side has <= 16 channels: Cout for a conv, Cin for a transposed conv
    ("k3s1d1", 3, 8, 2, 12, 20),     # 3 roles, two pixel parts
    ("k3s1d1", 16, 16, 3, 30, 40),   # 4 channel chunks x 2 row-channel groups
    ("k3s1d1", 8, 5, 2, 7, 12),      # odd height, 5 row channels
    ("k3s1d1", 3, 8, 2, 120, 160),
    ("k3s1d1", 16, 16, 1, 6, 320),   # two column tiles
    ("k3s1d2", 3, 8, 2, 12, 20),
    ("k3s1d2", 16, 16, 2, 13, 24),
    ("k3s2", 8, 16, 2, 24, 40),
    ("k3s2", 3, 8, 2, 26, 24),
    ("convT", 16, 8, 2, 12, 20),     # src = dy (8 ch, fine grid), row = x (16 ch, coarse grid)
    ("convT", 8, 32, 2, 15, 20),     # 8 row channels, 32 src channels
    ("k1", 8, 5, 2, 12, 20),
    ("k1", 16, 5, 2, 120, 160),
]


@pytest.mark.parametrize("geom,cin,cout,n,h,w", WGRAD_CASES)
def test_narrow_wgrad(geom, cin, cout, n, h, w):
    from robocupvision_b200 import _lib, ops
    g, x, wt, b = _mk(geom, cin, cout, n, h, w, seed=15)
    assert _engine(g, n, h, w, 2) == _lib.ENGINE_NARROW
    wt.requires_grad_(True); b.requires_grad_(True)
    y = _ref_conv(geom, x, wt, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(16))
    y.backward(dy)
    dw, db = ops.conv_wgrad(g, x.cuda(), dy.cuda(), want_bias=True, math=ops.MATH_AUTO)
    assert_close(f"narrow wgrad {geom} {cin}->{cout}", dw, wt.grad, 1e-5)
    assert_close(f"narrow bgrad {geom}", db, b.grad, 1e-5)
    # accumulates into a caller-provided gradient
    dw2 = torch.ones_like(dw)
    ops.conv_wgrad(g, x.cuda(), dy.cuda(), dw=dw2, math=ops.MATH_AUTO)
    assert_close("narrow wgrad accumulate", dw2, wt.grad + 1.0, 1e-5)


FULL_SIZE_LAYERS = [  # ROBO-UNet 160x120 layer table at the bench batch
    ("k3s1d1", 3, 8, 120, 160), ("k3s2", 8, 16, 120, 160), ("k3s1d1", 16, 16, 60, 80), ("k3s2", 16, 32, 60, 80),
    ("k3s1d1", 32, 32, 30, 40), ("k3s2", 32, 64, 30, 40), ("k3s1d1", 64, 64, 15, 20), ("k3s1d1", 64, 128, 15, 20),
    ("k3s1d1", 128, 128, 15, 20), ("k3s1d1", 128, 64, 15, 20), ("convT", 64, 32, 15, 20), ("convT", 32, 16, 30, 40),
    ("convT", 16, 8, 60, 80), ("k1", 8, 5, 120, 160),
]


@pytest.mark.parametrize("geom,cin,cout,h,w", FULL_SIZE_LAYERS)
def test_full_size_adjoint_identities(geom, cin, cout, h, w):
    """Size-independent properties at the BASELINE batch (64 frames), where the CPU oracle is too slow: a
    convolution is bilinear, so <dy, conv(x; w)> = <dgrad(dy; w), x> = <wgrad(x, dy), w>, and conv(x1 + x2) =
    conv(x1) + conv(x2).  Each engine (tensor-core, narrow-layer) computes the three sides with different
    kernels and different tilings of the 1.2 M-pixel grid; inner products are taken in fp64 on the device."""
    from robocupvision_b200 import ops
    n = 64
    k, s, p, d, tr = GEOMS[geom]
    g = ops.ConvGeom(cin, cout, k, s, p, d, tr)
    gen = torch.Generator(device="cuda").manual_seed(cin * 1000 + cout)
    x = torch.randn(n, cin, h, w, device="cuda", generator=gen)
    x2 = torch.randn(n, cin, h, w, device="cuda", generator=gen)
    wt = torch.randn((cin, cout, 3, 3) if tr else (cout, cin, k, k), device="cuda", generator=gen) / (cin * k * k) ** 0.5
    eng = [ops.conv_engine(g, n, h, w, dd, ops.MATH_AUTO) for dd in (0, 1, 2)]
    wp = {dd: ops.conv_pack(g, wt, dd, nhw=(n, h, w)) for dd in (0, 1) if eng[dd] == ops.ENGINE_UMMA}
    y = ops.conv_fwd(g, x, wt, None, math=ops.MATH_AUTO, wpacked=wp.get(0))
    dy = torch.randn(y.shape, device="cuda", generator=gen)
    dx = ops.conv_dgrad(g, dy, wt, (h, w), math=ops.MATH_AUTO, wpacked=wp.get(1))
    dw, _ = ops.conv_wgrad(g, x, dy, math=ops.MATH_AUTO)
    a = float((dy.double() * y.double()).sum())
    b = float((dx.double() * x.double()).sum())
    c = float((dw.double() * wt.double()).sum())
    scale = float(dy.double().norm() * y.double().norm())
    assert abs(a - b) <= 2e-6 * scale and abs(a - c) <= 2e-6 * scale, (a, b, c, scale, eng)
    y12 = ops.conv_fwd(g, x + x2, wt, None, math=ops.MATH_AUTO, wpacked=wp.get(0))
    y2 = ops.conv_fwd(g, x2, wt, None, math=ops.MATH_AUTO, wpacked=wp.get(0))
    assert_close(f"linearity {geom} {cin}->{cout}", y12, y + y2, 1e-5)
