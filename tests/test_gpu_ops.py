"""Per-kernel parity: every rcv_* entry point (through the C ABI) against the ATen CPU op the
reference delegates to, on seeded inputs.  fp32 tolerance: 1e-5 of the output range (the
reference's own accumulation-order noise is 1.4e-5)."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from util import assert_close

pytestmark = pytest.mark.gpu

GEOMS = {  # name -> (k, stride, pad, dil, transposed)
    "k3s1d1": (3, 1, 1, 1, False),
    "k3s1d2": (3, 1, 2, 2, False),
    "k3s2": (3, 2, 1, 1, False),
    "k1": (1, 1, 0, 1, False),
    "convT": (3, 2, 1, 1, True),
}
CHANS = [(3, 8), (8, 16), (16, 16), (32, 64), (64, 128), (128, 128), (128, 64), (16, 5), (24, 40)]
# engine -> (rcv_math, tolerance factor): "tc" = tcgen05 3xTF32 tiles with TMEM accumulators
# (the tensor core truncates on accumulate: ~3e-6 of the output range at K = 1152)
MATHS = {"simt": (0, 1.0), "tc": (1, 4.0)}


def _mk(geom, cin, cout, n, h, w, seed=0):
    from robocupvision_b200 import ops
    k, s, p, d, tr = GEOMS[geom]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wshape = (cin, cout, 3, 3) if tr else (cout, cin, k, k)
    wt = torch.randn(wshape, generator=g) * (1.0 / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g)
    return ops.ConvGeom(cin, cout, k, s, p, d, tr), x, wt, b


def _ref_conv(geom, x, w, b):
    k, s, p, d, tr = GEOMS[geom]
    if tr:
        return F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    return F.conv2d(x, w, b, s, p, d)


@pytest.mark.parametrize("engine", list(MATHS))
@pytest.mark.parametrize("geom", list(GEOMS))
@pytest.mark.parametrize("cin,cout", CHANS)
def test_conv_fwd(geom, cin, cout, engine):
    from robocupvision_b200 import ops
    math, f = MATHS[engine]
    g, x, w, b = _mk(geom, cin, cout, 3, 12, 20)
    ref = _ref_conv(geom, x, w, b)
    got = ops.conv_fwd(g, x.cuda(), w.cuda(), b.cuda(), math=math)
    assert_close(f"conv_fwd {geom} {cin}->{cout}", got, ref, 2e-6 * f)


@pytest.mark.parametrize("engine", list(MATHS))
@pytest.mark.parametrize("geom", ["k3s1d1", "k3s1d2", "k3s2", "k1", "convT"])
@pytest.mark.parametrize("hw", [(9, 7), (5, 3), (15, 20), (1, 1), (2, 130)])
def test_conv_fwd_ragged_sizes(geom, hw, engine):
    """odd extents (scalar store path), single pixels, M not a multiple of the tile."""
    from robocupvision_b200 import ops
    math, f = MATHS[engine]
    g, x, w, b = _mk(geom, 5, 7, 2, *hw, seed=3)
    ref = _ref_conv(geom, x, w, None)
    got = ops.conv_fwd(g, x.cuda(), w.cuda(), None, math=math)
    assert_close(f"conv_fwd {geom} {hw}", got, ref, 2e-6 * f)


def test_conv_wide_channel_tiles():
    """Cout > 128: several N tiles per pixel tile on the tensor-core engine."""
    from robocupvision_b200 import ops
    g, x, w, b = _mk("k3s1d1", 40, 200, 2, 9, 11, seed=11)
    ref = _ref_conv("k3s1d1", x, w, b)
    got = ops.conv_fwd(g, x.cuda(), w.cuda(), b.cuda(), math=1)
    assert_close("conv_fwd 40->200 tc", got, ref, 8e-6)


@pytest.mark.parametrize("math", [1, 2])  # RCV_MATH_TF32X3, RCV_MATH_AUTO (the fast modes run every tile whole)
@pytest.mark.parametrize("cin,cout,dil,nhw", [(128, 128, 1, (64, 15, 20)),   # 169 tiles on 148 SMs: left-over tiles split
                                              (64, 64, 1, (64, 15, 20)), (128, 64, 2, (64, 15, 20)),
                                              (128, 128, 1, (1, 15, 20)),    # 3 tiles: every tile split
                                              (64, 128, 2, (2, 15, 20)), (40, 200, 1, (1, 9, 11))])
def test_conv_split_reduction_workspace(cin, cout, dil, nhw, math):
    """rcv_conv_desc::workspace: the halo-staged kernel splits the reduction of left-over / few tiles across CTAs
    (partials through the workspace, last arriver runs the epilogue).  With and without the workspace the layer must
    agree to accumulation order, the workspace must be reusable without re-zeroing, and the fused epilogue (bias,
    ReLU, residual, BatchNorm statistics) must see every element exactly once."""
    from robocupvision_b200 import ops
    n, h, w_ = nhw
    g = ops.ConvGeom(cin, cout, 3, 1, dil, dil, False)
    gen = torch.Generator().manual_seed(cin * 7 + cout + n)
    x = torch.randn(n, cin, h, w_, generator=gen).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=gen) / (cin * 9) ** 0.5).cuda()
    b = torch.randn(cout, generator=gen).cuda()
    res = torch.randn(n, cout, h, w_, generator=gen).cuda()
    need = ops.conv_workspace_bytes(g, n, h, w_, ops.PACK_FWD, math)
    if cin % 32 == 0 and cin >= 64:
        assert need > 0, "this geometry is expected to use the split reduction"
    for fast in (ops.MATH_TF32, ops.MATH_BF16):
        assert ops.conv_workspace_bytes(g, n, h, w_, ops.PACK_FWD, fast) == 0
    ws = ops.new_workspace(max(need, 1024), x.device)
    wp = ops.conv_pack(g, wt, ops.PACK_FWD, math=math, nhw=nhw)
    outs = []
    for wsp in (None, ws, ws):
        stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        y = ops.conv_fwd(g, x, wt, b, epilogue=ops.EPI_RELU, residual=res, stats=stats, math=math, wpacked=wp,
                         workspace=wsp)
        outs.append((y, stats))
    torch.cuda.synchronize()
    assert int(ws[:1024].view(torch.int32).abs().sum()) == 0, "arrival counters must be left zeroed"
    for y, st in outs[1:]:
        assert_close("split vs whole", y, outs[0][0], 2e-6)
        assert_close("split stats", st, outs[0][1], 1e-6, atol=1e-3)
    assert torch.equal(outs[1][0], outs[2][0]), "same shares, same order: bitwise reproducible"
    if cin % 32 == 0:
        ref = F.relu(F.conv2d(x.cpu(), wt.cpu(), b.cpu(), 1, dil, dil)) + res.cpu()
        assert_close("split vs fp32", outs[1][0], ref, 8e-6)
    # input gradient through the same path
    dy = torch.randn(n, cout, h, w_, generator=gen).cuda()
    wpd = ops.conv_pack(g, wt, ops.PACK_DGRAD, math=math, nhw=nhw)
    ws2 = ops.new_workspace(max(ops.conv_workspace_bytes(g, n, h, w_, ops.PACK_DGRAD, math), 1024), x.device)
    d0 = ops.conv_dgrad(g, dy, wt, (h, w_), math=math, wpacked=wpd)
    d1 = ops.conv_dgrad(g, dy, wt, (h, w_), math=math, wpacked=wpd, workspace=ws2)
    assert_close("split dgrad", d1, d0, 4e-6)


@pytest.mark.parametrize("cin,cout,nhw", [(16, 32, (64, 60, 80)), (32, 64, (64, 30, 40)), (32, 32, (8, 60, 80)),
                                          (64, 64, (4, 30, 40)), (128, 128, (2, 30, 40)), (16, 32, (2, 6, 8))])
@pytest.mark.parametrize("math", [1, 3])  # RCV_MATH_TF32X3 (parity), RCV_MATH_TF32 (fast)
def test_stride2_conv_at_real_sizes(cin, cout, nhw, math):
    """Stride-2 3x3 convs on the tensor-core engine at the nets' real sizes: forward with the fused epilogue and
    BatchNorm statistics against F.conv2d, and the transposed convolution's input gradient (the same problem)."""
    from robocupvision_b200 import ops
    n, h, w_ = nhw
    g = ops.ConvGeom(cin, cout, 3, 2, 1, 1, False)
    gen = torch.Generator().manual_seed(cin * 3 + cout)
    x = torch.randn(n, cin, h, w_, generator=gen)
    wt = torch.randn(cout, cin, 3, 3, generator=gen) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=gen)
    assert ops.conv_engine(g, n, h, w_, ops.PACK_FWD, math) == ops.ENGINE_UMMA
    wp = ops.conv_pack(g, wt.cuda(), ops.PACK_FWD, math=math, nhw=nhw)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    got = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), epilogue=ops.EPI_RELU, stats=stats, math=math, wpacked=wp)
    ref = F.relu(F.conv2d(x, wt, b, 2, 1))
    tol = 8e-6 if math == 1 else 2e-3
    assert_close(f"s2 conv {cin}->{cout} {nhw}", got, ref, tol)
    assert_close("s2 stats", stats[:cout], ref.double().sum((0, 2, 3)), tol, atol=1e-2)
    # ConvTranspose2d(cout -> cin)'s input gradient is a stride-2 conv over dy with the transposed weight layout
    gt = ops.ConvGeom(cout, cin, 3, 2, 1, 1, True)
    wtt = torch.randn(cout, cin, 3, 3, generator=gen) / (cin * 9) ** 0.5          # (Cin_T = cout, Cout_T = cin, 3, 3)
    xt = torch.randn(n, cout, h // 2, w_ // 2, generator=gen, requires_grad=True)
    yt = F.conv_transpose2d(xt, wtt, None, stride=2, padding=1, output_padding=1)
    dyt = torch.randn(yt.shape, generator=gen)
    yt.backward(dyt)
    if ops.conv_engine(gt, n, h // 2, w_ // 2, ops.PACK_DGRAD, math) == ops.ENGINE_UMMA:
        wpd = ops.conv_pack(gt, wtt.cuda(), ops.PACK_DGRAD, math=math, nhw=(n, h // 2, w_ // 2))
        dx = ops.conv_dgrad(gt, dyt.cuda(), wtt.cuda(), (h // 2, w_ // 2), math=math, wpacked=wpd)
        assert_close(f"convT dgrad {cout}->{cin}", dx, xt.grad, 1.2e-5 if math == 1 else 3e-3)


@pytest.mark.parametrize("math", [2, 3])  # RCV_MATH_AUTO (parity: 3xTF32), RCV_MATH_TF32 (fast)
@pytest.mark.parametrize("cout", [16, 8, 5])
@pytest.mark.parametrize("nhw", [(64, 60, 80), (3, 12, 20), (5, 30, 40), (2, 9, 7), (1, 1, 1), (300, 6, 8)])
def test_conv16_persistent_tensor_core_kernel(nhw, cout, math):
    """16 -> <= 16 stride-1 3x3 layers on the persistent tensor-core kernel (resident weight panel, double-buffered
    patch and accumulators): forward with every fused epilogue piece (bias, ReLU, residual, BatchNorm statistics)
    and the input gradient, against ATen; many tiles per CTA (300 images), odd widths, a single pixel."""
    from robocupvision_b200 import ops
    n, h, w_ = nhw
    g = ops.ConvGeom(16, cout, 3, 1, 1, 1, False)
    gen = torch.Generator().manual_seed(n + cout)
    x = torch.randn(n, 16, h, w_, generator=gen)
    wt = torch.randn(cout, 16, 3, 3, generator=gen) / 12.0
    b = torch.randn(cout, generator=gen)
    res = torch.randn(n, cout, h, w_, generator=gen)
    assert ops.conv_engine(g, n, h, w_, ops.PACK_FWD, math) == ops.ENGINE_UMMA
    wp = ops.conv_pack(g, wt.cuda(), ops.PACK_FWD, math=math, nhw=nhw)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    got = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), epilogue=ops.EPI_RELU, residual=res.cuda(), stats=stats,
                       math=math, wpacked=wp)
    ref = F.relu(F.conv2d(x, wt, b, 1, 1)) + res
    tol = 8e-6 if math == 2 else 2e-3
    assert_close(f"conv16 fwd -> {cout} {nhw}", got, ref, tol)
    rd = ref.double()
    assert_close("conv16 stats sum", stats[:cout], rd.sum((0, 2, 3)), tol, atol=1e-2)
    assert_close("conv16 stats sumsq", stats[cout:], (rd * rd).sum((0, 2, 3)), tol, atol=1e-2)
    # the input gradient of a (cout' -> 16) layer reduces over 16 channels as well: take cin = cout
    g2 = ops.ConvGeom(cout, 16, 3, 1, 1, 1, False)
    w2 = torch.randn(16, cout, 3, 3, generator=gen) / 12.0
    dy = torch.randn(n, 16, h, w_, generator=gen)
    if ops.conv_engine(g2, n, h, w_, ops.PACK_DGRAD, math) == ops.ENGINE_UMMA:
        wpd = ops.conv_pack(g2, w2.cuda(), ops.PACK_DGRAD, math=math, nhw=nhw)
        prev = torch.randn(n, cout, h, w_, generator=gen)
        dx = ops.conv_dgrad(g2, dy.cuda(), w2.cuda(), (h, w_), residual=prev.cuda(), math=math, wpacked=wpd)
        dref = torch.nn.grad.conv2d_input((n, cout, h, w_), w2, dy, 1, 1) + prev
        assert_close(f"conv16 dgrad {nhw}", dx, dref, 1.2e-5 if math == 2 else 3e-3)


def test_pack_table_matches_single_packs():
    """rcv_conv_pack_table_* (all layers in one launch) writes the same panels as rcv_conv_pack."""
    from robocupvision_b200 import ops
    jobs, singles = [], []
    for i, (geom, cin, cout, d) in enumerate([("k3s1d1", 128, 128, 0), ("k3s1d1", 128, 128, 1), ("convT", 64, 32, 0),
                                              ("k3s2", 16, 32, 1), ("k3s2", 16, 32, 0), ("k3s2", 64, 128, 0), ("k3s1d2", 24, 40, 0), ("k1", 16, 5, 0)]):
        g, x, w, b = _mk(geom, cin, cout, 1, 4, 4, seed=20 + i)
        w = w.cuda()
        singles.append(ops.conv_pack(g, w, d))
        jobs.append((g, d, w, torch.zeros_like(singles[-1])))
    tbl = ops.PackTable(jobs)
    tbl.run()
    for (g, d, w, pk), ref in zip(jobs, singles):
        assert torch.equal(pk, ref)


def test_tc_requires_packed_weights():
    """RCV_MATH_TF32X3 through the raw C ABI without a packed panel is an error, not a fallback."""
    import ctypes as C
    from robocupvision_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(1, 8, 4, 4, device="cuda"); w = torch.zeros(8, 8, 3, 3, device="cuda"); y = torch.zeros(1, 8, 4, 4, device="cuda")
    d = _lib.ConvDesc(1, 8, 4, 4, 8, 3, 1, 1, 1, 0, 0, _lib.MATH_TF32X3)
    rc = lib.rcv_conv_fwd(C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), None, None, None, None,
                          None, C.c_void_p(y.data_ptr()), None, None)
    assert rc == _lib.RCV_ERR_BAD_ARG and b"packed" in lib.rcv_last_error()


@pytest.mark.parametrize("engine", list(MATHS))
@pytest.mark.parametrize("epi", ["none", "relu", "relu_affine", "affine_relu", "affine"])
@pytest.mark.parametrize("geom,cin,cout", [("k3s1d1", 8, 16), ("k3s1d2", 64, 128), ("convT", 32, 16), ("k3s2", 16, 32)])
def test_conv_epilogues(epi, geom, cin, cout, engine):
    from robocupvision_b200 import ops
    math, f = MATHS[engine]
    g, x, w, b = _mk(geom, cin, cout, 2, 12, 16, seed=1)
    gen = torch.Generator().manual_seed(9)
    sc, sh = torch.randn(cout, generator=gen), torch.randn(cout, generator=gen)
    v = _ref_conv(geom, x, w, b)
    res = torch.randn(v.shape, generator=gen)
    A, B = sc.view(1, -1, 1, 1), sh.view(1, -1, 1, 1)
    ref = {"none": v, "relu": F.relu(v), "relu_affine": A * F.relu(v) + B, "affine_relu": F.relu(A * v + B),
           "affine": A * v + B}[epi] + res
    code = {"none": ops.EPI_NONE, "relu": ops.EPI_RELU, "relu_affine": ops.EPI_RELU_AFFINE,
            "affine_relu": ops.EPI_AFFINE_RELU, "affine": ops.EPI_AFFINE}[epi]
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    got = ops.conv_fwd(g, x.cuda(), w.cuda(), b.cuda(), epilogue=code, scale=sc.cuda(), shift=sh.cuda(),
                       residual=res.cuda(), stats=stats, math=math)
    assert_close(f"epilogue {epi} {geom}", got, ref, 3e-6 * f)
    rd = ref.double()
    assert_close("stats sum", stats[:cout], rd.sum((0, 2, 3)), 1e-6 * f, atol=1e-3)
    assert_close("stats sumsq", stats[cout:], (rd * rd).sum((0, 2, 3)), 1e-6 * f, atol=1e-3)


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("geom,cin,cout,nhw", [("k3s1d1", 32, 64, (4, 30, 40)), ("k3s1d2", 64, 128, (3, 15, 20)),
                                               ("k3s1d1", 128, 128, (2, 9, 7)), ("k3s1d2", 128, 64, (5, 12, 16)),
                                               ("k3s1d1", 64, 32, (2, 24, 50))])
def test_conv_fwd_normalise_on_load(geom, cin, cout, nhw, relu):
    """rcv_conv_fwd_nl: the producer block's BatchNorm (per-channel scale/shift [+ReLU]) applied to the input while
    the halo-staged tensor-core kernel stages it == the conv of the normalised tensor; the zero padding stays zero."""
    from robocupvision_b200 import ops
    n, h, w_ = nhw
    g, x, w, b = _mk(geom, cin, cout, n, h, w_, seed=4)
    if not ops.conv_normalises_on_load(g, n, h, w_, ops.MATH_AUTO):
        pytest.skip("layer not on the halo-staged tensor-core kernel at this size")
    gen = torch.Generator().manual_seed(11)
    sc, sh = torch.randn(cin, generator=gen), torch.randn(cin, generator=gen)
    xn = sc.view(1, -1, 1, 1) * x + sh.view(1, -1, 1, 1)
    if relu:
        xn = F.relu(xn)
    ref = F.relu(_ref_conv(geom, xn, w, b))
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    wp = ops.conv_pack(g, w.cuda(), ops.PACK_FWD, nhw=(n, h, w_))
    got = ops.conv_fwd(g, x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_RELU, stats=stats, math=ops.MATH_AUTO,
                       wpacked=wp, in_affine=(sc.cuda(), sh.cuda(), relu))
    assert_close(f"conv_fwd_nl {geom} {cin}->{cout} relu={relu}", got, ref, 8e-6)
    same = ops.conv_fwd(g, xn.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_RELU, math=ops.MATH_AUTO, wpacked=wp)
    assert_close("conv_fwd_nl vs conv_fwd of the normalised tensor", got, same.cpu(), 2e-6)
    assert_close("stats sum", stats[:cout], ref.double().sum((0, 2, 3)), 4e-6, atol=1e-3)


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("geom,cin,cout,nhw", [("k3s1d1", 32, 64, (4, 30, 40)), ("k3s1d2", 64, 128, (3, 15, 20)),
                                               ("k3s1d1", 128, 128, (2, 9, 8)), ("k3s1d2", 128, 64, (5, 12, 16)),
                                               ("k3s1d1", 24, 40, (2, 18, 24)), ("k3s1d1", 64, 32, (2, 24, 52))])
def test_conv_wgrad_normalise_on_load(geom, cin, cout, nhw, relu):
    """rcv_conv_wgrad_nl: weight / bias gradient with the producer block's BatchNorm applied to x on load == the
    gradient computed from the normalised tensor (CPU autograd and the plain entry point)."""
    from robocupvision_b200 import ops
    n, h, w_ = nhw
    g, x, w, b = _mk(geom, cin, cout, n, h, w_, seed=6)
    assert ops.conv_wgrad_normalises_on_load(g, n, h, w_, ops.MATH_AUTO)
    gen = torch.Generator().manual_seed(13)
    sc, sh = torch.randn(cin, generator=gen), torch.randn(cin, generator=gen)
    xn = sc.view(1, -1, 1, 1) * x + sh.view(1, -1, 1, 1)
    if relu:
        xn = F.relu(xn)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    out = _ref_conv(geom, xn, wr, br)
    dy = torch.randn(out.shape, generator=gen)
    out.backward(dy)
    dw, db = ops.conv_wgrad(g, x.cuda(), dy.cuda(), want_bias=True, math=ops.MATH_AUTO,
                            in_affine=(sc.cuda(), sh.cuda(), relu))
    assert_close(f"wgrad_nl {geom} {cin}->{cout} relu={relu}", dw, wr.grad, 8e-6)
    assert_close("dbias", db, br.grad, 4e-6)
    dw2, _ = ops.conv_wgrad(g, xn.cuda(), dy.cuda(), want_bias=True, math=ops.MATH_AUTO)
    assert_close("wgrad_nl vs wgrad of the normalised tensor", dw, dw2.cpu(), 4e-6)


def test_conv_wgrad_normalise_on_load_refused_elsewhere():
    from robocupvision_b200 import ops
    for geom, cin, cout, math in [("k3s2", 32, 64, ops.MATH_AUTO), ("k3s1d1", 16, 8, ops.MATH_AUTO),
                                  ("k3s1d1", 32, 64, ops.MATH_FP32), ("convT", 32, 16, ops.MATH_AUTO),
                                  ("k1", 64, 64, ops.MATH_AUTO)]:
        g, x, w, b = _mk(geom, cin, cout, 2, 12, 16)
        assert not ops.conv_wgrad_normalises_on_load(g, 2, 12, 16, math)
        one = torch.ones(cin, device="cuda")
        dy = torch.zeros(2, cout, *g.out_hw(12, 16), device="cuda")
        with pytest.raises(RuntimeError):
            ops.conv_wgrad(g, x.cuda(), dy, math=math, in_affine=(one, one, False))
    g, x, w, b = _mk("k3s1d1", 32, 64, 2, 9, 7)   # odd row width: element-wise gather
    assert not ops.conv_wgrad_normalises_on_load(g, 2, 9, 7, ops.MATH_AUTO)


def test_conv_fwd_normalise_on_load_refused_elsewhere():
    """Engines other than the halo-staged tensor-core kernel refuse an input transform (no silent ignore)."""
    from robocupvision_b200 import ops
    for geom, cin, cout, math in [("k3s2", 32, 64, ops.MATH_AUTO), ("k3s1d1", 3, 8, ops.MATH_AUTO),
                                  ("k3s1d1", 32, 64, ops.MATH_FP32), ("convT", 32, 16, ops.MATH_AUTO)]:
        g, x, w, b = _mk(geom, cin, cout, 2, 12, 16)
        assert not ops.conv_normalises_on_load(g, 2, 12, 16, math)
        one = torch.ones(cin, device="cuda")
        with pytest.raises(RuntimeError):
            ops.conv_fwd(g, x.cuda(), w.cuda(), b.cuda(), math=math, in_affine=(one, one, False))


@pytest.mark.parametrize("engine", list(MATHS))
@pytest.mark.parametrize("geom", list(GEOMS))
@pytest.mark.parametrize("cin,cout", [(3, 8), (8, 16), (32, 64), (128, 128), (64, 32), (16, 5)])
def test_conv_dgrad_wgrad(geom, cin, cout, engine):
    from robocupvision_b200 import ops
    math, f = MATHS[engine]
    g, x, w, b = _mk(geom, cin, cout, 3, 12, 20, seed=5)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = _ref_conv(geom, x, w, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
    y.backward(dy)
    dx = ops.conv_dgrad(g, dy.cuda(), w.detach().cuda(), (12, 20), math=math)
    assert_close(f"dgrad {geom} {cin}->{cout}", dx, x.grad, 3e-6 * f)
    other = torch.randn(x.shape, generator=torch.Generator().manual_seed(7))
    dx2 = ops.conv_dgrad(g, dy.cuda(), w.detach().cuda(), (12, 20), residual=other.cuda(), math=math)
    assert_close("dgrad+residual", dx2, x.grad + other, 3e-6 * f)
    dw, db = ops.conv_wgrad(g, x.detach().cuda(), dy.cuda(), want_bias=True, math=math)
    assert_close(f"wgrad {geom} {cin}->{cout}", dw, w.grad, 1e-5)
    assert_close(f"bgrad {geom}", db, b.grad, 1e-5)


@pytest.mark.parametrize("geom,hw", [("k3s1d1", (18, 24)), ("k3s1d2", (18, 24)), ("k3s1d1", (7, 12)), ("k3s1d2", (5, 8))])
def test_wgrad_tc_quad_gather(geom, hw):
    """tensor-core wgrad, aligned-quad gather (stride-1 3x3, W % 4 == 0): row wraps inside a pixel chunk,
    chunk-edge neighbour loads, several pixel splits, a partial last k tile (47 channels = 141 units)."""
    from robocupvision_b200 import ops
    g, x, w, b = _mk(geom, 47, 24, 5, *hw, seed=13)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = _ref_conv(geom, x, w, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
    y.backward(dy)
    dw, db = ops.conv_wgrad(g, x.detach().cuda(), dy.cuda(), want_bias=True, math=1)
    assert_close(f"wgrad quad {geom} {hw}", dw, w.grad, 1e-5)
    assert_close("bgrad quad", db, b.grad, 1e-5)


def test_wgrad_tc_split_and_tail():
    """tensor-core wgrad: several pixel splits, k tiles with a ragged tail, pixel count not a multiple of 32."""
    from robocupvision_b200 import ops
    g, x, w, b = _mk("k3s1d1", 24, 40, 5, 18, 22, seed=12)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = _ref_conv("k3s1d1", x, w, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
    y.backward(dy)
    dw, db = ops.conv_wgrad(g, x.detach().cuda(), dy.cuda(), want_bias=True, math=1)
    assert_close("wgrad tc", dw, w.grad, 1e-5)
    assert_close("bgrad tc", db, b.grad, 1e-5)
    with pytest.raises(Exception, match="multiple of 4"):
        g2, x2, w2, b2 = _mk("k3s1d1", 6, 10, 3, 7, 9, seed=8)
        ops.conv_wgrad(g2, x2.cuda(), torch.zeros(3, 10, 7, 9, device="cuda"), math=1)


@pytest.mark.parametrize("geom", ["k3s1d1", "k3s1d2", "k1"])
def test_wgrad_ragged(geom):
    """pixel count not a multiple of 4/16: scalar load path and slab tails."""
    from robocupvision_b200 import ops
    g, x, w, b = _mk(geom, 6, 10, 3, 7, 9, seed=8)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = _ref_conv(geom, x, w, b)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
    y.backward(dy)
    dw, db = ops.conv_wgrad(g, x.detach().cuda(), dy.cuda(), want_bias=True)
    assert_close("wgrad ragged", dw, w.grad, 1e-5)
    assert_close("bgrad ragged", db, b.grad, 1e-5)
    dx = ops.conv_dgrad(g, dy.cuda(), w.detach().cuda(), (7, 9))
    assert_close("dgrad ragged", dx, x.grad, 3e-6)


def test_conv_rejects_unsupported():
    from robocupvision_b200 import ops, _lib
    g = ops.ConvGeom(4, 4, 5, 1, 2, 1)
    x = torch.zeros(1, 4, 8, 8, device="cuda")
    w = torch.zeros(4, 4, 5, 5, device="cuda")
    with pytest.raises(_lib.RcvError) as e:
        ops.conv_fwd(g, x, w)
    assert e.value.code == _lib.RCV_ERR_UNSUPPORTED
    with pytest.raises(RuntimeError):
        ops.conv_fwd(ops.ConvGeom(4, 4), x.cpu(), torch.zeros(4, 4, 3, 3))  # no CPU path


@pytest.mark.parametrize("order", ["relu_affine", "affine_relu"])
@pytest.mark.parametrize("shape", [(4, 8, 12, 20), (3, 128, 5, 4), (2, 5, 7, 9)])
def test_bn_train_fwd_bwd(order, shape):
    """conv-output statistics -> finalize -> apply, and the two-pass backward, against
    F.batch_norm(training=True) composed with ReLU in the block's order."""
    from robocupvision_b200 import ops
    n, c, h, w = shape
    gen = torch.Generator().manual_seed(11)
    v = torch.randn(shape, generator=gen) * 2 + 0.3
    gamma = torch.rand(c, generator=gen) + 0.5
    beta = torch.randn(c, generator=gen)
    rm, rv = torch.randn(c, generator=gen), torch.rand(c, generator=gen) + 0.5
    res = torch.randn(shape, generator=gen)
    dy = torch.randn(shape, generator=gen)
    vr = v.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    if order == "relu_affine":
        z_ref = F.relu(vr)
        y_ref = F.batch_norm(z_ref, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5) + res
    else:
        z_ref = vr
        y_ref = F.relu(F.batch_norm(z_ref, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)) + res
    y_ref.backward(dy)

    z = z_ref.detach().cuda()
    zd = z.double()
    stats = torch.cat([zd.sum((0, 2, 3)), (zd * zd).sum((0, 2, 3))])
    rm_g, rv_g = rm.cuda(), rv.cuda()
    scale, shift, mean, invstd = ops.bn_finalize(stats, n * h * w, gamma.cuda(), beta.cuda(), rm_g, rv_g, 0.1, 1e-5)
    y = ops.bn_apply(z, scale, shift, relu=(order == "affine_relu"), residual=res.cuda())
    assert_close("bn y", y, y_ref, 3e-6)
    assert_close("running_mean", rm_g, rm_ref, 1e-6)
    assert_close("running_var", rv_g, rv_ref, 1e-6)
    # the fused single-launch form must agree bit for bit with finalize + apply
    rm_f, rv_f = rm.cuda(), rv.cuda()
    y2, sc2, sh2, mean2, inv2 = ops.bn_finalize_apply(z, stats, gamma.cuda(), beta.cuda(), rm_f, rv_f, 0.1, 1e-5,
                                                      relu=(order == "affine_relu"), residual=res.cuda())
    assert torch.equal(y2, y) and torch.equal(sc2, scale) and torch.equal(sh2, shift)
    assert torch.equal(mean2, mean) and torch.equal(inv2, invstd)
    assert torch.equal(rm_f, rm_g) and torch.equal(rv_f, rv_g)
    code = ops.EPI_RELU_AFFINE if order == "relu_affine" else ops.EPI_AFFINE_RELU
    dconv, dgamma, dbeta, dbias = ops.bn_bwd(code, dy.cuda(), z, scale, shift, mean, invstd, want_dbias=True)
    assert_close("bn dconv", dconv, vr.grad, 1e-5)
    assert_close("bn dgamma", dgamma, gr.grad, 1e-5)
    assert_close("bn dbeta", dbeta, br.grad, 1e-5)
    assert_close("bn dbias", dbias, vr.grad.sum((0, 2, 3)), 1e-5, atol=1e-4)


def test_bn_fold_eval():
    from robocupvision_b200 import ops
    gen = torch.Generator().manual_seed(2)
    c = 37
    gamma, beta, mean = (torch.randn(c, generator=gen) for _ in range(3))
    var = torch.rand(c, generator=gen) * 1e-3  # released checkpoints have var << eps
    x = torch.randn(2, c, 6, 10, generator=gen)
    ref = F.batch_norm(x, mean, var, gamma, beta, False, 0.1, 1e-5)
    sc, sh = ops.bn_fold(gamma.cuda(), beta.cuda(), mean.cuda(), var.cuda(), 1e-5)
    got = ops.bn_apply(x.cuda(), sc, sh, relu=False)
    assert_close("bn eval", got, ref, 3e-6)


@pytest.mark.parametrize("shape", [(2, 8, 12, 20), (3, 5, 6, 2), (1, 32, 30, 40)])
def test_maxpool(shape):
    from robocupvision_b200 import ops
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(shape, generator=gen)
    x[0, 0, 0, 0:2] = 1.5  # tie inside a window: first wins
    x[0, 0, 1, 0:2] = 1.5
    x[-1, -1, 2, 1] = float("nan")
    xr = x.clone().requires_grad_(True)
    y_ref, idx_ref = F.max_pool2d(xr, 2, 2, return_indices=True)
    dy = torch.randn(y_ref.shape, generator=gen)
    y_ref.backward(dy)
    y, idx, code = ops.maxpool2x2_fwd(x.cuda(), want_idx=True, want_code=True)
    assert torch.equal(idx.cpu(), idx_ref), "pool indices must be bit-exact"
    assert torch.equal(torch.nan_to_num(y.cpu(), nan=123.0), torch.nan_to_num(y_ref.detach(), nan=123.0))
    dx = ops.maxpool2x2_bwd(dy.cuda(), code, shape[2:])
    assert torch.equal(dx.cpu(), xr.grad)


@pytest.mark.parametrize("shape", [(2, 3, 8, 12), (3, 8, 120, 160), (1, 1, 2, 2), (2, 16, 30, 40)])
@pytest.mark.parametrize("use_code", [False, True])
def test_maxunpool_pair(shape, use_code):
    """The pool-index / unpool pair: rcv_maxunpool2x2_fwd/bwd on the indices (or 2-bit codes) of rcv_maxpool2x2_fwd
    against F.max_unpool2d and its autograd, bit-exact (pure data movement), with and without the fused skip add."""
    from robocupvision_b200 import ops
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(shape, generator=gen)
    y_ref, idx_ref = F.max_pool2d(x, 2, 2, return_indices=True)
    y, idx, code = ops.maxpool2x2_fwd(x.cuda(), want_idx=True, want_code=True)
    assert torch.equal(idx.cpu(), idx_ref)
    v = torch.randn(y_ref.shape, generator=gen).requires_grad_(True)
    up_ref = F.max_unpool2d(v, idx_ref, 2, 2)
    dout = torch.randn(shape, generator=gen)
    up_ref.backward(dout)
    kw = dict(code=code) if use_code else dict(idx=idx)
    up = ops.maxunpool2x2(v.detach().cuda(), **kw)
    assert torch.equal(up.cpu(), up_ref.detach())
    skip = torch.randn(shape, generator=gen)
    up2 = ops.maxunpool2x2(v.detach().cuda(), skip=skip.cuda(), **kw)
    assert torch.equal(up2.cpu(), up_ref.detach() + skip)
    dv = ops.maxunpool2x2_bwd(dout.cuda(), **kw)
    assert torch.equal(dv.cpu(), v.grad)


@pytest.mark.parametrize("shape", [(2, 3, 8, 12), (2, 16, 60, 80), (1, 2, 1, 1), (3, 5, 1, 7), (2, 4, 15, 20), (1, 3, 5, 3)])
def test_bilinear_upsample_2x(shape):
    """rcv_upsample_bilinear2x_fwd/bwd against F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) and
    its autograd; tolerance 1e-6 of the output range (the weights 1/4, 3/4 are exact, the sums round)."""
    from robocupvision_b200 import ops
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(shape, generator=gen, requires_grad=True)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    dout = torch.randn(ref.shape, generator=gen)
    ref.backward(dout)
    got = ops.upsample_bilinear2x(x.detach().cuda())
    assert_close("bilinear fwd", got, ref, 1e-6)
    skip = torch.randn(ref.shape, generator=gen)
    got2 = ops.upsample_bilinear2x(x.detach().cuda(), skip=skip.cuda())
    assert_close("bilinear fwd + skip", got2, ref.detach() + skip, 1e-6)
    dx = ops.upsample_bilinear2x_bwd(dout.cuda())
    assert_close("bilinear bwd", dx, x.grad, 2e-6)


def test_resample_modules_autograd():
    """PoolWithIndices -> MaxUnpool2x2(+skip) and UpsampleBilinear2x(+skip) as autograd modules against the torch.nn
    layers they stand for."""
    from robocupvision_b200.model import MaxUnpool2x2, PoolWithIndices, UpsampleBilinear2x
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 8, 24, 32, generator=gen)
    s = torch.randn(2, 8, 24, 32, generator=gen)
    g = torch.randn(2, 8, 24, 32, generator=gen)
    xr, sr = x.clone().requires_grad_(True), s.clone().requires_grad_(True)
    yr, ir = F.max_pool2d(xr, 2, 2, return_indices=True)
    outr = F.max_unpool2d(yr * 2.0, ir, 2, 2) + sr
    outr.backward(g)
    xg, sg = x.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    y, i = PoolWithIndices()(xg)
    out = MaxUnpool2x2()(y * 2.0, i, sg)
    out.backward(g.cuda())
    assert torch.equal(out.detach().cpu(), outr.detach()) and torch.equal(xg.grad.cpu(), xr.grad)
    assert torch.equal(sg.grad.cpu(), sr.grad)
    xr2 = x[:, :, :12, :16].clone().requires_grad_(True)
    r2 = F.interpolate(xr2, scale_factor=2, mode="bilinear", align_corners=False) + sr.detach()
    r2.backward(g)
    xg2 = x[:, :, :12, :16].contiguous().cuda().requires_grad_(True)
    o2 = UpsampleBilinear2x()(xg2, s.cuda())
    o2.backward(g.cuda())
    assert_close("module bilinear", o2, r2, 1e-6)
    assert_close("module bilinear grad", xg2.grad, xr2.grad, 2e-6)


@pytest.mark.parametrize("cin,classes,hw,weighted", [(8, 5, (24, 32), True), (8, 5, (5, 3), True), (16, 5, (12, 16), True),
                                                     (16, 3, (7, 5), False), (8, 2, (6, 8), False), (8, 8, (6, 8), True)])
def test_head_ce_train(cin, classes, hw, weighted):
    """rcv_head_ce_train + rcv_ce_weight_sum: classifier conv (model.py:259), CrossEntropyLoss2d (model.py:76-82),
    argmax / correct count (train.py:70-71) and their backward in one pass, against autograd over F.conv2d +
    F.cross_entropy in float64; labels outside [0, C) contribute nothing."""
    from robocupvision_b200 import ops
    h, w_ = hw
    gen = torch.Generator().manual_seed(cin * 100 + classes + h)
    n = 3
    f = torch.randn(n, cin, h, w_, generator=gen)
    wt = torch.randn(classes, cin, 1, 1, generator=gen) * 0.5
    b = torch.randn(classes, generator=gen)
    y = torch.randint(0, classes, (n, h, w_), generator=gen)
    cw = (torch.rand(classes, generator=gen) + 0.5) if weighted else None
    fd = f.double().requires_grad_(True)
    wd = wt.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    z = F.conv2d(fd, wd, bd)
    loss = F.cross_entropy(z, y, weight=None if cw is None else cw.double())
    loss.backward()
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    corr = torch.zeros(1, dtype=torch.int64, device="cuda")
    dw = torch.zeros(classes, cin, 1, 1, device="cuda")
    db = torch.zeros(classes, device="cuda")
    cwd = None if cw is None else cw.cuda()
    ops.ce_weight_sum(y.cuda(), cwd, sums[1:2], classes)
    df = ops.head_ce_train(f.cuda(), wt.cuda(), b.cuda(), y.cuda(), cwd, sums, corr, dw, db)
    s = sums.cpu()
    assert abs(float(s[0] / s[1]) - float(loss)) <= 2e-6 * abs(float(loss))
    wsum = float((cw.double()[y]).sum()) if weighted else float(y.numel())
    assert abs(float(s[1]) - wsum) <= 1e-11 * wsum
    assert int(corr) == int((z.argmax(1) == y).sum())
    assert_close("dfeat", df, fd.grad.float(), 2e-6)
    assert_close("dweight", dw, wd.grad.float(), 5e-6)
    assert_close("dbias", db, bd.grad.float(), 5e-6)


@pytest.mark.parametrize("hw", [(12, 16), (5, 3), (15, 20)])   # float4 rows, scalar rows, 300-element planes
def test_channel_copy(hw):
    """rcv_channel_copy: torch.cat([a, b], 1) of ROBO_UNet --v2 (model.py:507) and the slices autograd cuts back
    out of its gradient; bit-exact (a copy)."""
    from robocupvision_b200 import ops
    h, w_ = hw
    gen = torch.Generator().manual_seed(h + w_)
    a = torch.randn(3, 8, h, w_, generator=gen).cuda()
    b = torch.randn(3, 5, h, w_, generator=gen).cuda()
    cat = ops.concat_channels(a, b)
    assert torch.equal(cat, torch.cat([a, b], 1))
    assert torch.equal(ops.channel_slice(cat, 8, 5), b)
    assert torch.equal(ops.channel_slice(cat, 0, 8), a)
    assert torch.equal(ops.channel_slice(cat, 3, 7), cat[:, 3:10].contiguous())
    from robocupvision_b200 import _lib
    with pytest.raises(_lib.RcvError):
        ops.channel_slice(cat, 10, 5)


@pytest.mark.parametrize("hw", [(12, 16), (5, 3)])   # float4 path and the scalar path
def test_partial_channel_residual(hw):
    """rcv_*::res_channels: a residual with fewer channels than the output is added to the FIRST channels only --
    LabelProp's `x[:, 0:8] += top` (model.py:565) -- in the BatchNorm apply passes and in the narrow-layer engine's
    epilogue; the other engines refuse it."""
    from robocupvision_b200 import _lib, ops
    h, w_ = hw
    gen = torch.Generator().manual_seed(h * w_)
    n, c, rc = 3, 16, 8
    z = torch.randn(n, c, h, w_, generator=gen)
    sc, sh = torch.randn(c, generator=gen), torch.randn(c, generator=gen)
    top = torch.randn(n, rc, h, w_, generator=gen)
    ref = F.relu(sc.view(1, -1, 1, 1) * z + sh.view(1, -1, 1, 1))
    ref[:, :rc] += top
    got = ops.bn_apply(z.cuda(), sc.cuda(), sh.cuda(), True, residual=top.cuda())
    assert_close("bn_apply partial residual", got, ref, 1e-6)
    # train-mode finalize + apply
    stats = torch.stack([z.double().sum((0, 2, 3)), (z.double() ** 2).sum((0, 2, 3))]).reshape(-1).cuda()
    gam, bet = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    y, *_ = ops.bn_finalize_apply(z.cuda(), stats, gam.cuda(), bet.cuda(), None, None, 0.1, 1e-5, relu=True,
                                  residual=top.cuda())
    bn_ref = F.relu(F.batch_norm(z, None, None, gam, bet, True, 0.1, 1e-5))
    bn_ref[:, :rc] += top
    assert_close("bn_finalize_apply partial residual", y, bn_ref, 2e-6)
    if w_ % 4 == 0:
        # transposed conv 16 -> 16 on the narrow engine (LabelProp upConv3), eval-mode epilogue + partial skip
        g = ops.ConvGeom(16, 16, 3, 2, 1, 1, True)
        x = torch.randn(n, 16, h, w_, generator=gen)
        wt = torch.randn(16, 16, 3, 3, generator=gen) / 12.0
        b = torch.randn(16, generator=gen)
        topT = torch.randn(n, rc, 2 * h, 2 * w_, generator=gen)
        assert ops.conv_takes_partial_residual(g, n, h, w_)
        out = ops.conv_fwd(g, x.cuda(), wt.cuda(), b.cuda(), epilogue=ops.EPI_AFFINE_RELU, scale=sc.cuda(), shift=sh.cuda(),
                           residual=topT.cuda(), math=ops.MATH_AUTO)
        cref = F.relu(sc.view(1, -1, 1, 1) * F.conv_transpose2d(x, wt, b, stride=2, padding=1, output_padding=1) + sh.view(1, -1, 1, 1))
        cref[:, :rc] += topT
        assert_close("narrow convT partial residual", out, cref, 3e-6)
        # a tensor-core layer refuses
        g2 = ops.ConvGeom(64, 64, 3, 1, 1, 1, False)
        x2 = torch.randn(2, 64, h, w_, generator=gen).cuda()
        w2 = torch.randn(64, 64, 3, 3, generator=gen).cuda()
        with pytest.raises(_lib.RcvError):
            ops.conv_fwd(g2, x2, w2, None, residual=torch.zeros(2, 8, h, w_, device="cuda"), math=ops.MATH_TF32X3)


@pytest.mark.parametrize("n,c", [(1, 5), (7, 5), (70, 5), (3, 2), (5, 8)])
def test_metric_tail_matches_reference_rule(n, c):
    """rcv_metric_tail: the validation loops' per-image IoU rule (train.py:148-153: inter / union per class, an image
    without the class counts 1) and the mean loss, against the oracle's restatement; counts are exact integers."""
    from robocupvision_b200 import ops
    from oracle import ref_metrics
    gen = torch.Generator().manual_seed(n * 10 + c)
    conf = torch.randint(0, 5000, (n, c, c), generator=gen, dtype=torch.int64)
    conf[0, :, 0] = 0
    conf[0, 0, :] = 0          # class 0 absent from image 0: union == 0 -> counts 1
    sums = torch.tensor([123.456, 78.9], dtype=torch.float64)
    iou, loss = ops.metric_tail(conf.cuda(), sums.cuda())
    ref = ref_metrics.iou_sums(conf.numpy())
    assert torch.allclose(iou.cpu(), torch.as_tensor(ref, dtype=torch.float64), rtol=0, atol=1e-12)
    assert abs(float(loss) - 123.456 / 78.9) < 1e-15


@pytest.mark.parametrize("c", [5, 4, 2, 8])
@pytest.mark.parametrize("weighted", [True, False])
def test_cross_entropy_argmax_confusion(c, weighted):
    from robocupvision_b200 import ops
    from oracle import ref_metrics, ref_model
    gen = torch.Generator().manual_seed(21)
    n, h, w = 3, 15, 20
    logits = torch.randn(n, c, h, w, generator=gen) * 3
    logits[0, :, 0, 0] = 0.25  # exact tie: argmax must be class 0
    target = torch.randint(0, c, (n, h, w), generator=gen)
    cw = (torch.rand(c, generator=gen) * 10 + 0.5) if weighted else None
    lr = logits.clone().requires_grad_(True)
    loss_ref = ref_model.cross_entropy_2d(lr, target, cw)
    (loss_ref * 1.7).backward()
    sums, am, conf, corr = ops.ce_fwd(logits.cuda(), target.cuda(), None if cw is None else cw.cuda(),
                                      want_argmax=True, want_conf=True, want_correct=True)
    loss = (sums[0] / sums[1]).item()
    assert abs(loss - loss_ref.item()) <= 2e-6 * max(1, abs(loss_ref.item()))
    am_ref = ref_metrics.argmax_first(logits.numpy())
    assert (am.cpu().numpy() == am_ref).all()
    conf_ref = ref_metrics.confusion_per_image(am_ref, target.numpy(), c)
    assert (conf.cpu().numpy() == conf_ref).all()
    assert int(corr) == int((am_ref == target.numpy()).sum())
    conf2 = ops.confusion(am, target.cuda(), c)
    assert (conf2.cpu().numpy() == conf_ref).all()
    dl = ops.ce_bwd(logits.cuda(), target.cuda(), None if cw is None else cw.cuda(), sums,
                    gscale=torch.tensor(1.7, device="cuda"))
    assert_close("ce dlogits", dl, lr.grad, 1e-5, atol=1e-9)


def test_adam_l1_matches_torch_adam():
    from robocupvision_b200 import ops
    gen = torch.Generator().manual_seed(31)
    n = 10007
    p0 = torch.randn(n, generator=gen)
    p0[::17] = 0.0
    mask = torch.rand(n, generator=gen) < 0.2
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p, m, v = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n, generator=gen)
        opt.zero_grad()
        loss = (p_ref * g).sum() + 1e-3 * p_ref.abs().sum()
        loss.backward()
        p_ref.grad[mask] = 0
        l1_ref = float(p_ref.detach().abs().sum())
        opt.step()
        l1 = torch.zeros(1, dtype=torch.float64, device="cuda")
        ops.adam_l1_step(p, g.cuda(), m, v, lr=1e-3, step=step, l1_decay=1e-3, mask=mask.cuda().to(torch.uint8),
                         l1_sum=l1)
        assert abs(float(l1) - l1_ref) <= 1e-5 * l1_ref
        assert_close(f"adam step {step}", p, p_ref, 2e-6)
    # device-resident step / lr (graph-replayable form)
    step_dev = torch.tensor([6], dtype=torch.int32, device="cuda")
    lr_dev = torch.tensor([1e-3], device="cuda")
    g = torch.randn(n, generator=gen)
    opt.zero_grad(); (p_ref * g).sum().backward(); opt.step()
    ops.adam_l1_step(p, g.cuda(), m, v, lr=123.0, step=0, step_dev=step_dev, lr_dev=lr_dev)
    assert_close("adam dev step", p, p_ref, 2e-6)


def test_sgd_matches_torch_sgd():
    from robocupvision_b200 import ops
    gen = torch.Generator().manual_seed(32)
    n = 5003
    p0 = torch.randn(n, generator=gen)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([p_ref], lr=0.1, momentum=0.5, weight_decay=1e-3)  # trainer.py:182-184
    p, buf = p0.cuda(), torch.zeros(n, device="cuda")
    for step in range(4):
        g = torch.randn(n, generator=gen)
        opt.zero_grad(); (p_ref * g).sum().backward(); opt.step()
        ops.sgd_step(p, g.cuda(), buf, lr=0.1, momentum=0.5, weight_decay=1e-3, first_step=(step == 0))
        assert_close(f"sgd step {step}", p, p_ref, 2e-6)


def test_label_ops():
    """maskLabel LUT kernel, labelToPred and the LabelProp batch assembly against the reference's own
    tensor code restated on CPU (transform.py:26-49, 172-183; labelPropTrain.py:178-193)."""
    from robocupvision_b200 import ops
    g = torch.Generator().manual_seed(3)
    lab = torch.randint(0, 5, (3, 12, 20), generator=g)
    for flags in [(False, False, False, False), (True, False, False, True), (False, True, True, False), (True, True, True, True)]:
        ref = lab.clone()
        lut = torch.tensor(ops.mask_label_lut(*flags))
        got = ops.mask_label_(lab.clone().cuda(), *flags)
        assert torch.equal(got.cpu(), lut[ref])
    # labelToPred: ones.scatter_(1, label, -1) * -1, viewed [B,H,W,C] and permuted (transform.py:172-183)
    B, H, W, C = 3, 12, 20, 5
    ref = (torch.ones(B * H * W, C).scatter_(1, lab.view(-1, 1), -1.0) * (-1)).view(B, H, W, C).permute(0, 3, 1, 2)
    assert torch.equal(ops.label_to_pred(lab.cuda(), C).cpu(), ref.contiguous())
    # LabelProp assembly of P frame pairs
    P = 4
    ya, yb = torch.randn(P, H, W, generator=g), torch.randn(P, H, W, generator=g)
    la, lb = torch.randint(0, 5, (P, H, W), generator=g), torch.randint(0, 5, (P, H, W), generator=g)
    inp = torch.zeros(2 * P, 8, H, W); tgt = torch.zeros(2 * P, H, W, dtype=torch.long)
    for q in range(P):
        preds = (torch.ones(2 * H * W, C).scatter_(1, torch.stack([la[q], lb[q]]).view(-1, 1), -1.0) * (-1)) \
            .view(2, H, W, C).permute(0, 3, 1, 2)
        inp[2 * q] = torch.cat([ya[q][None], yb[q][None], (ya[q] - yb[q])[None], preds[1]])
        inp[2 * q + 1] = torch.cat([yb[q][None], ya[q][None], (yb[q] - ya[q])[None], preds[0]])
        tgt[2 * q], tgt[2 * q + 1] = la[q], lb[q]
    gi, gt = ops.lp_assemble(ya.cuda(), yb.cuda(), la.cuda(), lb.cuda(), C)
    assert torch.equal(gi.cpu(), inp) and torch.equal(gt.cpu(), tgt)


def test_augment_matches_reference_color_jitter():
    """rcv_augment (Normalize + flip + ColorJitter in one pass, dataset.py:19-39, 123-131) against the output of
    the reference's own ColorJitter (golden) and, on a larger seeded batch, against the oracle restatement."""
    from oracle import ref_transforms as RT
    from robocupvision_b200 import ops
    from util import load_golden
    g = load_golden("augment_dice")
    sc = g["aug_scalars"]
    params = ops.color_jitter_params(g["aug_flip"].tolist(), sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3], "cuda")
    img, lab = ops.augment(torch.from_numpy(g["aug_in"]).cuda(), params, torch.from_numpy(g["aug_labels"]).cuda())
    assert_close("augment vs reference ColorJitter", img, torch.from_numpy(g["aug_out"]), 1e-6)
    assert torch.equal(lab.cpu(), torch.from_numpy(g["aug_labels_out"]))
    gen = torch.Generator().manual_seed(77)
    x = torch.rand(5, 3, 120, 160, generator=gen)
    y = torch.randint(0, 5, (5, 120, 160), generator=gen)
    flip = [True, False, True, True, False]
    b, c = [0.1, -0.2, 0.3, 0.0, -0.05], [1.1, 0.8, 1.25, 1.0, 0.7]
    s, h = [0.9, 1.2, 0.75, 1.0, 1.3], [0.3, -0.4, 0.5, 0.0, -0.1]
    img, lab = ops.augment(x.cuda(), ops.color_jitter_params(flip, b, c, s, h, "cuda"), y.cuda())
    for i in range(5):
        ri, rl = RT.normalize_flip_jitter(x[i], y[i], flip[i], b[i], c[i], s[i], h[i])
        assert_close(f"augment image {i}", img[i], ri, 1e-6)
        assert torch.equal(lab[i].cpu(), rl)
    img2, none = ops.augment(x.cuda(), ops.color_jitter_params(flip, b, c, s, h, "cuda"))
    assert none is None and torch.equal(img2, img)


def test_dice_loss_matches_reference():
    """DiceLoss (model.py:5-43) on the fused kernels: value and logits gradient against the reference's own
    (golden), and against the oracle on a second seeded case with non-unit upstream gradient."""
    from oracle import ref_transforms as RT
    from robocupvision_b200.model import DiceLoss
    from util import load_golden
    g = load_golden("augment_dice")
    logits = torch.from_numpy(g["dice_logits"]).cuda().requires_grad_(True)
    loss = DiceLoss(torch.from_numpy(g["dice_weights"]))(logits, torch.from_numpy(g["dice_true"]).cuda())
    loss.backward()
    assert abs(float(loss) - float(g["dice_loss"])) <= 1e-6
    ref_g = torch.from_numpy(g["dice_grad"])
    assert float((logits.grad.cpu() - ref_g).abs().max()) <= 1e-5 * float(ref_g.abs().max())
    gen = torch.Generator().manual_seed(5)
    z = torch.randn(4, 3, 30, 40, generator=gen) * 2
    t = torch.randint(0, 3, (4, 30, 40), generator=gen)
    w = torch.tensor([1.0, 4.0, 0.5])
    zr = z.clone().requires_grad_(True)
    (RT.dice_loss(zr, t, w) * 2.5).backward()
    zc = z.cuda().requires_grad_(True)
    lc = DiceLoss(w)(zc, t.unsqueeze(1).cuda())   # the [B,1,H,W] target form the docstring names
    (lc * 2.5).backward()
    assert abs(float(lc) - float(RT.dice_loss(z, t, w))) <= 1e-6
    assert float((zc.grad.cpu() - zr.grad).abs().max()) <= 1e-5 * float(zr.grad.abs().max())
