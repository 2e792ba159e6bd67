"""Test infrastructure: ATen-CPU stand-ins for robocupvision_b200.ops, written from the documented contracts of the
entry points (include/rcv_b200.h).  With engine.ops replaced by this module the REAL scheduling code of
engine.Plan (forward incl. the deferred-BatchNorm / normalise-on-load schedule, backward incl. skip routing,
multi-consumer gradient accumulation, statistic arenas) runs on CPU tensors, so its results can be compared with
autograd over the oracle without a GPU.  Never imported by the package."""
import torch
import torch.nn.functional as F

from robocupvision_b200 import ops as real
from robocupvision_b200.ops import *  # noqa: F401,F403  (constants: EPI_*, MATH_*, PACK_*)
from robocupvision_b200.ops import (EPI_AFFINE, EPI_AFFINE_RELU, EPI_NONE, EPI_RELU, EPI_RELU_AFFINE)

NOMINAL_NHW = real.NOMINAL_NHW
calls = []  # (name, detail) log, for assertions about the schedule


def _chk(t, dtype=torch.float32, name="tensor"):
    assert t.dtype == dtype, name
    return t.contiguous()


def conv_uses_tensor_cores(g, direction, math=real.MATH_AUTO):
    return False  # no weight panels on the CPU


conv_normalises_on_load = real.conv_normalises_on_load              # host logic of the library: real answers
conv_wgrad_normalises_on_load = real.conv_wgrad_normalises_on_load


def _add_res(v, residual):
    """+ residual; a residual with fewer channels goes to the first channels only (rcv_*::res_channels)."""
    if residual.shape[1] == v.shape[1]:
        return v + residual
    v = v.clone()
    v[:, :residual.shape[1]] += residual
    return v


def conv_takes_partial_residual(g, n, h, w, math=real.MATH_AUTO):
    return True


def _affine_in(x, in_affine):
    if in_affine is None:
        return x
    sc, sh, relu = in_affine
    t = x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    return F.relu(t) if relu else t


def _conv(g, x, w, b):
    if g.transposed:
        return F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    return F.conv2d(x, w, b, g.stride, g.pad, g.dil)


def conv_workspace_bytes(g, n, h, w, direction=0, math=real.MATH_AUTO):
    return 0


def conv_fwd(g, x, w, bias=None, epilogue=EPI_NONE, scale=None, shift=None, residual=None, stats=None,
             math=real.MATH_FP32, out=None, wpacked=None, in_affine=None, workspace=None):
    calls.append(("conv_fwd", in_affine is not None))
    if in_affine is not None:
        n, _, h, wd = x.shape
        assert real.conv_normalises_on_load(g, n, h, wd, math), "engine asked an engine that refuses in_affine"
    v = _conv(g, _affine_in(x, in_affine), w, bias)
    A = scale.view(1, -1, 1, 1) if scale is not None else None
    B = shift.view(1, -1, 1, 1) if shift is not None else None
    if epilogue == EPI_RELU:
        v = F.relu(v)
    elif epilogue == EPI_RELU_AFFINE:
        v = A * F.relu(v) + B
    elif epilogue == EPI_AFFINE_RELU:
        v = F.relu(A * v + B)
    elif epilogue == EPI_AFFINE:
        v = A * v + B
    if residual is not None:
        v = _add_res(v, residual)
    if stats is not None:
        c = v.shape[1]
        d = v.double()
        stats[:c] += d.sum((0, 2, 3))
        stats[c:] += (d * d).sum((0, 2, 3))
    return v


def conv_dgrad(g, dy, w, in_hw, residual=None, math=real.MATH_FP32, out=None, wpacked=None, workspace=None):
    n = dy.shape[0]
    x = torch.zeros(n, g.cin, *in_hw, requires_grad=True)
    with torch.enable_grad():
        y = _conv(g, x, w, None)
    (dx,) = torch.autograd.grad(y, x, dy)
    return dx if residual is None else dx + residual


def conv_wgrad(g, x, dy, dw=None, dbias=None, want_bias=False, math=real.MATH_FP32, in_affine=None):
    calls.append(("conv_wgrad", in_affine is not None))
    if in_affine is not None:
        n, _, h, wd = x.shape
        assert real.conv_wgrad_normalises_on_load(g, n, h, wd, math)
    wz = torch.zeros(g.weight_shape(), requires_grad=True)
    with torch.enable_grad():
        y = _conv(g, _affine_in(x, in_affine), wz, None)
    (gw,) = torch.autograd.grad(y, wz, dy)
    if dw is None:
        dw = torch.zeros(g.weight_shape())
    dw += gw
    if dbias is None and want_bias:
        dbias = torch.zeros(g.cout)
    if dbias is not None:
        dbias += dy.sum((0, 2, 3))
    return dw, dbias


def zeros(shape, dtype, device):
    return torch.zeros(shape, dtype=dtype, device=device)


def zero_(t):
    return t.zero_()


def bn_finalize(stats, count, gamma, beta, running_mean, running_var, momentum, eps, num_batches_tracked=None):
    calls.append(("bn_finalize", None))
    if num_batches_tracked is not None:
        num_batches_tracked += 1
    c = stats.numel() // 2
    mean = stats[:c] / count
    var = (stats[c:] / count - mean * mean).clamp_min(0)
    invstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.double() * invstd
    shift = beta.double() - mean * scale
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean.float())
        running_var.mul_(1 - momentum).add_(momentum * (var * count / max(count - 1, 1)).float())
    return scale.float(), shift.float(), mean.float(), invstd.float()


def bn_apply(z, scale, shift, relu, residual=None, out=None):
    calls.append(("bn_apply", None))
    y = z * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if relu:
        y = F.relu(y)
    return y if residual is None else _add_res(y, residual)


def bn_finalize_apply(z, stats, gamma, beta, running_mean, running_var, momentum, eps, relu, residual=None,
                      num_batches_tracked=None):
    calls.append(("bn_finalize_apply", None))
    count = z.numel() // z.shape[1]
    scale, shift, mean, invstd = bn_finalize(stats, count, gamma, beta, running_mean, running_var, momentum, eps,
                                             num_batches_tracked)
    calls.pop()
    y = bn_apply(z, scale, shift, relu, residual)
    calls.pop()
    return y, scale, shift, mean, invstd


def bn_fold(gamma, beta, mean, var, eps):
    inv = 1.0 / torch.sqrt(var + eps)
    return gamma * inv, beta - mean * gamma * inv


def bn_bwd(order, dy, z, scale, shift, mean, invstd, dgamma=None, dbeta=None, dbias=None, want_dbias=False, sums=None):
    """Gradient of  y = [relu](gamma * (z - mean) * invstd + beta)  with batch statistics, z = [relu](conv) for
    EPI_RELU_AFFINE: -> gradient with respect to the conv output (through that ReLU as well)."""
    g = dy
    if order == EPI_AFFINE_RELU:
        g = torch.where(z * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) > 0, dy, torch.zeros_like(dy))
    m = z.numel() // z.shape[1]
    xhat = (z - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    sum_g = g.double().sum((0, 2, 3))
    sum_gx = (g.double() * xhat.double()).sum((0, 2, 3))
    gamma = (scale / invstd).view(1, -1, 1, 1)
    dz = gamma * invstd.view(1, -1, 1, 1) * (g - (sum_g / m).float().view(1, -1, 1, 1)
                                             - xhat * (sum_gx / m).float().view(1, -1, 1, 1))
    if order == EPI_RELU_AFFINE:
        dz = torch.where(z > 0, dz, torch.zeros_like(dz))
    if dgamma is not None:
        dgamma += sum_gx.float()
    if dbeta is not None:
        dbeta += sum_g.float()
    if dbias is not None:
        dbias += dz.sum((0, 2, 3))
    return dz, dgamma, dbeta, dbias


def head_ce_supported(cin, classes):
    return cin in (8, 16) and 2 <= classes <= 8


def ce_weight_sum(target, class_w, out, classes):
    w = class_w if class_w is not None else torch.ones(classes)
    ok = (target >= 0) & (target < classes)
    out += w.double()[target.clamp(0, classes - 1)][ok].sum()


def head_ce_train(feat, weight, bias, target, class_w, sums, corr, dweight, dbias):
    """ATen stand-in of rcv_head_ce_train: autograd over conv1x1 + weighted NLL (sum form) / sums[1]."""
    calls.append(("head_ce_train",))
    f = feat.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    b = None if bias is None else bias.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        z = F.conv2d(f, w.view(w.shape[0], -1, 1, 1), b)
        nll = F.cross_entropy(z, target, weight=class_w, reduction="sum")
        (nll / float(sums[1])).backward()
    sums[0] += nll.detach().double()
    corr += (z.argmax(1) == target).sum()
    dweight += w.grad.view_as(dweight)
    if dbias is not None:
        dbias += b.grad
    return f.grad


def channel_slice(x, offset, count):
    return x[:, offset:offset + count].contiguous()


def concat_channels(a, b):
    return torch.cat([a, b], 1)


def relu_bwd(dy, y):
    return torch.where(y > 0, dy, torch.zeros_like(dy))


def maxpool2x2_fwd(x, want_idx=False, want_code=True):
    y, idx = F.max_pool2d(x, 2, 2, return_indices=True)
    return y, (idx if want_idx else None), idx  # the "code" of the stand-in is the flat index itself


def maxpool2x2_bwd(dy, code, in_hw):
    return F.max_unpool2d(dy, code, 2, 2, output_size=in_hw)
