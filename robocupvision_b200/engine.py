"""Execution plan for the encoder-decoder nets: a flat list of fused nodes built from the
``model.py``-compatible module tree, with a hand-scheduled forward and backward.

Each conv-type node is  conv [+bias] -> (ReLU, BatchNorm) in the block's order -> [+ skip]
and runs as:
  eval : 1 kernel  (implicit-GEMM conv with bias/ReLU/folded-BN/skip in the epilogue)
  train: conv (+bias, +ReLU for the `Conv` order, + per-channel sum / sum^2 in the epilogue)
         -> bn_finalize (tiny) -> bn_apply (scale/shift [+ReLU] [+skip])
  bwd  : bn_bwd_reduce -> bn_bwd_apply (-> dconv, dgamma, dbeta, dbias) -> dgrad (+ skip
         gradient summed in its epilogue) -> wgrad (split over pixels, RED atomics)
The autograd boundary is one ``torch.autograd.Function`` for the whole plan; parameter
gradients are views of one zero-filled flat buffer (or of the caller's arena).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .ops import (EPI_AFFINE_RELU, EPI_NONE, EPI_RELU, EPI_RELU_AFFINE, MATH_AUTO, MATH_FP32, PACK_DGRAD,
                  PACK_FWD, ConvGeom)

# Math mode of the conv engine for every plan: RCV_MATH_AUTO = tcgen05 3xTF32 tensor-core tiles
# wherever the reduction is long enough, CUDA cores elsewhere.  RCV_B200_MATH=fp32 forces the
# CUDA-core engine everywhere (A/B comparisons, bisecting a numerical difference).
import os as _os

DEFAULT_MATH = {"fp32": MATH_FP32, "auto": MATH_AUTO}[_os.environ.get("RCV_B200_MATH", "auto").lower()]


class Node:
    __slots__ = ("kind", "src", "conv", "bn", "order", "skip", "skip_mode", "geom", "_fold_key",
                 "_fold_val", "skip_ch", "_pack", "_pack_key", "_tc")

    def __init__(self, kind, src, conv=None, bn=None, order=EPI_NONE, skip=-1, skip_mode="add"):
        self.kind, self.src, self.conv, self.bn = kind, src, conv, bn
        self.order, self.skip, self.skip_mode = order, skip, skip_mode
        self.geom = ConvGeom.of(conv) if conv is not None else None
        self._fold_key = None
        self._fold_val = None
        self.skip_ch = 0
        self._pack = [None, None]      # persistent packed-weight buffers (fwd, dgrad)
        self._pack_key = [None, None]
        self._tc = [None, None]        # does this direction run on tensor cores

    def params(self) -> List[nn.Parameter]:
        out = []
        if self.conv is not None:
            out.append(self.conv.weight)
            if self.conv.bias is not None:
                out.append(self.conv.bias)
        if self.bn is not None:
            out += [self.bn.weight, self.bn.bias]
        return out

    def packed(self, direction: int, epoch: int, fresh: bool, math: int):
        """Weight panel of the tensor-core engine (None if this layer/direction stays on CUDA
        cores).  Re-packed when `fresh` (training: weights change every step) or when the weight
        tensor / the plan's epoch changed; the buffer itself is allocated once (CUDA-graph safe)."""
        if math == MATH_FP32:
            return None
        if self._tc[direction] is None:
            self._tc[direction] = ops.conv_uses_tensor_cores(self.geom, direction, math)
        if not self._tc[direction]:
            return None
        w = self.conv.weight
        key = (w.data_ptr(), w._version, epoch)
        buf = self._pack[direction]
        if buf is not None and buf.device != w.device:
            buf = None
        if buf is None or fresh or key != self._pack_key[direction]:
            buf = ops.conv_pack(self.geom, w.detach(), direction, out=buf)
            self._pack[direction] = buf
            self._pack_key[direction] = key
        return buf

    def folded(self, epoch: int = 0):
        bn = self.bn
        ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in ts) + (epoch,)
        if key != self._fold_key:
            self._fold_val = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                         bn.running_var, bn.eps)
            self._fold_key = key
        return self._fold_val


class PlanBuilder:
    """acts[0] is the plan input; node t produces acts[t+1]."""

    def __init__(self):
        self.nodes: List[Node] = []

    def conv(self, src: int, conv: nn.Module, bn: Optional[nn.Module], order: int, skip: int = -1,
             skip_mode: str = "add", skip_ch: int = 0) -> int:
        if bn is None and order not in (EPI_NONE, EPI_RELU):
            raise ValueError("conv without BatchNorm takes EPI_NONE or EPI_RELU")
        if bn is not None and order not in (EPI_RELU_AFFINE, EPI_AFFINE_RELU):
            raise ValueError("conv with BatchNorm takes EPI_RELU_AFFINE or EPI_AFFINE_RELU")
        nd = Node("conv", src, conv, bn, order, skip, skip_mode)
        nd.skip_ch = skip_ch
        self.nodes.append(nd)
        return len(self.nodes)

    def pool(self, src: int) -> int:
        self.nodes.append(Node("pool", src))
        return len(self.nodes)


class Plan:
    def __init__(self, builder: PlanBuilder, outputs: Sequence[int]):
        self.nodes = builder.nodes
        self.outputs = list(outputs)
        self.params: List[nn.Parameter] = []
        seen = set()
        for nd in self.nodes:
            for p in nd.params():
                if id(p) not in seen:
                    seen.add(id(p))
                    self.params.append(p)
        self.math = DEFAULT_MATH
        # bumped whenever parameters / BN buffers may have been written behind torch's back (raw
        # kernels of a training forward or of TrainStep): invalidates folded-BN and packed caches
        self.epoch = 0
        self.n_stats = 0
        self._sum_off = {}
        for t, nd in enumerate(self.nodes):
            if nd.bn is not None:
                self._sum_off[t] = self.n_stats
                self.n_stats += 2 * nd.geom.cout

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, training: bool, save: bool):
        """-> (outputs, saved).  training selects batch statistics for BatchNorm nodes whose
        module is in training mode; save keeps what backward needs."""
        x = ops._chk(x, name="input")
        if x.dim() != 4:
            raise ValueError(f"expected NCHW input, got shape {tuple(x.shape)}")
        dev = x.device
        acts: List[torch.Tensor] = [x]
        saved: List[Optional[tuple]] = [None] * len(self.nodes)
        stats_arena = None
        soff = 0
        nbt = []
        if training:
            self.epoch += 1
        for t, nd in enumerate(self.nodes):
            src = acts[nd.src]
            if nd.kind == "pool":
                y, _, code = ops.maxpool2x2_fwd(src, want_idx=False, want_code=save)
                saved[t] = (code,)
                acts.append(y)
                continue
            g, conv, bn = nd.geom, nd.conv, nd.bn
            w = conv.weight.detach()
            b = conv.bias.detach() if conv.bias is not None else None
            skip = acts[nd.skip] if (nd.skip >= 0 and nd.skip_mode == "add") else None
            wp = nd.packed(PACK_FWD, self.epoch, training, self.math)
            if save and (nd.src != 0 or x.requires_grad):
                nd.packed(PACK_DGRAD, self.epoch, training, self.math)
            if bn is None:
                y = ops.conv_fwd(g, src, w, b, epilogue=nd.order, math=self.math, wpacked=wp)
                saved[t] = (y if nd.order == EPI_RELU else None,)
            elif training and bn.training:
                if stats_arena is None:
                    stats_arena = torch.zeros(self.n_stats, device=dev, dtype=torch.float64)
                stats = stats_arena[soff:soff + 2 * g.cout]
                soff += 2 * g.cout
                z = ops.conv_fwd(g, src, w, b, epilogue=EPI_RELU if nd.order == EPI_RELU_AFFINE else EPI_NONE,
                                 stats=stats, math=self.math, wpacked=wp)
                count = z.numel() // g.cout
                if bn.momentum is None:
                    momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)
                else:
                    momentum = bn.momentum
                track = bn.track_running_stats and bn.running_mean is not None
                scale, shift, mean, invstd = ops.bn_finalize(
                    stats, count, bn.weight.detach(), bn.bias.detach(),
                    bn.running_mean if track else None, bn.running_var if track else None, momentum, bn.eps)
                y = ops.bn_apply(z, scale, shift, relu=(nd.order == EPI_AFFINE_RELU), residual=skip)
                saved[t] = (z, scale, shift, mean, invstd)
                if track and bn.num_batches_tracked is not None:
                    nbt.append(bn.num_batches_tracked)
            else:
                scale, shift = nd.folded(self.epoch)
                y = ops.conv_fwd(g, src, w, b, epilogue=nd.order, scale=scale, shift=shift, residual=skip,
                                 math=self.math, wpacked=wp)
            if nd.skip >= 0 and nd.skip_mode == "partial":
                y[:, :nd.skip_ch] += acts[nd.skip]
            elif nd.skip >= 0 and nd.skip_mode == "cat":
                y = torch.cat([y, acts[nd.skip]], 1)
            acts.append(y)
        if nbt:
            torch._foreach_add_(nbt, 1)
        outs = [acts[i] for i in self.outputs]
        return outs, ((acts, saved) if save else None)

    # ------------------------------------------------------------------ backward
    def backward(self, saved_all, gouts: Sequence[Optional[torch.Tensor]], x_needs_grad: bool,
                 grad_views: Optional[Dict[int, torch.Tensor]] = None, node_done=None):
        """-> (dx or None, {id(param): grad}).  grad_views, if given, maps id(param) to zero-filled
        tensors that receive the gradients (the train step's flat arena).  node_done(t) is called
        once node t's parameter gradients are final (nodes are visited last to first)."""
        acts, saved = saved_all
        dev = acts[0].device
        if grad_views is None:
            total = sum(p.numel() for p in self.params)
            flat = torch.zeros(total, device=dev, dtype=torch.float32)
            grad_views, o = {}, 0
            for p in self.params:
                grad_views[id(p)] = flat[o:o + p.numel()].view(p.shape)
                o += p.numel()
        grads: List[Optional[torch.Tensor]] = [None] * len(acts)

        def add_to(i, g):
            grads[i] = g if grads[i] is None else grads[i] + g

        for oi, g in zip(self.outputs, gouts):
            if g is not None:
                add_to(oi, g)
        sums_arena = torch.zeros(max(self.n_stats, 1), device=dev, dtype=torch.float64)
        for t in range(len(self.nodes) - 1, -1, -1):
            self._backward_node(t, acts, saved, grads, grad_views, sums_arena, x_needs_grad, add_to)
            if node_done is not None:
                node_done(t)
        return grads[0], grad_views

    def _backward_node(self, t, acts, saved, grads, grad_views, sums_arena, x_needs_grad, add_to):
        if True:
            nd = self.nodes[t]
            g = grads[t + 1]
            grads[t + 1] = None
            if nd.bn is not None:
                soff = self._sum_off[t]
                sums = sums_arena[soff:soff + 2 * nd.geom.cout]
            if g is None:
                return
            g = ops._chk(g, name="grad")
            src = acts[nd.src]
            in_hw = (src.shape[2], src.shape[3])
            if nd.kind == "pool":
                add_to(nd.src, ops.maxpool2x2_bwd(g, saved[t][0], in_hw))
                return
            geom, conv, bn = nd.geom, nd.conv, nd.bn
            if nd.skip >= 0:
                if nd.skip_mode == "add":
                    add_to(nd.skip, g)
                elif nd.skip_mode == "partial":
                    add_to(nd.skip, g[:, :nd.skip_ch].contiguous())
                else:  # cat
                    add_to(nd.skip, g[:, geom.cout:].contiguous())
                    g = g[:, :geom.cout].contiguous()
            w = conv.weight.detach()
            has_bias = conv.bias is not None
            if bn is not None:
                if saved[t] is None or len(saved[t]) != 5:
                    raise NotImplementedError(
                        "backward through an eval-mode BatchNorm block is not on the hot path "
                        "(call model.train() before the forward pass you differentiate)")
                z, scale, shift, mean, invstd = saved[t]
                dconv, _, _, _ = ops.bn_bwd(nd.order, g, z, scale, shift, mean, invstd,
                                            dgamma=grad_views[id(bn.weight)], dbeta=grad_views[id(bn.bias)],
                                            dbias=grad_views[id(conv.bias)] if has_bias else None, sums=sums)
                wg_bias = None
            else:
                dconv = ops.relu_bwd(g, saved[t][0]) if nd.order == EPI_RELU else g
                wg_bias = grad_views[id(conv.bias)] if has_bias else None
            if nd.src != 0 or x_needs_grad:
                wp = nd._pack[PACK_DGRAD] if (self.math != MATH_FP32 and nd._tc[PACK_DGRAD]) else None
                grads[nd.src] = ops.conv_dgrad(geom, dconv, w, in_hw, residual=grads[nd.src], math=self.math,
                                               wpacked=wp)
            ops.conv_wgrad(geom, src, dconv, dw=grad_views[id(conv.weight)], dbias=wg_bias, math=self.math)


class _PlanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: Plan, training: bool, need: bool, x: torch.Tensor, *params):
        # (grad mode is always off inside Function.forward, so `need` is decided by the caller)
        outs, saved = plan.forward(x, training, save=need and training)
        ctx.plan, ctx.saved, ctx.training = plan, saved, training
        ctx.x_needs_grad = x.requires_grad
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.saved is None:
            raise NotImplementedError(
                "robocupvision_b200: backward needs a training-mode forward (model.train()); "
                "eval-mode forwards keep no activations")
        dx, gv = ctx.plan.backward(ctx.saved, gouts, ctx.x_needs_grad)
        ctx.saved = None
        return (None, None, None, dx, *[gv[id(p)] for p in ctx.plan.params])


def run_plan(plan: Plan, x: torch.Tensor, training: bool) -> Tuple[torch.Tensor, ...]:
    if not x.is_cuda:
        raise RuntimeError(
            "robocupvision_b200 runs on CUDA (sm_100a) only: move the model and its input to the GPU. "
            "There is no CPU fallback; the CPU oracle under oracle/ is test infrastructure.")
    if x.dtype != torch.float32:
        raise TypeError(f"robocupvision_b200: fp32 input expected, got {x.dtype}")
    need = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in plan.params))
    return _PlanFn.apply(plan, training, need, x, *plan.params)
