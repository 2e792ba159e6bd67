"""Test infrastructure: run an engine.Plan node by node with ATen CPU ops.

The plan is the host-side product of the drop-in modules (which conv feeds which, epilogue order, skip wiring);
the kernels that execute its nodes are pinned per op on the GPU.  Interpreting the same node list on the CPU
checks the wiring of every model family against the oracle without a GPU (tests/test_plan_host.py)."""
import torch
import torch.nn.functional as F

from robocupvision_b200.ops import EPI_AFFINE, EPI_AFFINE_RELU, EPI_NONE, EPI_RELU, EPI_RELU_AFFINE


def run_plan_cpu(plan, x, training=False):
    """Forward of `plan` (eval-mode BatchNorm unless training) -> list of output tensors."""
    acts = [x]
    for nd in plan.nodes:
        src = acts[nd.src]
        if nd.kind == "pool":
            acts.append(F.max_pool2d(src, 2, 2))
            continue
        g, conv, bn = nd.geom, nd.conv, nd.bn
        if g.transposed:
            y = F.conv_transpose2d(src, conv.weight, conv.bias, stride=2, padding=1, output_padding=1)
        else:
            y = F.conv2d(src, conv.weight, conv.bias, g.stride, g.pad, g.dil)

        def norm(t):
            return F.batch_norm(t, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)

        if nd.order == EPI_RELU:
            y = F.relu(y)
        elif nd.order == EPI_RELU_AFFINE:
            y = norm(F.relu(y))
        elif nd.order == EPI_AFFINE_RELU:
            y = F.relu(norm(y))
        elif nd.order == EPI_AFFINE:
            y = norm(y)
        else:
            assert nd.order == EPI_NONE and bn is None
        if nd.skip >= 0:
            s = acts[nd.skip]
            if nd.skip_mode == "add":
                y = y + s
            elif nd.skip_mode == "partial":
                y = torch.cat([y[:, :nd.skip_ch] + s, y[:, nd.skip_ch:]], 1)
            else:
                assert nd.skip_mode == "cat"
                y = torch.cat([y, s], 1)
        acts.append(y)
    return [acts[o] for o in plan.outputs]
