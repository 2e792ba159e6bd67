"""Export wire format of the reference (paramSave.py:5-17): the state_dict flattened in key order
into one raw float64 file that an external engine reads next to a hand-written net.cfg.

Host-side code (the reference calls ``.numpy()`` on CPU tensors); CUDA tensors are copied back first.
``weightsLP/weights.dat`` in the reference tree is exactly this flatten of
``pth/bestModelLPFinetunedPruned.pth`` (tests/golden/weightsLP_head.npz pins head, tail, length, sha256).
"""
from __future__ import annotations

import os
from typing import Mapping

import numpy as np
import torch


def flatten_state_dict(state_dict: Mapping[str, torch.Tensor], skipClassifier: bool = False) -> np.ndarray:
    """float64 concatenation of every entry in key order (paramSave.py:9-16)."""
    parts = [np.empty(0)]
    for name, t in state_dict.items():
        if "classifier" in name and skipClassifier:
            print("Classifier module skipped")
            continue
        a = t.detach().cpu().numpy()
        parts.append(a.reshape(a.size).astype(np.float64, copy=False))
    return np.concatenate(parts)


def saveParams(path, model, fName: str = "weights.dat", skipClassifier: bool = False):
    """Drop-in for paramSave.saveParams (same arguments, same file)."""
    if not os.path.exists(path):
        os.makedirs(path)
    flatten_state_dict(model.state_dict(), skipClassifier).tofile(os.path.join(path, fName))


# ------------------------------------------------------------------------------------------------ net.cfg
# The layer list the external inference engine reads next to weights.dat (weights/net.cfg, weightsLP/net.cfg,
# weightsVGA/net.cfg in the reference tree: hand-written, darknet-style).  Here it is emitted from the module's own
# execution plan, so it cannot drift from the net that produced weights.dat.
def net_cfg_sections(model, height=None, width=None, channels=None, downscale=4):
    """-> [(section, [(key, value), ...]), ...] in file order.  Sections as in the reference's files:
    net, convolutional, batchnorm, transposedconv, shortcut (from = absolute index of the layer whose output is
    added), softmax; plus, for nets the reference ships no cfg for, maxpool (2x2, --UNet) and route (layers = the
    concatenated layers, --v2)."""
    from .ops import EPI_AFFINE, EPI_AFFINE_RELU, EPI_NONE, EPI_RELU, EPI_RELU_AFFINE
    plan = model._get_plan()
    first = next(nd for nd in plan.nodes if nd.kind == "conv")
    h, w = getattr(model, "img_shape", (120, 160))
    out = [("net", [("height", height or h), ("width", width or w), ("channels", channels or first.geom.cin),
                    ("downscale", downscale)])]
    layer_of = {}  # activation index -> index of the layer (section after [net]) that produced it

    def emit(section, kv, act):
        out.append((section, kv))
        layer_of[act] = len(out) - 2

    for t, nd in enumerate(plan.nodes):
        act = t + 1
        if nd.kind == "pool":
            emit("maxpool", [("size", 2), ("stride", 2)], act)
            continue
        g, conv = nd.geom, nd.conv
        conv_act = "relu" if nd.order in (EPI_RELU, EPI_RELU_AFFINE) else "linear"
        if g.transposed:
            emit("transposedconv", [("filters", g.cout), ("size", g.k), ("stride", g.stride), ("pad", g.pad),
                                    ("outpad", 1), ("activation", conv_act)], act)
        else:
            kv = [("filters", g.cout), ("size", g.k), ("stride", g.stride), ("pad", g.pad)]
            if conv.bias is None or g.dil != 1:
                kv.append(("dilation", g.dil))
            kv.append(("activation", conv_act))
            if conv.bias is None:
                kv.append(("hasBias", 0))
            emit("convolutional", kv, act)
        if nd.bn is not None:
            emit("batchnorm", [("activation", "relu" if nd.order == EPI_AFFINE_RELU else "linear")], act)
        elif nd.order not in (EPI_NONE, EPI_RELU):
            raise ValueError(f"net_cfg: node {t} has epilogue {nd.order} without a BatchNorm module")
        if nd.skip >= 0:
            if nd.skip_mode == "cat":
                emit("route", [("layers", f"{layer_of[act]},{layer_of[nd.skip]}")], act)
            else:  # "add", and LabelProp's add into the first channels (the shortcut adds the common channels)
                emit("shortcut", [("activation", "linear"), ("from", layer_of[nd.skip])], act)
    out.append(("softmax", []))
    return out


def net_cfg(model, **kw) -> str:
    """The cfg text (same keys, order and spelling as the reference's files)."""
    lines = []
    for section, kv in net_cfg_sections(model, **kw):
        lines.append(f"[{section}]")
        for k, v in kv:
            lines.append(f"{k} = {v}" if section == "batchnorm" else f"{k}={v}")
        lines.append("")
    return "\n".join(lines)


def parse_net_cfg(text: str):
    """Inverse of net_cfg, whitespace-insensitive: -> [(section, [(key, value-as-str), ...]), ...]."""
    out = []
    for raw in text.splitlines():
        ln = raw.strip()
        if not ln or ln.startswith("#"):
            continue
        if ln.startswith("["):
            out.append((ln.strip("[]").strip(), []))
        else:
            k, v = ln.split("=", 1)
            out[-1][1].append((k.strip(), v.strip()))
    return out


def saveNetCfg(path, model, fName: str = "net.cfg", **kw):
    """Write net.cfg next to weights.dat (the reference's weights*/ directories hold the pair)."""
    if not os.path.exists(path):
        os.makedirs(path)
    with open(os.path.join(path, fName), "w") as f:
        f.write(net_cfg(model, **kw))
