"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Restatement of one training step of train.py:43-74 (ROBO_UNet) on the functional oracle:
zero_grad -> forward -> CrossEntropyLoss2d -> + decay * l1reg -> backward -> [grad mask] ->
Adam.step -> argmax / correct-pixel count.  The optimiser is torch.optim.Adam itself (the
reference's own call, train.py:357-363), fed the oracle's leaf tensors.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ref_model as R


class OracleTrainer:
    def __init__(self, sd: Dict[str, torch.Tensor], forward: Callable, class_weights=None, lr=1e-3,
                 l1_decay=1e-6, masks: Optional[List[torch.Tensor]] = None, optimizer="adam",
                 momentum=0.0, weight_decay=0.0):
        self.sd = R.leaf_state_dict(sd)
        self.forward = forward
        self.keys = R.param_keys(self.sd)
        self.params = [self.sd[k] for k in self.keys]
        self.w = None if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32)
        self.l1 = l1_decay
        self.masks = masks
        if optimizer == "adam":
            self.opt = torch.optim.Adam(self.params, lr=lr)
        else:
            self.opt = torch.optim.SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)

    def step(self, x: torch.Tensor, y: torch.Tensor):
        self.opt.zero_grad()
        pred = self.forward(self.sd, x, training=True)
        loss = R.cross_entropy_2d(pred, y, self.w)
        reg = torch.zeros(())
        if self.masks is None and self.l1:
            reg = self.l1 * R.l1reg(self.params)
            loss = loss + reg
        loss.backward()
        if self.masks is not None:
            i = 0
            for p in self.params:
                if p.dim() > 1:
                    if p.grad is not None:
                        p.grad[self.masks[i]] = 0
                    i += 1
        grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in zip(self.keys, self.params)}
        self.opt.step()
        pred_class = pred.detach().argmax(1)
        correct = int((pred_class == y).sum())
        return float(loss.detach()), float(reg), correct, pred.detach(), grads
