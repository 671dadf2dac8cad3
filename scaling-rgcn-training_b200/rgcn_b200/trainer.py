"""One training step of the reference's ``Trainer.train`` loop body
(/root/reference/model/modelTrainer.py:61-69) and its two losses (model/evaluation.py:33-42),
mirrored so bench.py and the tests can drive the engine the way the reference does."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
from torch import Tensor, nn


def ce_loss(pred: Tensor, targets: Tensor) -> Tensor:
    return nn.functional.cross_entropy(pred, targets.argmax(-1))


def bce_loss(pred: Tensor, targets: Tensor) -> Tensor:
    return nn.functional.binary_cross_entropy(pred, targets)


def identity(x: Tensor) -> Tensor:
    return x


def get_losst(dataset: str, sumModel: bool = False) -> Tuple[Callable, Callable]:
    """(loss, activation) selector, evaluation.py:44-48."""
    if sumModel or dataset == 'AIFB':
        return bce_loss, torch.sigmoid
    return ce_loss, identity


def make_optimizer(model: nn.Module, lr: float = 0.01, weight_decay: float = 5e-5, fused: bool = True):
    """Adam as modelTrainer.py:44 builds it (weight_d = 5e-5 from main.py:52): the engine's one-pass
    FusedAdam (same update rule) by default, torch.optim.Adam with fused=False."""
    if fused:
        from .optim import FusedAdam
        return FusedAdam(model.parameters(), lr=lr, weight_decay=weight_decay)
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_step(model: nn.Module, training_data, optimizer, loss_f: Callable, activation: Callable) -> float:
    model.train()
    optimizer.zero_grad()
    out = model(training_data, activation)
    targets = training_data.y_train.to(torch.float32)
    loss = loss_f(out[training_data.x_train], targets)
    loss.backward()
    optimizer.step()
    return loss.item()
