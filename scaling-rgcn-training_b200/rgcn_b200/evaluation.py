"""Device-side mirror of the reference's ``evaluate`` (/root/reference/model/evaluation.py:14-31;
SURVEY.md section 8f rank 1).  The reference builds ``torch.zeros(pred.shape)`` on the host,
scatters with a device index and hands device tensors to sklearn — it only runs on CPU
(SURVEY F4).  Here the predictions and the per-class confusion counts are computed on the GPU
(``rgcn_eval_counts``) and only 3C+1 integers reach the host; the metrics are the ones sklearn
computes for multilabel-indicator input: subset accuracy, F1 weighted by support, macro F1
(``zero_division=0``).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib


def confusion_counts(pred: Tensor, x: Tensor, y: Tensor, sigmoid_path: bool) -> np.ndarray:
    """-> int64 [3C+1]: tp[C], fp[C], fn[C], exact-match rows."""
    lib = _lib.load()
    if not pred.is_cuda:
        raise _lib.EngineError('confusion_counts: CUDA tensors required (no CPU path)')
    pred = pred.detach()
    pred = pred if pred.stride(1) == 1 else pred.contiguous()
    dev = pred.device
    c = pred.size(1)
    x = x.to(device=dev, dtype=torch.int64).contiguous()
    y = y.to(device=dev, dtype=torch.int64).contiguous()
    if y.shape != (x.numel(), c):
        raise ValueError('confusion_counts: y must be [len(x), num_classes]')
    counts = torch.empty(3 * c + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.rgcn_eval_counts(pred.data_ptr(), pred.stride(0), c, x.data_ptr(), x.numel(), y.data_ptr(),
                                  1 if sigmoid_path else 0, counts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, 'rgcn_eval_counts')
    return counts.cpu().numpy()


def metrics_from_counts(counts: np.ndarray, n_rows: int) -> Tuple[float, float, float]:
    """(accuracy, f1 weighted, f1 macro) exactly as sklearn defines them on indicator matrices."""
    c = (counts.size - 1) // 3
    tp, fp, fn = (counts[i * c:(i + 1) * c].astype(np.float64) for i in range(3))
    denom = 2 * tp + fp + fn
    f1 = np.divide(2 * tp, denom, out=np.zeros(c), where=denom > 0)
    support = tp + fn
    acc = float(counts[3 * c]) / n_rows if n_rows else 0.0
    f1_w = float((f1 * support).sum() / support.sum()) if support.sum() > 0 else 0.0
    f1_m = float(f1.mean()) if c else 0.0
    return acc, f1_w, f1_m


def evaluate(model: nn.Module, activation: Callable, training_data, x: Tensor, y: Tensor,
             report: bool = False) -> Tuple[float, float, float]:
    """Same signature and return value as the reference's evaluate."""
    with torch.no_grad():
        pred = model(training_data, activation)
    counts = confusion_counts(pred, x, y, sigmoid_path=activation is torch.sigmoid)
    acc, f1_w, f1_m = metrics_from_counts(counts, x.numel())
    if report:
        c = (counts.size - 1) // 3
        print('class  tp  fp  fn')
        for k in range(c):
            print(k, int(counts[k]), int(counts[c + k]), int(counts[2 * c + k]))
    return acc, f1_w, f1_m
