"""Binary-change metrics with the reference's names (utils/metric_tool.py) and a device-resident confusion matrix.

The reference updates `ConfuseMatrixMeter` every training step from `pred.cpu().numpy()` / `target.cpu().numpy()`
(scripts/train_BCD.py:203-225: a 16.8 MB int64 device->host copy + np.bincount per step, which serialises the GPU).
Here the matrix stays on the device — accumulated by the loss kernel itself (losses.bce_dice_loss(cm=...)) or by
`c3d_confusion_matrix` — and is read back once, when scores are asked for.  Score formulas are the reference's
(cm2F1 :68-81, cm2score :84-108), evaluated in float64 on the 2x2 matrix.
"""
from __future__ import annotations

import numpy as np
import torch

from . import losses

_EPS = np.finfo(np.float32).eps


def cm2F1(confusion_matrix) -> float:
    """utils/metric_tool.py:68-81."""
    hist = np.asarray(confusion_matrix, dtype=np.float64)
    tp, fn, fp = hist[1, 1], hist[1, 0], hist[0, 1]
    recall = tp / (tp + fn + _EPS)
    precision = tp / (tp + fp + _EPS)
    return 2 * recall * precision / (recall + precision + _EPS)


def cm2score(confusion_matrix) -> dict:
    """utils/metric_tool.py:84-108."""
    hist = np.asarray(confusion_matrix, dtype=np.float64)
    tp, fn, fp, tn = hist[1, 1], hist[1, 0], hist[0, 1], hist[0, 0]
    oa = (tp + tn) / (tp + fn + fp + tn + _EPS)
    recall = tp / (tp + fn + _EPS)
    precision = tp / (tp + fp + _EPS)
    f1 = 2 * recall * precision / (recall + precision + _EPS)
    iou = tp / (tp + fp + fn + _EPS)
    pre = ((tp + fn) * (tp + fp) + (tn + fp) * (tn + fn)) / (tp + fp + tn + fn) ** 2
    kappa = (oa - pre) / (1 - pre)
    return {'Kappa': kappa, 'IoU': iou, 'F1': f1, 'OA': oa, 'recall': recall, 'precision': precision, 'Pre': pre}


class ConfuseMatrixMeter:
    """Device-resident counterpart of utils/metric_tool.py:46-62.  `sum` is an int64 (n_class, n_class) CUDA tensor
    (hist[gt][pred]); `update_cm` takes device tensors and launches one kernel, no host synchronisation.
    `get_scores()` / `value_f1()` copy the matrix to the host."""

    def __init__(self, n_class: int, device="cuda"):
        self.n_class = n_class
        self.sum = torch.zeros(n_class, n_class, dtype=torch.int64, device=device)
        self.val = torch.zeros_like(self.sum)

    def clear(self) -> None:
        self.sum.zero_()
        self.val.zero_()

    def update_cm(self, pr: torch.Tensor, gt: torch.Tensor) -> None:
        """hist of this batch into `val`, added to `sum` (weight 1, as every reference call site)."""
        self.val.zero_()
        losses.confusion_matrix(gt, pr, self.n_class, self.val)
        self.sum += self.val

    def value_f1(self) -> float:
        """F1 of the last batch — what the reference's update_cm returns (reads the matrix back)."""
        return float(cm2F1(self.val.cpu().numpy()))

    def get_scores(self) -> dict:
        return cm2score(self.sum.cpu().numpy())
