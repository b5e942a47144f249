"""Device-side change/damage-map metrics.

Mirrors the reference's ``ConfuseMatrixMeter`` / ``cm2score`` (misc/metric_tool.py:48-63, 75-138, 141-158) with
the confusion matrix accumulated ON the GPU from uint8 class maps (``dahitra_confusion_matrix``), so an evaluation
loop no longer moves an int64 map to the host and bincounts it per batch (models/evaluator.py:95-104).  Only the
nc x nc int64 matrix ever crosses PCIe.  ``cm2score`` / ``cm2F1`` restate the reference formulas term by term
(float64 numpy, same epsilons) so the reported numbers are identical.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_EPS = np.finfo(np.float32).eps


def confusion_matrix(pred: torch.Tensor, gt: torch.Tensor, n_class: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """cm[g][p] += count — pred / gt: CUDA integer class maps of equal numel (any integer dtype; gt values outside
    [0, n_class) are ignored like the reference's mask).  Returns (and accumulates into) a CUDA int64 [n_class, n_class]."""
    if not (pred.is_cuda and gt.is_cuda):
        raise RuntimeError("dahitra_b200.metrics: tensors must be on a CUDA device")
    if pred.numel() != gt.numel():
        raise RuntimeError("pred and gt must have the same number of pixels")
    p8 = pred if pred.dtype == torch.uint8 else pred.to(torch.uint8)
    if gt.dtype == torch.uint8:
        g8 = gt
    else:                                   # negative / large labels must stay outside [0, n_class)
        g8 = torch.where((gt >= 0) & (gt < n_class), gt, torch.full_like(gt, 255)).to(torch.uint8)
    p8, g8 = p8.contiguous(), g8.contiguous()
    if out is None:
        out = torch.zeros((n_class, n_class), dtype=torch.int64, device=pred.device)
    if p8.numel() == 0:
        return out
    with torch.cuda.device(pred.device):
        rc = _lib.load().dahitra_confusion_matrix(p8.data_ptr(), g8.data_ptr(), p8.numel(), n_class, out.data_ptr(),
                                                  torch.cuda.current_stream(pred.device).cuda_stream)
    _lib.check(rc, "dahitra_confusion_matrix")
    return out


def cm2F1(hist) -> float:
    """misc/metric_tool.py:75-96"""
    hist = np.asarray(hist, dtype=np.float64)
    tp = np.diag(hist)
    recall = tp / (hist.sum(axis=1) + _EPS)
    precision = tp / (hist.sum(axis=0) + _EPS)
    f1 = 2 * recall * precision / (recall + precision + _EPS)
    return float(np.nanmean(f1))


def cm2score(hist) -> dict:
    """misc/metric_tool.py:99-138 — acc, miou, mf1 and per-class iou / F1 / precision / recall."""
    hist = np.asarray(hist, dtype=np.float64)
    n_class = hist.shape[0]
    tp = np.diag(hist)
    sum_a1, sum_a0 = hist.sum(axis=1), hist.sum(axis=0)
    acc = tp.sum() / (hist.sum() + _EPS)
    recall = tp / (sum_a1 + _EPS)
    precision = tp / (sum_a0 + _EPS)
    f1 = 2 * recall * precision / (recall + precision + _EPS)
    iu = tp / (sum_a1 + sum_a0 - tp + _EPS)
    score = {"acc": acc, "miou": np.nanmean(iu), "mf1": np.nanmean(f1)}
    score.update({f"iou_{i}": iu[i] for i in range(n_class)})
    score.update({f"F1_{i}": f1[i] for i in range(n_class)})
    score.update({f"precision_{i}": precision[i] for i in range(n_class)})
    score.update({f"recall_{i}": recall[i] for i in range(n_class)})
    return score


class DeviceConfuseMatrixMeter:
    """Drop-in for the way the harness uses ``ConfuseMatrixMeter`` (update_cm / get_scores / clear), fed with CUDA
    tensors; the running matrix lives on the device."""

    def __init__(self, n_class: int):
        self.n_class = n_class
        self.sum = None

    def update_cm(self, pr: torch.Tensor, gt: torch.Tensor, weight: int = 1) -> float:
        cur = confusion_matrix(pr, gt, self.n_class)
        self.sum = cur * weight if self.sum is None else self.sum + cur * weight
        return cm2F1(cur.cpu().numpy())

    def get_scores(self) -> dict:
        return cm2score(self.sum.cpu().numpy())

    def clear(self):
        self.sum = None
