"""xBD harness pieces that sit directly on the forward (SURVEY.md §8 f2 tail / config 3), restated on the device:

* ``flip4_tta`` — the 4-flip test-time augmentation of xBD_code/predict_test_cls.py:69-91: the network sees the image,
  its vertical flip, its horizontal flip and both; the sigmoid outputs are flipped back and averaged.
* ``damage_map`` — the validation rule of xBD_code/train.py:266-273: localisation = sigmoid(ch 0) > thr, damage class =
  argmax over sigmoid(ch 1..4), multiplied by the localisation mask.

Plain torch ops on CUDA tensors around the native forward; nothing here touches the host.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def flip4_tta(net, x: torch.Tensor) -> torch.Tensor:
    """x: (B, 6, H, W) normalised pre|post stack on the device -> (B, nc, H, W) averaged sigmoid probabilities.
    Flip order and un-flip bookkeeping as in predict_test_cls.py:69-91 (img[::-1], img[:, ::-1], img[::-1, ::-1])."""
    flips = ((), (2,), (3,), (2, 3))
    acc = None
    for dims in flips:
        xi = torch.flip(x, dims) if dims else x
        p = torch.sigmoid(net(xi.contiguous()))
        p = torch.flip(p, dims) if dims else p
        acc = p if acc is None else acc + p
    return acc / len(flips)


@torch.no_grad()
def damage_map(out: torch.Tensor, thr: float = 0.3, probabilities: bool = False) -> torch.Tensor:
    """out: (B, 5, H, W) logits (or probabilities) -> (B, H, W) int64: argmax over channels 1..4 (0..3), zeroed where the
    localisation channel is at or below `thr` (train.py:266-273)."""
    p = out if probabilities else torch.sigmoid(out)
    return p[:, 1:].argmax(dim=1) * (p[:, 0] > thr)
