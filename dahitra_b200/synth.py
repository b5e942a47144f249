"""Deterministic synthetic weights and inputs (there is no network for datasets or checkpoints): shared by
bench.py, the tools, the tests and the oracle pin script.  Weights are drawn per key from a torch CPU generator (bit-reproducible across
machines running the same torch build), at scales that exercise every folded term: BatchNorm
running statistics and affine parameters are non-trivial, biases are non-zero, LayerNorm affine is
non-identity.  ``style='default'`` gives O(0.3) logits (catches what the tiny N(0,0.02) define_G init
hides — SURVEY.md §8c); ``style='small'`` mimics define_G's N(0, 0.02).
"""
from __future__ import annotations

import hashlib

import torch


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:6], "little")


def synth_state_dict(template: dict, seed: int = 0, style: str = "default") -> dict:
    """template: {key: tensor} giving keys/shapes/dtypes (e.g. a fresh module's state_dict)."""
    out = {}
    first_key_of = {}      # alias keys (nn.ModuleList views of the same tensor, xBD variant) share one value
    for key, t in template.items():
        owner = first_key_of.setdefault((t.data_ptr(), tuple(t.shape)), key) if t.numel() > 0 else key
        if owner != key:
            out[key] = out[owner]
            continue
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        shp = tuple(t.shape)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("num_batches_tracked"):
            v = torch.tensor(7, dtype=t.dtype)
        elif leaf == "running_mean":
            v = 0.1 * torch.randn(shp, generator=g)
        elif leaf == "running_var":
            v = 0.5 + torch.rand(shp, generator=g)
        elif t.dim() == 1 and leaf == "weight":                     # BN / LN gamma
            v = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif leaf == "bias":
            v = 0.05 * torch.randn(shp, generator=g)
        elif key.startswith("pos_embedding"):
            v = (1.0 if style == "default" else 0.5) * torch.randn(shp, generator=g)
        elif t.dim() >= 2:                                           # conv / linear weights
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            std = (1.0 / fan_in) ** 0.5 if style == "default" else 0.02
            if style == "default" and ("to_q" in key or "to_k" in key or "to_qkv" in key):
                std *= 3.0                                           # make the attention softmax non-uniform
            v = std * torch.randn(shp, generator=g)
        else:
            v = torch.randn(shp, generator=g)
        out[key] = v.to(t.dtype)
    return out


def synth_pair(B: int, H: int, W: int, seed: int = 1, kind: str = "normal"):
    g = torch.Generator().manual_seed(seed)
    if kind == "normal":
        return torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
    if kind == "uniform":      # what the loaders produce: [-1, 1]
        return torch.rand(B, 3, H, W, generator=g) * 2 - 1, torch.rand(B, 3, H, W, generator=g) * 2 - 1
    if kind == "u8":           # uint8 image statistics through (x/255 - .5)/.5 (datasets/data_utils.py:106-111)
        a = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
        b = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
        return (a / 255 - 0.5) / 0.5, (b / 255 - 0.5) / 0.5
    raise ValueError(kind)


def fingerprint(sd: dict) -> dict:
    """Per-key float64 sums — small enough to commit, strong enough to detect RNG / layout drift."""
    return {k: float(v.double().sum()) for k, v in sd.items()}
