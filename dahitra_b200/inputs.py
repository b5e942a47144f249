"""Device-side input path: decoded uint8 HWC images go to the GPU as they are (4x fewer bytes over
PCIe than fp32 tensors) and are normalised — and, for the 1024² -> 256² evaluation of config 5, tiled — by one kernel,
bit-identically to the reference loaders (datasets/data_utils.py:65-66,104-111; xBD_code/utils.py:112-116)."""
from __future__ import annotations

import torch

from . import _lib

KINDS = {"levir": 0, "xbd": 1}


def normalize_u8(img: torch.Tensor, kind: str = "levir", tile: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """img: CUDA uint8 (N,H,W,3) -> fp32 (N*T,3,h,w), T = tiles per image (1 when tile == 0)."""
    if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
        raise RuntimeError("dahitra_b200.inputs: expected a CUDA uint8 tensor of shape (N,H,W,3)")
    if kind not in KINDS:
        raise ValueError(kind)
    img = img.contiguous()
    N, H, W, _ = img.shape
    th, tw = (tile, tile) if tile else (H, W)
    T = (H // th) * (W // tw)
    if out is None:
        out = torch.empty((N * T, 3, th, tw), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        rc = _lib.load().dahitra_prepare_input_u8(img.data_ptr(), N, H, W, KINDS[kind], tile, out.data_ptr(),
                                                  torch.cuda.current_stream(img.device).cuda_stream)
    _lib.check(rc, "dahitra_prepare_input_u8")
    return out
