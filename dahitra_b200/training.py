"""Native pieces of the TRAINING step: ``torch.autograd.Function``s over hand-written sm_100a kernels.

The training step of the reference (models/trainer.py:247-262: forward, CE loss, backward, optimizer step) runs on the drop-in
module's autograd route.  What that route spends its device time on is the pixel decoders
(``profiles/r02_train_step_ablation.json``: 27 of 36 ms of forward + backward at batch 8), so those run natively:

* ``pixel_decoder(x, tables, heads)`` — all layers of one ``TransformerDecoder`` call (reference models/help_funcs.py:66-114,
  170-186) for every pixel in ONE forward launch, and the whole backward in ONE launch that returns dL/dx and the gradient of the
  per-(image, layer) tables (``modules.PixelDecoder.train_tables`` builds those from the parameters and the 4 memory tokens with
  differentiable torch ops on tensors of a few KB, so autograd carries the table gradients on to Wq / Wk / Wv / Wo, the
  LayerNorm affines, the MLP, the tokens and everything upstream of the tokens).

There is no fallback: the functions raise if the CUDA library is missing or the tensors are not on a GPU.
"""
from __future__ import annotations

import torch

from . import _lib


def train_tab_floats(heads: int) -> int:
    """DH_TRAIN_TAB_FLOATS of include/dahitra_b200.h"""
    return 65 * 4 * heads + 96 + 2048


class _PixelDecoderTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tables, heads):
        if not (x.is_cuda and tables.is_cuda):
            raise RuntimeError("dahitra_b200.training: tensors must be on a CUDA device — there is no CPU path")
        if x.dtype != torch.float32 or tables.dtype != torch.float32:
            raise RuntimeError("dahitra_b200.training: fp32 tensors only")
        B, C, N = x.shape
        depth = tables.shape[1]
        if C != 32 or tables.shape[0] != B or tables.shape[2] != train_tab_floats(heads):
            raise RuntimeError(f"pixel_decoder: x {tuple(x.shape)} / tables {tuple(tables.shape)} do not fit heads={heads}")
        x, tables = x.contiguous(), tables.contiguous()
        lib = _lib.load()
        xs = torch.empty((depth, B, 32, N), dtype=torch.float32, device=x.device)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = lib.dahitra_pixel_decoder_train_fwd(x.data_ptr(), tables.data_ptr(), xs.data_ptr(), out.data_ptr(), B, N, heads,
                                                     depth, torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc, "dahitra_pixel_decoder_train_fwd")
        ctx.save_for_backward(xs, tables)
        ctx.heads = heads
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        xs, tables = ctx.saved_tensors
        depth, B, _, N = xs.shape
        lib = _lib.load()
        dout = dout.contiguous()
        nblk = lib.dahitra_pixel_decoder_train_blocks(N)
        dx = torch.empty_like(dout)
        partial = torch.empty((B, nblk, depth, tables.shape[2]), dtype=torch.float32, device=dout.device)
        with torch.cuda.device(dout.device):
            rc = lib.dahitra_pixel_decoder_train_bwd(dout.data_ptr(), xs.data_ptr(), tables.data_ptr(), dx.data_ptr(),
                                                     partial.data_ptr(), B, N, ctx.heads, depth,
                                                     torch.cuda.current_stream(dout.device).cuda_stream)
        _lib.check(rc, "dahitra_pixel_decoder_train_bwd")
        return dx, partial.sum(1), None


def pixel_decoder(x: torch.Tensor, tables: torch.Tensor, heads: int) -> torch.Tensor:
    """x (B, 32, N) fp32 CUDA (channel-planar = flattened NCHW), tables (B, depth, DH_TRAIN_TAB_FLOATS(heads)) -> (B, 32, N).
    Differentiable in x and tables (first order)."""
    return _PixelDecoderTrain.apply(x, tables, heads)
