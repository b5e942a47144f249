"""Checkpoint I/O helpers.

The reference saves ``{'model_G_state_dict': net_G.state_dict(), ...}`` (models/trainer.py:150-160) and loads it
strictly (models/evaluator.py:68-75); the xBD scripts save ``{'state_dict': model.state_dict(), ...}`` of an
``nn.DataParallel`` wrapper, i.e. with ``module.`` prefixes (xBD_code/train.py:447-457).  These helpers turn either
into a plain state_dict for the drop-in modules, convert between the LEVIR (425 keys) and xBD (700 keys: the
``nn.ModuleList`` containers add ``*_layers.N.*`` aliases of the per-level keys) key layouts, and export the prepared
(BN-folded, re-laid-out) weight slots for consumers of the C ABI that do not go through PyTorch.  The ``*_bin`` functions at
the end are the file formats of the plain-C host ``examples/dahitra_infer.c``.

Converting the KEY LAYOUT does not make the two variants compute the same function (the xBD forward runs one decoder
pass per level and applies positional terms on one level only, xBD_code/zoo/model_transformer_encoding.py:358-406).
"""
from __future__ import annotations

import re

import numpy as np
import torch

# xBD container name -> per-level attribute prefix; container index i <-> level i + 2
_ALIASES = {"conv_squeeze_layers": "conv_squeeze", "conv_tokens_layers": "conv_token", "conv_decode_layers": "conv_decode",
            "transformer_layers": "transformer", "transformer_decoder_layers": "transformer_decoder"}
_ALIAS_RE = re.compile(r"^(%s)\.(\d)\.(.*)$" % "|".join(_ALIASES))
_LEVIR_ONLY = ("pos_embedding_2", "pos_embedding_decoder_2")


def extract_state_dict(obj) -> dict:
    """checkpoint object (what torch.load returns) or state_dict -> plain state_dict without DataParallel prefixes"""
    sd = obj
    if isinstance(obj, dict):
        for key in ("model_G_state_dict", "state_dict"):
            if key in obj and isinstance(obj[key], dict):
                sd = obj[key]
                break
    if not isinstance(sd, dict) or not sd or not all(isinstance(k, str) for k in sd):
        raise ValueError("dahitra_b200.checkpoints: no state_dict found in the checkpoint object")
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}


def load_checkpoint(path: str) -> dict:
    """torch.load on the CPU + extract_state_dict (reference checkpoints are plain pickles of tensors and numbers)"""
    return extract_state_dict(torch.load(path, map_location="cpu", weights_only=False))


def xbd_to_levir(sd: dict, template: dict | None = None) -> dict:
    """drop the ModuleList alias keys; LEVIR-only keys (scale-2 positional embeddings, never used by the forward) are
    taken from `template` when given, else zero-filled with the reference shapes"""
    out = {k: v for k, v in sd.items() if not _ALIAS_RE.match(k)}
    shapes = {"pos_embedding_2": (1, 8, 32), "pos_embedding_decoder_2": (1, 32, 64, 64)}
    for k in _LEVIR_ONLY:
        if k not in out:
            out[k] = template[k].clone() if template is not None and k in template else torch.zeros(shapes[k])
    return out


def levir_to_xbd(sd: dict) -> dict:
    """add the ModuleList alias keys (sharing storage with the per-level tensors) and drop the LEVIR-only keys"""
    out = {k: v for k, v in sd.items() if k not in _LEVIR_ONLY}
    for cont, prefix in _ALIASES.items():
        for k, v in sd.items():
            m = re.match(r"^%s_(\d)\.(.*)$" % prefix, k)
            if m and 2 <= int(m.group(1)) <= 5:
                out[f"{cont}.{int(m.group(1)) - 2}.{m.group(2)}"] = v
    return out


def export_prepared(module, path: str) -> dict:
    """write the prepared weight slots of `module` (what dahitra_forward's pointer table refers to, fp32, keyed by
    the DH_W_* slot names of include/dahitra_b200.h) to an .npz; returns {slot: shape}"""
    from .engine import prepare_weights, DH_VARIANT_LEVIR, DH_VARIANT_XBD
    variant = DH_VARIANT_LEVIR if module.VARIANT == "levir" else DH_VARIANT_XBD
    P = prepare_weights(module.state_dict(), variant, module.output_nc)
    arrays = {k: v.numpy() for k, v in P.items() if v is not None}
    np.savez(path, **arrays)
    return {k: tuple(a.shape) for k, a in arrays.items()}


# ---- flat binary files of the plain-C host (examples/dahitra_infer.c; formats in its header comment) -----------------------
def export_state_dict_bin(sd: dict, path: str) -> int:
    """write the floating-point tensors of a state_dict (reference keys, models/trainer.py:150-158) as a DHSD0001 file — the
    `dh_tensor` list dahitra_prepare_weights takes, for a host without PyTorch; returns the number of tensors written"""
    import struct
    keep = [(k, v.detach().cpu().contiguous()) for k, v in extract_state_dict(sd).items()
            if torch.is_tensor(v) and v.dtype.is_floating_point and v.dim() <= 4]
    with open(path, "wb") as f:
        f.write(b"DHSD0001" + struct.pack("<i", len(keep)))
        for k, v in keep:
            v = v if v.dtype == torch.float64 else v.float()
            name = k.encode()
            shape = list(v.shape) + [1] * (4 - v.dim())
            data = v.numpy().tobytes()
            f.write(struct.pack("<i", len(name)) + name + struct.pack("<ii4qq", int(v.dtype == torch.float64), v.dim(), *shape, len(data)))
            f.write(data)
    return len(keep)


def write_pairs_bin(path: str, x1: torch.Tensor, x2: torch.Tensor | None = None) -> None:
    """DHIN0001 input file: two (B,3,H,W) tensors (LEVIR: pre, post) or one (B,6,H,W) tensor (xBD: stacked on the channels)"""
    import struct
    ts = [x1] if x2 is None else [x1, x2]
    B, C, H, W = ts[0].shape
    if C != (6 if x2 is None else 3) or any(t.shape != ts[0].shape for t in ts):
        raise ValueError("write_pairs_bin: expected two (B,3,H,W) tensors or one (B,6,H,W) tensor")
    with open(path, "wb") as f:
        f.write(b"DHIN0001" + struct.pack("<4i", B, C, H, W))
        for t in ts:
            f.write(t.detach().cpu().float().contiguous().numpy().tobytes())


def read_result_bin(path: str):
    """DHOUT001 result file -> (logits (B,nc,H,W) fp32, class map (B,H,W) uint8)"""
    import struct
    raw = open(path, "rb").read()
    if raw[:8] != b"DHOUT001":
        raise ValueError(f"{path} is not a DHOUT001 file")
    B, nc, H, W = struct.unpack("<4i", raw[8:24])
    n = B * nc * H * W
    if len(raw) != 24 + 4 * n + B * H * W:
        raise ValueError(f"{path}: {len(raw)} bytes for B={B} nc={nc} H={H} W={W}")
    logits = torch.from_numpy(np.frombuffer(raw, dtype="<f4", count=n, offset=24).reshape(B, nc, H, W).copy())
    cmap = torch.from_numpy(np.frombuffer(raw, dtype=np.uint8, count=B * H * W, offset=24 + 4 * n).reshape(B, H, W).copy())
    return logits, cmap
