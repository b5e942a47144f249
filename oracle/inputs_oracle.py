"""ORACLE — test infrastructure, NOT product code.

numpy restatement of the reference's input normalisation and evaluation tiling, used to check
dahitra_b200.inputs.  fp32 arithmetic in the reference's operation order: parity is bit-exact.
"""
import numpy as np


def normalize_levir(u8_hwc):
    """TF.to_tensor (uint8 HWC -> float32 CHW / 255) then TF.normalize(mean .5, std .5) — datasets/data_utils.py:104-111"""
    x = u8_hwc.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    return (x - np.float32(0.5)) / np.float32(0.5)


def normalize_xbd(u8_hwc):
    """preprocess_inputs — xBD_code/utils.py:112-116 (x /= 127; x -= 1 in float32), then HWC -> CHW as the loaders do"""
    x = np.asarray(u8_hwc, dtype='float32')
    x /= 127
    x -= 1
    return x.transpose(2, 0, 1)


def tiles(u8_hwc, size=256):
    """evaluation patches of a big image in the reference's order (datasets/data_utils.py:65-66, with patch 0 at the
    origin instead of the `if patch:` fall-through to (256,256)): x0 = size*(p // n), y0 = size*(p % n)"""
    n = u8_hwc.shape[0] // size
    out = []
    for p in range(n * (u8_hwc.shape[1] // size)):
        x0, y0 = size * (p // n), size * (p % n)
        out.append(u8_hwc[y0:y0 + size, x0:x0 + size, :])
    return out
