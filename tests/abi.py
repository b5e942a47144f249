"""Test helper: call the per-kernel C-ABI entry points with torch CUDA tensors (NHWC fp32)."""
import torch

from dahitra_b200 import _lib


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def conv2d(in0, in1, w, bias, res, relu, K, stride, pad, up=1, flags=0):
    lib = _lib.load()
    N, inH, inW, C0 = in0.shape
    C1 = 0 if in1 is None else in1.shape[-1]
    Cout = w.shape[1]
    H, W = inH * up, inW * up
    OH, OW = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
    out = torch.empty((N, OH, OW, Cout), device=in0.device, dtype=torch.float32)
    wt = None
    if flags:                                              # K-major [2][Cout][K] (TF32 hi, lo) for the tcgen05 paths
        from dahitra_b200.engine import kmajor_split
        wt = kmajor_split(w.detach().cpu().double().t()).float().contiguous().to(w.device)
    rc = lib.dahitra_conv2d(_p(in0), _p(in1), C0, C1, N, inH, inW, up, K, K, stride, pad, Cout, _p(w), _p(wt), _p(bias),
                            _p(res), int(relu), _p(out), flags, _stream())
    _lib.check(rc, "dahitra_conv2d")
    return out


def conv2d_up2_tc(x, pswt, psb, relu, flags=0):
    lib = _lib.load()
    N, H, W, _ = x.shape
    out = torch.empty((N, 2 * H, 2 * W, 32), device=x.device, dtype=torch.float32)
    _lib.check(lib.dahitra_conv2d_up2_tc(_p(x), N, H, W, _p(pswt), _p(psb), int(relu), _p(out), flags, _stream()),
               "dahitra_conv2d_up2_tc")
    return out


def stem(x, w, b):
    lib = _lib.load()
    N, _, H, W = x.shape
    out = torch.empty((N, H // 2, W // 2, 64), device=x.device, dtype=torch.float32)
    _lib.check(lib.dahitra_stem(_p(x), 3 * H * W, N, H, W, _p(w), _p(b), _p(out), _stream()), "dahitra_stem")
    return out


def stem_tc(x, wtc, b, x3=0):
    lib = _lib.load()
    N, _, H, W = x.shape
    out = torch.empty((N, H // 2, W // 2, 64), device=x.device, dtype=torch.float32)
    _lib.check(lib.dahitra_stem_tc(_p(x), 3 * H * W, N, H, W, _p(wtc), _p(b), _p(out), x3, _stream()), "dahitra_stem_tc")
    return out


def maxpool(x):
    lib = _lib.load()
    N, H, W, C = x.shape
    out = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), device=x.device, dtype=torch.float32)
    _lib.check(lib.dahitra_maxpool3x3s2(_p(x), N, H, W, C, _p(out), _stream()), "dahitra_maxpool3x3s2")
    return out


def squeeze_tokens(feat, wsq, wtok):
    lib = _lib.load()
    N, npix, Cin = feat.shape
    nchunk = (npix + 255) // 256
    xs = torch.empty((N, npix, 32), device=feat.device, dtype=torch.float32)
    parts = torch.empty((N, nchunk, 4, 34), device=feat.device, dtype=torch.float32)
    _lib.check(lib.dahitra_squeeze_tokens(_p(feat), N, npix, Cin, _p(wsq), _p(wtok), _p(xs), _p(parts), _stream()),
               "dahitra_squeeze_tokens")
    return xs, parts


def token_encoder(parts, B, enc, heads, add_pos):
    lib = _lib.load()
    mem = torch.empty((B, 3, 4, 32), device=parts.device, dtype=torch.float32)
    _lib.check(lib.dahitra_token_encoder(_p(parts), B, parts.shape[1], _p(enc), heads, int(add_pos), _p(mem), _stream()),
               "dahitra_token_encoder")
    return mem


def decoder_tables(mem, first_call, ncalls, dec, heads, depth):
    lib = _lib.load()
    B = mem.shape[0]
    tabf = 32 * 4 * heads + 4 * heads + 4 * heads * 32 + 32
    tab = torch.empty((ncalls * B, depth, tabf), device=mem.device, dtype=torch.float32)
    _lib.check(lib.dahitra_decoder_tables(_p(mem), B, first_call, ncalls, _p(dec), heads, depth, _p(tab), _stream()),
               "dahitra_decoder_tables")
    return tab


def pixel_decoder(x, pos, tab, dec, h, w, heads, depth, skip=None, skip_up=1):
    lib = _lib.load()
    nimg = x.shape[0]
    out = torch.empty_like(x)
    _lib.check(lib.dahitra_pixel_decoder(_p(x), _p(pos), _p(tab), _p(dec), nimg, h, w, heads, depth, _p(skip), skip_up,
                                         _p(out), _stream()), "dahitra_pixel_decoder")
    return out


def decoder_tables_tc(mem, first_call, ncalls, dec, heads, depth):
    lib = _lib.load()
    B = mem.shape[0]
    tab = torch.empty((ncalls * B, depth, 4128), device=mem.device, dtype=torch.float32)
    _lib.check(lib.dahitra_decoder_tables_tc(_p(mem), B, first_call, ncalls, _p(dec), heads, depth, _p(tab), _stream()),
               "dahitra_decoder_tables_tc")
    return tab


def pixel_decoder_tc(x, pos, tab, dectc, h, w, heads, depth, skip=None, skip_up=1, x3=0):
    lib = _lib.load()
    out = torch.empty_like(x)
    _lib.check(lib.dahitra_pixel_decoder_tc(_p(x), _p(pos), _p(tab), _p(dectc), x.shape[0], h, w, heads, depth, _p(skip),
                                            skip_up, x3, _p(out), _stream()), "dahitra_pixel_decoder_tc")
    return out


def classifier(x, w, b, nc, want_argmax=True):
    lib = _lib.load()
    N, H, W, _ = x.shape
    logits = torch.empty((N, nc, H, W), device=x.device, dtype=torch.float32)
    am = torch.empty((N, H, W), device=x.device, dtype=torch.uint8) if want_argmax else None
    _lib.check(lib.dahitra_classifier(_p(x), N, H, W, nc, _p(w), _p(b), _p(logits), _p(am), _stream()), "dahitra_classifier")
    return logits, am


# ---- split16 activation format (conv_tc3.cu): a tensor of n elements = two FP16 planes (hi | lo) in a buffer of n floats
def split_pack(x):
    lib = _lib.load()
    x = x.contiguous()
    out = torch.empty_like(x)                       # same bytes: 2 planes x 2 bytes per element
    _lib.check(lib.dahitra_split_pack(_p(x), x.numel(), _p(out), _stream()), "dahitra_split_pack")
    return out


def split_unpack(s):
    lib = _lib.load()
    out = torch.empty_like(s)
    _lib.check(lib.dahitra_split_unpack(_p(s), s.numel(), _p(out), _stream()), "dahitra_split_unpack")
    return out


def maxpool_split(s):
    lib = _lib.load()
    N, H, W, C = s.shape
    out = torch.empty((N, H // 2, W // 2, C), device=s.device, dtype=torch.float32)
    _lib.check(lib.dahitra_maxpool3x3s2_split(_p(s), N, H, W, C, _p(out), _stream()), "dahitra_maxpool3x3s2_split")
    return out


def conv2d_split(in0, in1, w, bias, res, relu, K, stride, res_split=False, out_split=False, mode=0, wtok=None, sched=0):
    """in0 / in1: split16 buffers shaped like their fp32 tensors (N, H, W, C); w: [K*K*Cin][Cout] fp32 (split on the host
    exactly as the engine does).  mode 1: w is the [128][288] phase filter already K-major.  Returns out (fp32 NHWC, or a
    split16 buffer of that shape) and, in mode 2, the partials."""
    lib = _lib.load()
    from dahitra_b200.engine import kmajor_split
    N, inH, inW, C0 = in0.shape
    C1 = 0 if in1 is None else in1.shape[-1]
    wk = w.detach().cpu().double()
    wt = kmajor_split(wk if mode == 1 else wk.t()).float().contiguous().to(in0.device)
    Cout = wt.shape[1]
    OH, OW = inH // stride, inW // stride
    shape = (N, 2 * OH, 2 * OW, 32) if mode == 1 else (N, OH, OW, Cout)
    out = torch.empty(shape, device=in0.device, dtype=torch.float32)
    parts = None
    if mode == 2:
        nchunk = ((inH + 15) // 16) * ((inW + 7) // 8)
        parts = torch.empty((N, nchunk, 4, 34), device=in0.device, dtype=torch.float32)
    rc = lib.dahitra_conv2d_split(_p(in0), _p(in1), C0, C1, 0, 0, N, inH, inW, K, stride, Cout, _p(wt), _p(bias), _p(res),
                                  int(res_split), int(relu), _p(out), int(out_split), mode | sched, _p(wtok), _p(parts), _stream())
    _lib.check(rc, "dahitra_conv2d_split")
    return (out, parts) if mode == 2 else out
