"""GPU: every kernel behind the C ABI against the oracle / an fp64 torch statement of the same op."""
import pytest
import torch
import torch.nn.functional as F

import abi
import emulate as E
from oracle import dahitra_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def close(a, b, rtol=2e-5, atol=None):
    a, b = a.double().cpu(), b.double().cpu()
    atol = (atol if atol is not None else 2e-5 * float(b.abs().max()))
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), f"max|d|={float((a - b).abs().max()):.3e} (ref max {float(b.abs().max()):.3e}), {int(bad.sum())} bad"


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


CONVS = [  # (N, H, W, C0, C1, Cout, K, stride, up, res, relu, bias)
    (2, 16, 16, 64, 0, 64, 3, 1, 1, True, True, True),       # layer1 block conv2
    (1, 32, 32, 64, 0, 128, 3, 2, 1, False, True, True),     # layer2.0.conv1 (stride 2)
    (1, 32, 32, 64, 0, 128, 1, 2, 1, False, False, True),    # layer2.0 downsample
    (2, 16, 16, 128, 0, 256, 1, 1, 1, False, False, True),   # layer3.0 downsample
    (1, 16, 16, 256, 0, 256, 3, 1, 1, True, True, True),     # layer3
    (2, 16, 16, 32, 32, 32, 3, 1, 1, False, False, False),   # conv_decode on a virtual concat
    (1, 16, 32, 32, 0, 32, 3, 1, 2, False, True, True),      # conv_layerN on a virtually upsampled input
    (1, 32, 32, 64, 64, 128, 3, 1, 1, False, True, True),    # conv_layer2_0.0
    (1, 32, 32, 128, 0, 32, 3, 1, 1, True, False, True),     # conv_layer2_0.3 + out_3
    (1, 20, 28, 32, 0, 64, 3, 1, 1, False, True, True),      # ragged: not a multiple of the 8x16 tile
    (1, 7, 9, 32, 0, 32, 3, 2, 1, False, False, False),      # ragged + stride 2
]


@pytest.mark.parametrize("cfg", CONVS)
def test_conv2d(cfg):
    N, H, W, C0, C1, Cout, K, stride, up, res, relu, bias = cfg
    x0 = rnd(N, H, W, C0, seed=1)
    x1 = rnd(N, H, W, C1, seed=2) if C1 else None
    w = rnd(K * K * (C0 + C1), Cout, seed=3, scale=(K * K * (C0 + C1)) ** -0.5)
    b = rnd(Cout, seed=4) if bias else None
    pad = K // 2
    OH, OW = (H * up + 2 * pad - K) // stride + 1, (W * up + 2 * pad - K) // stride + 1
    r = rnd(N, OH, OW, Cout, seed=5) if res else None
    y = abi.conv2d(x0, x1, w, b, r, relu, K, stride, pad, up)
    xin = x0 if x1 is None else torch.cat([x0, x1], -1)
    ref = E.conv_nhwc(xin.double(), w.double(), None if b is None else b.double(), K, stride, pad,
                      None if r is None else r.double(), relu, up)
    assert y.shape == ref.shape
    close(y, ref)


TC_CONVS = [  # stride-1, un-upsampled shapes the tcgen05 kernel takes: (N, H, W, C0, C1, Cout, K, res, relu, bias)
    (2, 16, 16, 32, 0, 32, 1, False, False, False),          # smallest: one K step
    (1, 16, 16, 32, 0, 32, 3, False, False, False),          # 9 taps, zero padding through TMA OOB fill
    (2, 64, 64, 64, 0, 64, 3, True, True, True),             # layer1 conv2 (+identity, ReLU)
    (2, 32, 32, 128, 0, 128, 3, True, True, True),           # layer2
    (2, 16, 16, 128, 0, 256, 1, False, False, True),         # layer3.0 downsample (1x1)
    (1, 16, 16, 256, 0, 256, 3, True, True, True),           # layer3 (two N tiles, 72 K steps)
    (2, 32, 32, 32, 32, 32, 3, False, False, False),         # conv_decode on a virtual concat (two tensor maps)
    (1, 64, 64, 64, 64, 128, 3, False, True, True),          # conv_layer2_0.0
    (1, 64, 64, 128, 0, 32, 3, True, False, True),           # conv_layer2_0.3 + out_3
    (1, 20, 28, 32, 0, 64, 3, False, True, True),            # ragged edges
    (3, 16, 24, 64, 0, 128, 3, True, True, True),            # 9 M tiles: the last CTA pair has a tile without a partner
]


@pytest.mark.parametrize("shape", [(2, 16, 16), (1, 64, 64), (1, 24, 40)])
def test_conv2d_up2_tcgen05(shape):
    """nearest-x2 upsample + 3x3 conv (conv_layer4/3/2) as one low-res tcgen05 conv with a pixel-shuffle store."""
    from dahitra_b200.engine import upsample_phase_filter
    N, H, W = shape
    x = rnd(N, H, W, 32, seed=1)
    g = torch.Generator().manual_seed(2)
    w = torch.randn(32, 32, 3, 3, generator=g, dtype=torch.float64) * (288 ** -0.5)
    b = torch.randn(32, generator=g, dtype=torch.float64)
    from dahitra_b200.engine import kmajor_split
    wt, pb = upsample_phase_filter(w, b)
    wt = kmajor_split(wt).float().contiguous().to(DEV)
    ref = F.relu(F.conv2d(x.double().cpu().permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3), w, b, 1, 1))
    for flags, tol in ((0, 4e-3), (64, 4e-3), (2, 2e-5), (128, 4e-3), (130, 2e-5), (514, 6e-5), (642, 6e-5), (1024, 4e-2), (2562, 6e-5), (6658, 6e-5)):   # halo-reuse 1xTF32, per-tap kernel, 3xTF32, CTA pairs, bf16 corrections, bf16 operands, f16 main
        y = abi.conv2d_up2_tc(x, wt, pb.float().to(DEV), True, flags)
        torch.cuda.synchronize()
        close(y, ref.permute(0, 2, 3, 1), rtol=tol / 2, atol=tol)


@pytest.mark.parametrize("cfg", [(1, 64, 64, 64, 128, 3), (2, 32, 32, 64, 128, 1), (1, 24, 40, 32, 64, 3), (3, 36, 20, 64, 32, 3),
                                 (1, 128, 128, 128, 256, 3)])
def test_conv2d_stride2_tcgen05(cfg):
    """stride-2 convs (layer2.0 conv1 / downsample): the TMA box walks the input with element strides {1,2,2,1}."""
    N, H, W, Cin, Cout, K = cfg
    x = rnd(N, H, W, Cin, seed=1)
    w = rnd(K * K * Cin, Cout, seed=3, scale=(K * K * Cin) ** -0.5)
    b = rnd(Cout, seed=4)
    ref = E.conv_nhwc(x.double(), w.double(), b.double(), K, 2, K // 2, None, True, 1)
    # flags: 1|4 = halo-reuse kernel on the four phase images (1xTF32), 1|4|2 = same with 3xTF32, 1|4|64 = per-tap kernel
    for flags, tol in ((5, 4e-3), (7, 2e-5), (69, 4e-3), (133, 4e-3), (135, 2e-5), (519, 6e-5), (647, 6e-5), (1029, 4e-2), (2567, 6e-5), (6663, 6e-5)):   # + 128: CTA pairs; + 512: bf16 corrections; 1024: bf16 operands; 2048: f16 main product
        y = abi.conv2d(x, None, w, b, None, True, K, 2, K // 2, 1, flags=flags)
        torch.cuda.synchronize()
        assert y.shape == ref.shape
        print(f"[tc] stride-2 {cfg} flags={flags}: max|d|={float((y.double() - ref).abs().max()):.3e}")
        close(y, ref, rtol=tol / 2, atol=tol if not (flags & 2) else tol * float(ref.abs().max()))


@pytest.mark.parametrize("scale,tol", [(1.0, 6e-5), (1e-6, 6e-5), (3e4, 6e-3), (1e-9, 6e-3)])
def test_conv2d_f16_main_operand_range(scale, tol):
    """default error-compensated mode (FP16 main product + BF16 corrections): operands far outside the FP16 range stay
    finite and fall back on the BF16 terms — fp32-grade inside [1e-7, 65504], bf16-grade (stated: 6e-3 of the output
    range) beyond; the TF32-main variant (flags without 2048) has no such caveat."""
    N, H, W, C, Cout, K = 1, 32, 32, 64, 64, 3
    x = rnd(N, H, W, C, seed=1) * scale
    w = rnd(K * K * C, Cout, seed=3, scale=(K * K * C) ** -0.5)
    ref = E.conv_nhwc(x.double(), w.double(), None, K, 1, 1, None, False, 1)
    y = abi.conv2d(x, None, w, None, None, False, K, 1, 1, 1, flags=2563 | 4096)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    err = float((y.double().cpu() - ref.cpu()).abs().max()) / float(ref.abs().max())
    print(f"[tc] f16-main operand scale {scale:g}: max|d| / max|ref| = {err:.3e}")
    assert err <= tol
    y2 = abi.conv2d(x, None, w, None, None, False, K, 1, 1, 1, flags=515)      # TF32 main product: range-independent
    assert float((y2.double().cpu() - ref.cpu()).abs().max()) / float(ref.abs().max()) <= 6e-5


@pytest.mark.parametrize("cfg", TC_CONVS)
def test_conv2d_tcgen05(cfg):
    """tcgen05/TMEM/TMA implicit GEMM vs fp64: TF32 operands (10-bit mantissa), fp32 accumulate."""
    N, H, W, C0, C1, Cout, K, res, relu, bias = cfg
    x0 = rnd(N, H, W, C0, seed=1)
    x1 = rnd(N, H, W, C1, seed=2) if C1 else None
    Kt = K * K * (C0 + C1)
    w = rnd(Kt, Cout, seed=3, scale=Kt ** -0.5)
    b = rnd(Cout, seed=4) if bias else None
    r = rnd(N, H, W, Cout, seed=5) if res else None
    xin = x0 if x1 is None else torch.cat([x0, x1], -1)
    ref = E.conv_nhwc(xin.double(), w.double(), None if b is None else b.double(), K, 1, K // 2,
                      None if r is None else r.double(), relu, 1)
    # flags: 1 = halo-reuse kernel 1xTF32, 1|64 = per-tap kernel 1xTF32, 1|2 = halo-reuse kernel 3xTF32
    for flags, name, tol in ((1, "halo 1xTF32", 4e-3), (65, "per-tap 1xTF32", 4e-3), (3, "halo 3xTF32", 2e-5),
                             (129, "pair 1xTF32", 4e-3), (131, "pair 3xTF32", 2e-5), (515, "3xTF32 bf16-corr", 6e-5),
                             (643, "pair 3xTF32 bf16-corr", 6e-5), (1025, "bf16 operands", 4e-2),
                             (2563, "f16 main + bf16 corr", 6e-5), (2049, "f16 operands", 4e-3),
                             (6659, "f16 main, folded corr", 6e-5)):
        y = abi.conv2d(x0, x1, w, b, r, relu, K, 1, K // 2, 1, flags=flags)
        torch.cuda.synchronize()
        d = (y.double().cpu() - ref.cpu()).abs()
        print(f"[tc] {cfg} {name}: max|d|={float(d.max()):.3e} mean|d|={float(d.mean()):.3e} ref_absmax={float(ref.abs().max()):.3e}")
        # 1xTF32: ~2^-11 relative per product over Kt unit-variance terms; 3xTF32: fp32-grade (same bar as the FFMA kernel)
        close(y, ref, rtol=tol / 2, atol=tol if not (flags & 2) else tol * float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 96, 160), (1, 256, 256)])
def test_stem(shape):
    N, H, W = shape
    x = rnd(N, 3, H, W, seed=1)
    w = rnd(147, 64, seed=2, scale=147 ** -0.5)
    b = rnd(64, seed=3)
    y = abi.stem(x, w, b)
    close(y, E.stem(x.double(), w.double(), b.double()))


@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 96, 160), (1, 256, 256), (1, 40, 72)])
def test_stem_tcgen05(shape):
    """tensor-core stem (on-chip im2col + tcgen05, TF32) vs fp64."""
    from dahitra_b200.engine import stem_tc_image
    N, H, W = shape
    x = rnd(N, 3, H, W, seed=1)
    w = rnd(147, 64, seed=2, scale=147 ** -0.5)
    b = rnd(64, seed=3)
    wtc = stem_tc_image(w.cpu()).float().to(DEV)
    ref = E.stem(x.double(), w.double(), b.double())
    y = abi.stem_tc(x, wtc, b, x3=0)
    torch.cuda.synchronize()
    close(y, ref, rtol=2e-3, atol=4e-3)
    y3 = abi.stem_tc(x, wtc, b, x3=1)                      # error-compensated: same bar as the fp32 CUDA-core stem
    torch.cuda.synchronize()
    close(y3, ref)
    yf = abi.stem_tc(x, wtc, b, x3=2)                      # folded FP16 operands (the default mode's stem): same bar
    torch.cuda.synchronize()
    close(yf, ref)
    y1 = abi.stem_tc(x, wtc, b, x3=3)                      # single-pass FP16 (round-to-nearest operands: tighter than TF32)
    torch.cuda.synchronize()
    close(y1, ref, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("shape", [(2, 32, 32, 64), (1, 18, 30, 128)])
def test_maxpool_bit_exact(shape):
    x = rnd(*shape, seed=1)
    y = abi.maxpool(x)
    assert torch.equal(y, E.maxpool(x))


@pytest.mark.parametrize("k,npix", [(5, 256), (4, 1024), (3, 4096), (3, 1000)])
def test_tokens_and_encoder(k, npix, levir_template):
    """squeeze + spatial-softmax tokenizer + 1-layer encoder vs the oracle (networks.py:1273-1286)."""
    from dahitra_b200.engine import prepare_weights
    sd = synth.synth_state_dict(levir_template, seed=3, style="default")
    P = {n: (None if v is None else v.to(DEV)) for n, v in prepare_weights(sd, 0, 2).items()}
    cin, heads = O.LEVELS[k]["cin"], O.LEVELS[k]["heads"]
    B = 3
    h = npix // 8 if npix == 1000 else int(npix ** 0.5)
    w = npix // h
    feat = F.relu(rnd(2 * B, npix, cin, seed=7))
    s = f"DH_W_LV{k}_"
    xs, parts = abi.squeeze_tokens(feat, P[s + "SQ"], P[s + "TOK"])
    mem = abi.token_encoder(parts, B, P[s + "ENC"], heads, True)
    f = feat.cpu().double().transpose(1, 2).reshape(2 * B, cin, h, w)
    xs_ref = O.squeeze(sd, f, k, torch.float64)
    close(xs, xs_ref.flatten(2).transpose(1, 2))
    tok = torch.cat([O.semantic_tokens(sd, xs_ref[:B], k, torch.float64), O.semantic_tokens(sd, xs_ref[B:], k, torch.float64)], 1)
    tok = O.token_encoder(sd, tok, k, torch.float64)
    ref = torch.stack([tok[:, :4], tok[:, 4:], (tok[:, 4:] - tok[:, :4]).abs()], 1)
    close(mem, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))


@pytest.mark.parametrize("k,h,w", [(5, 16, 16), (4, 32, 32), (3, 64, 64), (3, 24, 40)])
def test_pixel_decoder(k, h, w, levir_template):
    """tables + streaming decoder vs TransformerDecoder as written (help_funcs.py:66-114,170-186)."""
    from dahitra_b200.engine import prepare_weights
    sd = dict(synth.synth_state_dict(levir_template, seed=3, style="default"))
    heads, depth = O.LEVELS[k]["heads"], O.LEVELS[k]["depth"]
    pos_nchw = torch.randn(1, 32, h, w, generator=torch.Generator().manual_seed(5))
    sd[f"pos_embedding_decoder_{k}"] = pos_nchw
    P = {n: (None if v is None else v.to(DEV)) for n, v in prepare_weights(sd, 0, 2).items()}
    B = 2
    x = rnd(B, h * w, 32, seed=8)
    mem = rnd(B, 3, 4, 32, seed=9)
    s = f"DH_W_LV{k}_"
    tab = abi.decoder_tables(mem, 0, 3, P[s + "DEC"], heads, depth)
    xn = x.cpu().double().transpose(1, 2).reshape(B, 32, h, w)
    for call in range(3):
        y = abi.pixel_decoder(x, P[s + "POS"], tab[call * B:(call + 1) * B].contiguous(), P[s + "DEC"], h, w, heads, depth)
        ref = O.pixel_decoder(sd, xn, mem[:, call].cpu().double(), k, torch.float64)
        close(y, ref.flatten(2).transpose(1, 2), rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
    # skip hooks: x2-upsampled coarser map, and same-size map
    if h % 2 == 0 and w % 2 == 0:
        sk = rnd(B, h // 2, w // 2, 32, seed=10)
        y2 = abi.pixel_decoder(x, P[s + "POS"], tab[:B].contiguous(), P[s + "DEC"], h, w, heads, depth, sk, 2)
        y1 = abi.pixel_decoder(x, P[s + "POS"], tab[:B].contiguous(), P[s + "DEC"], h, w, heads, depth)
        up = sk.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(B, h * w, 32)
        assert torch.equal(y2, y1 + up)
    sk = rnd(B, h, w, 32, seed=11)
    y3 = abi.pixel_decoder(x, None, tab[:B].contiguous(), P[s + "DEC"], h, w, heads, depth, sk, 1)
    y0 = abi.pixel_decoder(x, None, tab[:B].contiguous(), P[s + "DEC"], h, w, heads, depth)
    assert torch.equal(y3, y0 + sk.reshape(B, h * w, 32))


@pytest.mark.parametrize("k,h,w", [(5, 16, 16), (4, 32, 32), (3, 64, 64), (3, 24, 40)])
def test_pixel_decoder_tcgen05(k, h, w, levir_template):
    """tensor-core decoder (decoder_tc.cu) vs TransformerDecoder as written; TF32 operands -> 2e-3-level tolerance,
    and the skip hooks."""
    from dahitra_b200.engine import prepare_weights
    sd = dict(synth.synth_state_dict(levir_template, seed=3, style="default"))
    heads, depth = O.LEVELS[k]["heads"], O.LEVELS[k]["depth"]
    sd[f"pos_embedding_decoder_{k}"] = torch.randn(1, 32, h, w, generator=torch.Generator().manual_seed(5))
    P = {n: (None if v is None else v.to(DEV)) for n, v in prepare_weights(sd, 0, 2).items()}
    B = 2
    x = rnd(B, h * w, 32, seed=8)
    mem = rnd(B, 3, 4, 32, seed=9)
    s = f"DH_W_LV{k}_"
    tab = abi.decoder_tables_tc(mem, 0, 3, P[s + "DEC"], heads, depth)
    xn = x.cpu().double().transpose(1, 2).reshape(B, 32, h, w)
    for call in range(3):
        ref = O.pixel_decoder(sd, xn, mem[:, call].cpu().double(), k, torch.float64).flatten(2).transpose(1, 2)
        for x3 in (1, 0):
            y = abi.pixel_decoder_tc(x, P[s + "POS"], tab[call * B:(call + 1) * B].contiguous(), P[s + "DECTC"], h, w, heads,
                                     depth, x3=x3)
            torch.cuda.synchronize()
            d = (y.double().cpu() - ref).abs()
            print(f"[dec-tc] level {k} {h}x{w} call {call} {'3xTF32' if x3 else '1xTF32'}: max|d|={float(d.max()):.3e} "
                  f"mean|d|={float(d.mean()):.3e} ref_absmax={float(ref.abs().max()):.3e}")
            if x3:      # error-compensated: same tolerance as the fp32 CUDA-core decoder
                close(y, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
            else:       # single-pass TF32 on random (ill-conditioned) tables: percent-level worst case
                assert float(d.mean()) < 2e-3 * float(ref.abs().max()) and float(d.max()) < 5e-2 * float(ref.abs().max())
    sk = rnd(B, h, w, 32, seed=11)
    y3 = abi.pixel_decoder_tc(x, None, tab[:B].contiguous(), P[s + "DECTC"], h, w, heads, depth, sk, 1, x3=1)
    y0 = abi.pixel_decoder_tc(x, None, tab[:B].contiguous(), P[s + "DECTC"], h, w, heads, depth, x3=1)
    assert torch.equal(y3, y0 + sk.reshape(B, h * w, 32))


@pytest.mark.parametrize("nc", [2, 5])
def test_classifier_and_argmax(nc):
    x = rnd(2, 48, 80, 32, seed=1)
    w = rnd(9, nc, 32, seed=2, scale=0.1)
    b = rnd(nc, seed=3, scale=0.1)
    logits, am = abi.classifier(x, w, b, nc)
    wt = w.reshape(3, 3, nc, 32).permute(2, 3, 0, 1).double()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), wt, b.double(), 1, 1)
    close(logits, ref)
    assert torch.equal(am.long(), logits.argmax(1))        # the fused map is the argmax of the logits it wrote


def test_argument_errors():
    from dahitra_b200 import _lib
    lib = _lib.load()
    x = rnd(1, 8, 8, 48, seed=1)                           # 48 channels: not a multiple of 32
    w = rnd(9 * 48, 32, seed=2)
    out = torch.empty(1, 8, 8, 32, device=DEV)
    rc = lib.dahitra_conv2d(x.data_ptr(), None, 48, 0, 1, 8, 8, 1, 3, 3, 1, 1, 32, w.data_ptr(), None, None, None, 0,
                            out.data_ptr(), 0, None)
    assert rc == -2
    rc = lib.dahitra_conv2d(x.data_ptr() + 4, None, 32, 0, 1, 8, 8, 1, 3, 3, 1, 1, 32, w.data_ptr(), None, None, None, 0,
                            out.data_ptr(), 0, None)
    assert rc == -3
    with pytest.raises(RuntimeError, match="code -2"):
        _lib.check(-2, "x")


# ------------------------------------------------------------------------------------------------ split16 path (conv_tc3.cu)
def f16_split_ref(x):
    """host statement of the split16 format: hi = f16(a) saturating, lo = f16(2^11 (a - hi)); returns hi + 2^-11 lo (fp64)"""
    x = x.detach().cpu().float()
    hi = x.clamp(-65504.0, 65504.0).half()
    lo = ((x - hi.float()) * 2048.0).clamp(-65504.0, 65504.0).half()
    return hi.double() + lo.double() / 2048.0


def test_split16_pack_roundtrip():
    g = torch.Generator().manual_seed(1)
    x = torch.cat([torch.randn(4096, generator=g) * s for s in (1.0, 1e-3, 30.0, 1e-6, 3e4)] + [torch.zeros(64), torch.tensor([7e4, -7e4, 65504.0, 1e-9] * 2)])
    xd = x.to(DEV)
    s = abi.split_pack(xd)
    y = abi.split_unpack(s)
    ref = f16_split_ref(x)
    assert torch.equal(y.double().cpu(), ref)                     # bit-exact against the host statement of the format
    inr = (x.abs() > 6.2e-5) & (x.abs() < 6.5e4)
    rel = ((y.cpu() - x).abs() / x.abs().clamp_min(1e-30))[inr]
    assert float(rel.max()) <= 2.0 ** -21                          # 22 significant bits inside the normal FP16 range
    assert float((y.cpu() - x).abs()[x.abs() <= 6.2e-5].max()) <= 3.1e-11
    assert torch.equal(abi.split_unpack(abi.split_pack(y)), y)     # idempotent


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 20, 28, 128), (3, 64, 64, 64)])
def test_maxpool_split16(shape):
    x = rnd(*shape, seed=3)
    xq = abi.split_unpack(abi.split_pack(x))                       # representable values: the pool must be exact on them
    y = abi.split_unpack(abi.maxpool_split(abi.split_pack(x)))
    ref = F.max_pool2d(xq.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    # the pool orders the (hi, lo) pairs lexicographically: identical to the numeric order except for candidates that tie to
    # within the rounding of lo (2^-22 relative), where either one may win
    assert float(((y - ref).abs() / ref.abs().clamp_min(1e-6)).max()) <= 2.0 ** -20
    assert float((y == ref).float().mean()) >= 0.9999


SPLIT_CONVS = [  # (N, H, W, C0, C1, Cout, K, stride, res ('', 'f32', 'split'), relu, bias, out_split)
    (2, 16, 16, 32, 0, 32, 1, 1, '', False, False, False),         # smallest: one chunk, one tap
    (1, 16, 16, 32, 0, 32, 3, 1, '', False, False, False),         # 9 taps, zero padding through TMA OOB fill
    (2, 64, 64, 64, 0, 64, 3, 1, 'split', True, True, True),       # layer1 conv2 (+identity, ReLU), split16 in / res / out
    (2, 32, 32, 128, 0, 128, 3, 1, 'split', True, True, True),     # layer2
    (1, 64, 64, 64, 0, 128, 3, 2, '', True, True, True),           # layer2.0.conv1 (stride 2: four phase halos)
    (2, 32, 32, 64, 0, 128, 1, 2, '', False, True, True),          # layer2.0 downsample (1x1 stride 2)
    (2, 16, 16, 128, 0, 256, 1, 1, '', False, True, True),         # layer3.0 downsample (two N tiles)
    (1, 16, 16, 256, 0, 256, 3, 1, 'split', True, True, True),     # layer3 (72 (tap, chunk) steps)
    (2, 32, 32, 32, 32, 32, 3, 1, '', False, False, False),        # conv_decode on a virtual concat -> fp32
    (1, 64, 64, 64, 64, 128, 3, 1, '', True, True, True),          # conv_layer2_0.0
    (1, 64, 64, 128, 0, 32, 3, 1, 'f32', False, True, True),       # conv_layer2_0.3 + out_3 (fp32 residual)
    (1, 20, 28, 32, 0, 64, 3, 1, '', True, True, False),           # ragged edges
    (3, 36, 20, 64, 0, 32, 3, 2, '', False, False, True),          # ragged + stride 2
    (5, 48, 40, 64, 0, 64, 3, 1, 'split', True, True, True),       # more tiles than one wave of a small grid
]


SPLIT_CONVS += [
    (9, 32, 24, 128, 0, 128, 3, 1, 'split', True, True, True),     # 54 M tiles: resident / streaming, pairs with several tiles each
    (3, 16, 24, 64, 0, 128, 3, 1, '', True, True, True),           # 9 M tiles: the last CTA pair has a tile without a partner
    (40, 64, 64, 64, 0, 64, 3, 1, 'split', True, True, True),      # 1280 tiles: every CTA walks many tiles with the filter resident
    (24, 64, 64, 128, 0, 32, 3, 1, 'f32', False, True, True),      # K = 1152 -> 32 (conv_layer2_0.3): resident in a pair, 2 halo buffers alone
]


@pytest.mark.parametrize("sched", [0, 16, 32, 64, 32 | 64], ids=["auto", "cg1", "cg2", "stream", "cg2-stream"])
@pytest.mark.parametrize("cfg", SPLIT_CONVS)
def test_conv2d_split16(cfg, sched):
    """conv_tc3: split16 operands, three FP16 partial products, vs fp64 on the SAME representable inputs / weights; every
    scheduling variant (single CTAs / CTA pairs, filter resident in shared memory / streamed)."""
    N, H, W, C0, C1, Cout, K, stride, res, relu, bias, out_split = cfg
    x0 = rnd(N, H, W, C0, seed=1)
    x1 = rnd(N, H, W, C1, seed=2) if C1 else None
    Kt = K * K * (C0 + C1)
    w = rnd(Kt, Cout, seed=3, scale=Kt ** -0.5)
    b = rnd(Cout, seed=4) if bias else None
    OH, OW = H // stride, W // stride
    r = rnd(N, OH, OW, Cout, seed=5) if res else None
    s0, s1 = abi.split_pack(x0), (abi.split_pack(x1) if C1 else None)
    rs = None if r is None else (abi.split_pack(r) if res == 'split' else r)
    y = abi.conv2d_split(s0, s1, w, b, rs, relu, K, stride, res_split=(res == 'split'), out_split=out_split, sched=sched)
    torch.cuda.synchronize()
    xq = f16_split_ref(x0) if x1 is None else torch.cat([f16_split_ref(x0), f16_split_ref(x1)], -1)
    rq = None if r is None else (f16_split_ref(r) if res == 'split' else r.double().cpu())
    ref = E.conv_nhwc(xq, f16_split_ref(w), None if b is None else b.double().cpu(), K, stride, K // 2, rq, relu, 1)
    if out_split:
        ref = f16_split_ref(ref)
        y = abi.split_unpack(y)
    assert y.shape == ref.shape
    d = (y.double().cpu() - ref).abs()
    print(f"[tc3] {cfg}: max|d|={float(d.max()):.3e} mean|d|={float(d.mean()):.3e} ref_absmax={float(ref.abs().max()):.3e}")
    close(y, ref, rtol=1e-5, atol=2e-5 * float(ref.abs().max()))    # fp32 accumulation over up to 2304 terms (same bar as 3xTF32)
    # and against the un-quantised fp64 convolution: the format itself costs ~2^-22 per operand
    full = E.conv_nhwc((x0 if x1 is None else torch.cat([x0, x1], -1)).double().cpu(), w.double().cpu(),
                       None if b is None else b.double().cpu(), K, stride, K // 2, None if r is None else r.double().cpu(), relu, 1)
    close(y, full, rtol=1e-5, atol=2e-5 * float(full.abs().max()))


@pytest.mark.parametrize("shape", [(2, 16, 16), (1, 64, 64), (1, 24, 40), (12, 64, 64)])
def test_conv2d_up2_split16(shape):
    """nearest-x2 upsample + 3x3 conv (conv_layer4/3/2) on a split16 input, pixel-shuffle store to fp32."""
    from dahitra_b200.engine import upsample_phase_filter
    N, H, W = shape
    x = rnd(N, H, W, 32, seed=1)
    g = torch.Generator().manual_seed(2)
    w = torch.randn(32, 32, 3, 3, generator=g, dtype=torch.float64) * (288 ** -0.5)
    b = torch.randn(32, generator=g, dtype=torch.float64)
    wt, pb = upsample_phase_filter(w, b)
    ref = F.relu(F.conv2d(x.double().cpu().permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3), w, b, 1, 1))
    for sched in (0, 16, 32, 64):
        y = abi.conv2d_split(abi.split_pack(x), None, wt, pb.float().to(DEV), None, True, 3, 1, mode=1, sched=sched)
        torch.cuda.synchronize()
        close(y, ref.permute(0, 2, 3, 1), rtol=1e-5, atol=2e-5 * float(ref.abs().max()))


@pytest.mark.parametrize("cfg", [(2, 16, 16, 256), (2, 32, 32, 128), (2, 64, 64, 64), (1, 24, 40, 64), (1, 256, 256, 64)])
@pytest.mark.parametrize("xs_split", [False, True])
def test_squeeze_tokens_split16(cfg, xs_split):
    """conv_tc3 tok epilogue: 1x1 squeeze + ReLU + per-tile softmax partials, merged by the token encoder's rule, vs fp64
    (reference models/networks.py:1177-1189, 1273-1280)."""
    N, H, W, Cin = cfg
    feat = rnd(N, H, W, Cin, seed=1)
    wsq = rnd(Cin, 32, seed=2, scale=Cin ** -0.5)
    wtok = rnd(32, 4, seed=3, scale=1.0)
    xs, parts = abi.conv2d_split(abi.split_pack(feat), None, wsq, None, None, False, 1, 1, out_split=xs_split, mode=2, wtok=wtok)
    torch.cuda.synchronize()
    if xs_split:
        xs = abi.split_unpack(xs)
    fq = f16_split_ref(feat).reshape(N, H * W, Cin)
    xr = torch.relu(fq @ f16_split_ref(wsq))
    close(xs.reshape(N, H * W, 32), f16_split_ref(xr) if xs_split else xr, rtol=1e-5, atol=2e-5 * float(xr.abs().max()))
    # merge the partials as token_encoder_kernel does and compare with the fp64 tokenizer
    p = parts.double().cpu()                                       # [N][nchunk][4][34]
    M = p[..., 0].max(dim=1, keepdim=True).values
    sc = torch.exp(p[..., 0] - M)
    S = (p[..., 1] * sc).sum(1)
    T = (p[..., 2:] * sc[..., None]).sum(1)
    tok = T / S[..., None]                                         # [N][4][32]
    a = torch.softmax(torch.einsum("npc,cl->nlp", xr, wtok.double().cpu()), dim=-1)
    ref = torch.einsum("nlp,npc->nlc", a, xr)
    close(tok, ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max()))
