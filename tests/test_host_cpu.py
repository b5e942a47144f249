"""CPU: host-side logic — weight preparation algebra, the C-ABI library surface, module behaviour without a GPU."""
import ctypes
import os
import re

import pytest
import torch

import emulate as E
from oracle import dahitra_oracle as O
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dahitra_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "dahitra_b200.h")).read()
    declared = set(re.findall(r"\b(dahitra_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 13
    from dahitra_b200 import _lib
    assert declared == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dahitra_version() == 2
    assert b"ok" == lib.dahitra_error_string(0)
    assert b"workspace" in lib.dahitra_error_string(-4)


def test_slot_table_matches_preparation(lib, levir_template):
    from dahitra_b200.engine import slot_names, prepare_weights
    names = slot_names()
    assert names[0] == "DH_W_STEM_W" and "DH_W_CLS_B" in names and len(names) == len(set(names))
    hdr = open(os.path.join(ROOT, "include", "dahitra_b200.h")).read()
    start = hdr.index("enum dh_weight_slot")
    enum_body = hdr[start:hdr.index("DH_W_COUNT", start)]
    assert names == re.findall(r"\b(DH_W_[A-Z0-9_]+)\b", re.sub(r"/\*.*?\*/", "", enum_body, flags=re.S))
    P = prepare_weights(levir_template, 0, 2)
    assert set(P) == set(names)
    assert P["DH_W_LV3_ENC"].numel() == 8 * 32 + 64 + 2 * 8 * 1024 + 32 + 64 + 1024 + 32 + 1024 + 32
    assert P["DH_W_LV5_DEC"].numel() == 4 * (64 + 2 * 4 * 1024 + 32 + 1024 + 32 + 1024 + 32)


def test_workspace_and_argument_checks_without_gpu(lib):
    assert lib.dahitra_workspace_bytes(0, 1, 256, 256, 2, 0) > 0
    assert lib.dahitra_workspace_bytes(0, 1, 250, 256, 2, 0) == 0         # not a multiple of 32
    assert lib.dahitra_workspace_bytes(7, 1, 256, 256, 2, 0) == 0         # unknown variant
    big = lib.dahitra_workspace_bytes(1, 8, 1024, 1024, 5, 0)
    assert 1 << 30 < big < 40 << 30
    # host-side validation happens before any CUDA call, so these are safe on a GPU-less box
    rc = lib.dahitra_forward(None, 60, None, None, 0, None, None, None, 0, 0, 1, 256, 256, 2, 0, None)
    assert rc == -1
    rc = lib.dahitra_conv2d(None, None, 32, 0, 1, 8, 8, 1, 3, 3, 1, 1, 32, None, None, None, None, 0, None, 0, None)
    assert rc == -1


def test_module_refuses_cpu_inference(levir_template):
    from dahitra_b200.networks import BASE_Transformer_UNet
    net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8).eval()
    x = torch.zeros(1, 3, 256, 256)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        net(x, x)


def test_define_G_scope():
    from dahitra_b200.networks import define_G

    class A:
        net_G = "base_resnet18"
    with pytest.raises(NotImplementedError):
        define_G(A())


@pytest.mark.parametrize("variant", ["levir"])
def test_prepared_weights_algebra_matches_oracle(variant, levir_template):
    """Runs the kernels' algebra (collapsed attention, folded BN/LN, re-laid-out filters) in torch fp64 on the
    prepared packs and compares with the fp64 oracle: validates engine.prepare_weights end to end."""
    from dahitra_b200.engine import prepare_weights
    sd = synth.synth_state_dict(levir_template, seed=3, style="default")
    P = {k: (None if v is None else v.double()) for k, v in prepare_weights(sd, 0, 2).items()}
    x1, x2 = synth.synth_pair(1, 256, 256, seed=2, kind="uniform")
    y = O.forward_levir(sd, x1, x2, dtype=torch.float64)
    e = E.forward(P, x1.double(), x2.double(), "levir", 2)
    # packs are stored in fp32: expect ~1e-6 relative, nothing structural
    assert float((y - e).abs().max()) < 5e-5 * float(y.abs().max())


def test_upsample_phase_filter_algebra():
    """upsample(x2, nearest) -> conv3x3  ==  conv3x3 (32 -> 4*32) on the low-res map -> pixel shuffle."""
    import torch.nn.functional as F
    from dahitra_b200.engine import upsample_phase_filter
    g = torch.Generator().manual_seed(0)
    w = torch.randn(32, 32, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(32, generator=g, dtype=torch.float64)
    x = torch.randn(2, 32, 6, 10, generator=g, dtype=torch.float64)
    ref = F.conv2d(x.repeat_interleave(2, 2).repeat_interleave(2, 3), w, b, 1, 1)
    wt, pb = upsample_phase_filter(w, b)                       # [128][9*32] K-major, [128]
    w3 = wt.reshape(128, 3, 3, 32).permute(0, 3, 1, 2)         # -> OIHW
    y = F.conv2d(x, w3, pb, 1, 1)                              # (2, 128, 6, 10): channel = phase*32 + co
    y = y.reshape(2, 2, 2, 32, 6, 10).permute(0, 3, 4, 1, 5, 2).reshape(2, 32, 12, 20)
    assert float((y - ref).abs().max()) < 1e-12


def test_filter_planes_of_the_precision_modes():
    """engine.kmajor_split: every plane decodes to what its mode's MMAs expect, and the products each mode issues
    reproduce a.w to that mode's accuracy (emulated in fp64 on operands rounded like the kernels round them)."""
    from dahitra_b200.engine import kmajor_split, tf32_split
    g = torch.Generator().manual_seed(1)
    cout, K, M = 64, 9 * 64, 200
    w = torch.randn(cout, K, generator=g, dtype=torch.float64) * K ** -0.5
    a = torch.randn(M, K, generator=g, dtype=torch.float64)
    ref = a @ w.T
    P = kmajor_split(w)
    assert P.shape == (5, cout, K) and P.dtype == torch.float32
    hi, lo = P[0].double(), P[1].double()
    assert torch.equal(hi, tf32_split(w)[0]) and float((w - hi - lo).abs().max()) < 2 ** -21 * float(w.abs().max())
    b16 = P[2].contiguous().view(torch.int16).reshape(2, cout, K).view(torch.bfloat16).double()       # {bf16(w), bf16(w - hi)}
    f16 = P[3].contiguous().view(torch.int16).reshape(2, cout, K)
    wh, wr = f16[0].view(torch.float16).double(), f16[1].view(torch.bfloat16).double()               # {f16(w), bf16(w - f16 w)}
    ws = P[4].contiguous().view(torch.int16).reshape(2, cout, K)[0].view(torch.float16).double()      # f16(2^11 (w - f16 w))
    assert torch.equal(b16[0], w.float().bfloat16().double()) and torch.equal(wh, w.float().half().double())
    ah, a_t = a.float().half().double(), tf32_split(a)[0]
    ar = (a - ah).float().bfloat16().double()
    err = lambda y: float((y - ref).abs().max()) / float(ref.abs().max())
    e_f16 = err(ah @ wh.T)                                                       # single-pass FP16 (mode "f16")
    e_x3 = err(a_t @ hi.T + (a - a_t).float().bfloat16().double() @ b16[0].T + a.float().bfloat16().double() @ b16[1].T)   # XM = 2
    e_main16 = err(ah @ wh.T + ar @ b16[0].T + ah.float().bfloat16().double() @ wr.T)          # XM = 4
    e_fold = err(ah @ wh.T + (ah @ ws.T) / 2048.0 + ar @ b16[0].T)                              # XM = 6 (default)
    assert 1e-5 < e_f16 < 2e-3
    assert e_x3 < 3e-6 and e_main16 < 3e-6 and e_fold < 2e-6
    assert e_fold <= e_main16 * 1.05                          # 11-bit scaled remainder beats the 8-bit bf16 one


def test_stem_folded_fp16_image_algebra():
    """DH_W_STEM_WTC 16-bit images: de-swizzled, the three products of the folded FP16 stem
    f16(a).f16(w) + 2^-11 f16(a).f16(2^11 r_w) + bf16(r_a).bf16(w) over K = (ci, r, s8) reproduce the 7x7 stride-2 conv."""
    import torch.nn.functional as F
    from dahitra_b200.engine import stem_tc_image
    g = torch.Generator().manual_seed(0)
    w = torch.randn(7, 7, 3, 64, generator=g, dtype=torch.float64) * 147 ** -0.5            # [r][s][ci][co]
    x = torch.randn(1, 3, 20, 24, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w.permute(3, 2, 0, 1), None, 2, 3)                                      # (1, 64, 10, 12)
    img = stem_tc_image(w.reshape(147, 64))
    assert img.shape == (43008,) and img.dtype == torch.float32
    n = torch.arange(128)[:, None].expand(128, 64)
    k = torch.arange(64)[None, :].expand(128, 64)
    idx = n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7))
    main = img[24576:36864].view(torch.int16).view(3, 128 * 64)
    corr = img[36864:43008].view(torch.int16).view(3, 64 * 64)
    wm = torch.cat([main[kt][idx.reshape(-1)].view(128, 64).view(torch.float16).double() for kt in range(3)], 1)    # [128][192]
    wc = torch.cat([corr[kt][idx[:64].reshape(-1)].view(64, 64).view(torch.bfloat16).double() for kt in range(3)], 1)  # [64][192]
    # im2col rows in the kernel's K order: group = ci*7 + r, 8 consecutive columns starting one left of the first tap
    xp = F.pad(x, (4, 4, 3, 3))[0]                                                            # col 0 = image col -4
    A = torch.zeros(10 * 12, 192, dtype=torch.float64)
    for oy in range(10):
        for ox in range(12):
            for ci in range(3):
                for r in range(7):
                    A[oy * 12 + ox, (ci * 7 + r) * 8:(ci * 7 + r) * 8 + 8] = xp[ci, 2 * oy + r, 2 * ox:2 * ox + 8]
    ah = A.float().half().double()
    al = (A - ah).float().bfloat16().double()
    y = ah @ wm[:64].T + (ah @ wm[64:].T) / 2048.0 + al @ wc.T
    y = y.T.reshape(1, 64, 10, 12)
    assert float((y - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    assert float((y - ref).abs().max()) < 0.02 * float((ah @ wm[:64].T - ref.reshape(64, -1).T).abs().max())   # the corrections do the work


def test_tensor_core_decoder_pack(levir_template):
    """DH_W_LVk_DECTC: swizzled W1f / W2 images and cumulative biases reproduce the CUDA-core pack's algebra."""
    from dahitra_b200.engine import prepare_weights
    sd = synth.synth_state_dict(levir_template, seed=3, style="default")
    P = prepare_weights(sd, 0, 2)
    r = torch.arange(32)[:, None].expand(32, 32)
    k = torch.arange(32)[None, :].expand(32, 32)
    idx = (r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3))).reshape(-1)
    for lvl, heads, depth in ((5, 4, 4), (3, 8, 8)):
        tc = P[f"DH_W_LV{lvl}_DECTC"].view(depth, 4192)
        cum = torch.zeros(32)
        for l in range(depth):
            p = E.unpack_dec_layer(P[f"DH_W_LV{lvl}_DEC"], heads, l)
            w1h, w2h = tc[l, :1024][idx].view(32, 32), tc[l, 1024:2048][idx].view(32, 32)      # B[n=o][k=c], B[n=c][k=o]
            w1l, w2l = tc[l, 2144:3168][idx].view(32, 32), tc[l, 3168:4192][idx].view(32, 32)
            for t in (w1h, w2h, w1l, w2l):                 # every tile is exactly TF32-representable
                assert int((t.view(torch.int32) & 0x1FFF).abs().sum()) == 0
            assert torch.allclose(w1h + w1l, p["W1f"].T, rtol=1e-6, atol=1e-9)
            assert torch.allclose(w2h + w2l, p["W2t"].T, rtol=1e-6, atol=1e-9)
            assert float((w1h - p["W1f"].T).abs().max()) <= 2 ** -11 * float(p["W1f"].abs().max())
            assert torch.allclose(tc[l, 2048:2080], p["b1f"], atol=1e-7)
            cum = cum + p["bo"]
            assert torch.allclose(tc[l, 2080:2112], cum, atol=1e-6)
            cum = cum + p["b2"]
            assert torch.allclose(tc[l, 2112:2144], cum, atol=1e-6)


def test_prepared_weights_algebra_xbd():
    from dahitra_b200.engine import prepare_weights
    from dahitra_b200.xbd import BASE_Transformer_UNet as X
    net = X(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned",
            with_decoder_pos="learned", enc_depth=1, dec_depth=8)
    # run at 256x256 with the H/16-level embedding cropped to 16x16: the xBD variant itself only runs at 1024^2
    sd = synth.synth_state_dict(net.state_dict(), seed=6, style="default")
    sd = dict(sd)
    sd["pos_embedding_decoder_3"] = sd["pos_embedding_decoder_3"][:, :, :16, :16].contiguous()
    P = {k: (None if v is None else v.double()) for k, v in prepare_weights(sd, 1, 5).items()}
    g = torch.Generator().manual_seed(7)
    x = torch.rand(1, 6, 256, 256, generator=g) * 2 - 1
    y = O.forward_xbd(sd, x, dtype=torch.float64)
    e = E.forward(P, x[:, :3].double(), x[:, 3:].double(), "xbd", 5)
    assert float((y - e).abs().max()) < 5e-5 * float(y.abs().max())


def test_precision_mode_selection(monkeypatch):
    """engine default = tf32x3; DAHITRA_MODE / DAHITRA_FLAGS / set_mode() override it; unknown names are rejected."""
    from dahitra_b200 import engine as E
    monkeypatch.delenv("DAHITRA_MODE", raising=False)
    monkeypatch.delenv("DAHITRA_FLAGS", raising=False)
    e = E.NativeEngine()
    assert E.DEFAULT_MODE == "tf32x3" and e.mode == "tf32x3" and e.flags == E.MODES["tf32x3"]
    e.set_mode("fp32")
    assert e.flags == 0 and e.mode == "fp32"
    e.set_mode(29)
    assert e.mode == "tf32_fast"
    with pytest.raises(ValueError):
        e.set_mode("int8")
    monkeypatch.setenv("DAHITRA_MODE", "tf32")
    assert E.NativeEngine().flags == E.MODES["tf32"]
    monkeypatch.setenv("DAHITRA_FLAGS", "0")
    assert E.NativeEngine().mode == "fp32"


def test_header_is_plain_c_and_mode_macros_match_engine(tmp_path):
    """include/dahitra_b200.h compiles as C (the drop-in boundary is a C ABI) and its DH_FLAGS_* mode macros are the
    bitmasks dahitra_b200.engine.MODES uses."""
    import shutil
    import subprocess
    from dahitra_b200 import engine as E
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "chk.c"
    src.write_text('#include "dahitra_b200.h"\n#include <stdio.h>\n'
                   'int main(void){ printf("%d %d %d %d %d\\n", DH_FLAGS_TF32X3, DH_FLAGS_TF32, DH_FLAGS_F16, DH_FLAGS_BF16, (int)DH_W_COUNT); return 0; }\n')
    exe = tmp_path / "chk"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(v) for v in out[:4]] == [E.MODES["tf32x3"], E.MODES["tf32"], E.MODES["f16"], E.MODES["bf16"]]
    assert int(out[4]) == len(E.slot_names())


# ---------------------------------------------------------------------------------------------- engine cache hygiene
def test_prepared_weight_cache_sees_every_in_place_update():
    """The prepared-weight cache is keyed on the live tensors' version counters: load_state_dict on a PARENT module,
    in-place copies in eval mode (EMA / SWA), optimizer steps and requires_grad_ toggles + writes all change the key."""
    import torch
    from dahitra_b200.networks import BASE_Transformer_UNet
    net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8).eval()
    eng = net._engine
    fp0 = eng._live_fingerprint(net)
    assert eng._live_fingerprint(net) == fp0                       # stable while nothing changes
    wrapper = torch.nn.Sequential(net)
    wrapper.load_state_dict({k: v.clone() for k, v in wrapper.state_dict().items()})
    fp1 = eng._live_fingerprint(net)
    assert fp1 != fp0
    with torch.no_grad():
        net.classifier.weight.copy_(net.classifier.weight * 0.5)   # EMA-style update in eval mode
    fp2 = eng._live_fingerprint(net)
    assert fp2 != fp1
    opt = torch.optim.SGD([net.classifier.bias], lr=0.1)
    net.classifier.bias.grad = torch.ones_like(net.classifier.bias)
    opt.step()
    assert eng._live_fingerprint(net) != fp2
    fp3 = eng._live_fingerprint(net)
    with torch.no_grad():
        net.resnet.bn1.running_mean.add_(1.0)                      # buffers count too
    assert eng._live_fingerprint(net) != fp3


def test_engine_survives_deepcopy_and_pickle():
    import copy
    import ctypes as C
    import pickle
    import torch
    from dahitra_b200.networks import BASE_Transformer_UNet
    net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8).eval()
    net.set_mode("tf32")
    # what the engine holds after a GPU forward: a ctypes pointer table (not picklable by itself)
    net._engine._preps[("cuda:0", "levir")] = type("P", (), {"table": (C.c_void_p * 3)()})()
    net._engine._ws[("cuda:0", 0)] = torch.zeros(4)
    twin = copy.deepcopy(net)
    assert twin._engine is not net._engine and twin._engine._preps == {} and twin._engine._ws == {}
    assert twin._engine.mode == "tf32"
    blob = pickle.dumps(net)
    back = pickle.loads(blob)
    assert back._engine._preps == {} and back._engine.mode == "tf32"
    assert len(back.state_dict()) == 425


def test_xbd_variant_without_token_pos_prepares():
    """reference default with_pos=None (xBD_code/zoo/model_transformer_encoding.py constructor): no pos_embedding_3."""
    from dahitra_b200.xbd import BASE_Transformer_UNet as X
    from dahitra_b200.engine import prepare_weights
    net = X(3, 5, None, with_decoder_pos='learned')
    assert "pos_embedding_3" not in net.state_dict()
    P = prepare_weights(net.state_dict(), 1, 5)
    assert float(P["DH_W_LV5_ENC"][:256].abs().sum()) == 0.0


# ---------------------------------------------------------------------------------------------- C-ABI weight preparation
@pytest.mark.parametrize("variant", ["levir", "xbd"])
def test_c_prepare_weights_matches_python_specification(lib, variant):
    """dahitra_prepare_weights (csrc/prepare.cu, plain C++ on the host) against engine.prepare_weights slot by slot: bit-exact
    for everything that is a re-layout / fold / format conversion, <= 1 ulp(fp32) for the fp64 matrix products whose
    summation order differs from torch's einsum."""
    import torch
    from dahitra_b200 import synth
    from dahitra_b200.engine import prepare_weights, prepare_weights_c, slot_names
    if variant == "levir":
        from dahitra_b200.networks import BASE_Transformer_UNet
        net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
        vid, nc = 0, 2
    else:
        from dahitra_b200.xbd import BASE_Transformer_UNet as X
        net = X(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned", with_decoder_pos="learned", enc_depth=1, dec_depth=8)
        vid, nc = 1, 5
    sd = synth.synth_state_dict(net.state_dict(), seed=21, style="default")
    P = prepare_weights(sd, vid, nc)
    flat, offs = prepare_weights_c(sd, vid, nc)
    names = slot_names()
    assert len(offs) == len(names)
    products = ("_ENC", "_DEC", "_DECTC")                  # slots containing fp64 matrix products
    for name, off in zip(names, offs):
        ref = P[name]
        if ref is None:
            assert off == -1, name
            continue
        assert off >= 0 and off % 64 == 0, name
        got = flat[off:off + ref.numel()].view(ref.shape)
        if name.endswith(products):
            assert torch.allclose(got, ref, rtol=3e-7, atol=1e-30), (name, float((got - ref).abs().max()))
            assert float((got != ref).float().mean()) < 1e-3, name
        else:
            assert torch.equal(got.view(torch.int32), ref.view(torch.int32)), (name, float((got - ref).abs().max()))
    # errors: a missing key is reported, not papered over
    sd2 = {k: v for k, v in sd.items() if k != "resnet.layer2.0.bn1.running_var"}
    with pytest.raises(RuntimeError, match="weight"):
        prepare_weights_c(sd2, vid, nc)


def test_collapsed_decoder_training_route_equals_as_written():
    """modules.PixelDecoder.forward_collapsed (the training route's default) is the same function as the reference's
    as-written decoder (help_funcs.py:66-114,170-186): outputs and ALL parameter / input gradients agree in fp64."""
    import torch
    from dahitra_b200 import modules as M
    torch.manual_seed(3)
    for heads, depth in ((4, 2), (8, 3)):
        dec = M.PixelDecoder(32, depth, heads, 64, 32).double()
        for p in dec.parameters():                       # non-trivial LayerNorm affine / biases
            p.data.add_(0.1 * torch.randn_like(p))
        x = torch.randn(2, 50, 32, dtype=torch.float64, requires_grad=True)
        m = torch.randn(2, 4, 32, dtype=torch.float64, requires_grad=True)
        w = torch.randn(2, 50, 32, dtype=torch.float64)
        outs = []
        for fn in (dec, dec.forward_collapsed):
            for t in list(dec.parameters()) + [x, m]:
                t.grad = None
            y = fn(x, m)
            (y * w).sum().backward()
            outs.append((y.detach().clone(), [t.grad.clone() for t in list(dec.parameters()) + [x, m]]))
        assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-11, atol=1e-12)
        for ga, gb in zip(outs[0][1], outs[1][1]):
            assert torch.allclose(ga, gb, rtol=1e-9, atol=1e-11), float((ga - gb).abs().max())


def test_train_tables_reproduce_the_decoder():
    """modules.PixelDecoder.train_tables (host half of the native training decoder) + the kernel's algebra on the table layout
    (tests/emulate.py) is the same function as the reference's as-written decoder (help_funcs.py:66-114,170-186): outputs and
    ALL parameter / input gradients agree in fp64."""
    import torch
    import emulate as E
    from dahitra_b200 import modules as M
    from dahitra_b200.training import train_tab_floats
    torch.manual_seed(4)
    for heads, depth in ((4, 2), (8, 3)):
        dec = M.PixelDecoder(32, depth, heads, 64, 32).double()
        for p in dec.parameters():
            p.data.add_(0.1 * torch.randn_like(p))
        x = torch.randn(2, 50, 32, dtype=torch.float64, requires_grad=True)
        m = torch.randn(2, 4, 32, dtype=torch.float64, requires_grad=True)
        w = torch.randn(2, 50, 32, dtype=torch.float64)
        outs = []
        for native in (False, True):
            for t in list(dec.parameters()) + [x, m]:
                t.grad = None
            if native:
                tab = dec.train_tables(m)
                assert tab.shape == (2, depth, train_tab_floats(heads))
                y = E.train_decoder_from_tables(x.transpose(1, 2), tab, heads).transpose(1, 2)
            else:
                y = dec(x, m)
            (y * w).sum().backward()
            outs.append((y.detach().clone(), [t.grad.clone() for t in list(dec.parameters()) + [x, m]]))
        assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-11, atol=1e-12)
        for ga, gb in zip(outs[0][1], outs[1][1]):
            assert torch.allclose(ga, gb, rtol=1e-9, atol=1e-11), float((ga - gb).abs().max())


def test_unused_parameter_names_are_exactly_the_ones_without_a_path_to_the_output():
    """48 LEVIR parameters (scale-2 modules, conv_pred, layer4, fc) never receive a gradient (9.18 M of 13.38 M);
    freeze_unused_parameters() marks exactly those."""
    import contextlib, io
    import torch
    from dahitra_b200.networks import define_G

    class A:
        net_G = "newUNetTrans"
    with contextlib.redirect_stdout(io.StringIO()):
        net = define_G(A())
    names = net.unused_parameter_names()
    assert len(names) == 48
    n_unused = sum(p.numel() for n, p in net.named_parameters() if n in set(names))
    n_all = sum(p.numel() for p in net.parameters())
    assert abs(n_unused - 9.18e6) < 0.02e6
    # autograd agrees: run the stock route on CPU (the public forward refuses CPU tensors; the route itself is plain torch)
    net.native_training = False
    x = torch.randn(1, 3, 256, 256)
    net._forward_autograd(x, x.flip(-1)).sum().backward()
    no_grad = sorted(n for n, p in net.named_parameters() if p.grad is None)
    assert no_grad == sorted(names)
    assert net.freeze_unused_parameters() == sorted(names)
    assert all(p.requires_grad != (n in set(names)) for n, p in net.named_parameters())


def test_filters_outside_the_fp16_operand_range_are_reported(levir_template):
    """the default mode's convolution operands are FP16 pairs (saturating at 65504): a checkpoint whose BatchNorm-folded filters
    leave that range, or hold a NaN, is reported when its weights are prepared instead of silently losing accuracy"""
    import warnings
    from dahitra_b200.engine import PreparedWeights
    from dahitra_b200 import synth
    sd = synth.synth_state_dict(levir_template, seed=4, style="default")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        P = PreparedWeights(sd, 0, 2, "cpu")
    assert not w and 0 < P.filter_absmax[0] < 65504 and P.filter_absmax[1].startswith("DH_W_")
    big = dict(sd)
    big["resnet.layer2.0.bn1.running_var"] = torch.full_like(sd["resnet.layer2.0.bn1.running_var"], 1e-30)   # fold scale 1/sqrt(eps) = 316
    big["resnet.layer2.0.bn1.weight"] = sd["resnet.layer2.0.bn1.weight"] * 1e4
    with pytest.warns(RuntimeWarning, match="DH_W_L2_0_C1_W.*outside the FP16 operand range"):
        P = PreparedWeights(big, 0, 2, "cpu")
    assert P.filter_absmax[0] > 65504
    bad = dict(sd)
    bad["classifier.weight"] = sd["classifier.weight"].clone()
    bad["classifier.weight"][1, 3, 0, 2] = float("nan")
    with pytest.warns(RuntimeWarning, match="DH_W_CLS_W"):
        P = PreparedWeights(bad, 0, 2, "cpu")
    assert P.filter_absmax[0] != P.filter_absmax[0]


def test_build_staleness_is_decided_by_content_not_file_times(lib, monkeypatch):
    """the in-tree library travels between machines (file times do not survive every copy): `needs_build` compares a content hash
    of the sources / headers / flags with the stamp the build left next to the library"""
    from dahitra_b200 import _lib
    assert not _lib.needs_build() and os.path.exists(_lib.STAMP_PATH)
    src = os.path.join(_lib._CSRC, "aux.cu")
    st = os.stat(src)
    try:
        os.utime(src, (st.st_atime, os.path.getmtime(_lib.LIB_PATH) + 1000))      # "newer" than the library, same bytes
        assert not _lib.needs_build()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    monkeypatch.setattr(_lib, "_source_hash", lambda: "0" * 64)                      # any edited source
    assert _lib.needs_build()
    monkeypatch.undo()
    monkeypatch.setenv("DAHITRA_DEBUG_BUILD", "1")                                  # other flags = another library
    assert _lib.needs_build()


def test_imagenet_trunk_file_is_loaded_like_the_reference_does(tmp_path, monkeypatch):
    """reference models/networks.py:1096: resnet18(pretrained=True), then init_weights re-draws Conv / BatchNorm affines
    (:88-105) — the ImageNet BatchNorm running statistics survive define_G.  Here the same file is read from a local path only
    (never downloaded); without it the statistics start at 0 / 1; a path asked for explicitly must exist."""
    from dahitra_b200 import modules as M
    from dahitra_b200.networks import define_G
    from dahitra_b200.xbd import BASE_Transformer_UNet as X

    class Args:
        net_G = "newUNetTrans"
    monkeypatch.delenv("DAHITRA_RESNET18_CKPT", raising=False)
    monkeypatch.setattr(torch.hub, "get_dir", lambda: str(tmp_path / "hub"))          # an empty hub cache
    torch.manual_seed(0)
    plain = define_G(Args(), gpu_ids=[])
    assert float(plain.resnet.bn1.running_mean.abs().sum()) == 0.0
    # a file in the layout of resnet18-5c106cde.pth: torchvision keys, no num_batches_tracked entries
    torch.manual_seed(99)
    donor = M.Trunk()
    for m in donor.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2.0)
    ckpt = {k: v for k, v in donor.state_dict().items() if not k.endswith("num_batches_tracked")}
    path = str(tmp_path / "resnet18-5c106cde.pth")
    torch.save(ckpt, path)
    monkeypatch.setenv("DAHITRA_RESNET18_CKPT", path)
    torch.manual_seed(0)
    net = define_G(Args(), gpu_ids=[])
    for k, v in net.resnet.state_dict().items():
        if "running_" in k:
            assert torch.equal(v, ckpt[k]), k                                          # survive init_weights
        elif k.endswith("weight") and v.dim() == 4:
            assert torch.equal(v, plain.resnet.state_dict()[k]) and not torch.equal(v, ckpt[k]), k   # re-drawn, same RNG stream as without the file
    assert torch.equal(net.resnet.fc.weight, plain.resnet.fc.weight)                  # init_weights re-draws Linear too
    # the xBD variant has no init_weights pass: the whole ImageNet trunk stays (xBD_code/zoo/model_transformer_encoding.py:195)
    xnet = X(3, 5, with_pos="learned")
    assert torch.equal(xnet.resnet.layer1[0].conv1.weight, ckpt["layer1.0.conv1.weight"])
    # and the same through the hub cache the reference's own call fills
    monkeypatch.delenv("DAHITRA_RESNET18_CKPT")
    os.makedirs(tmp_path / "hub" / "checkpoints")
    os.replace(path, tmp_path / "hub" / "checkpoints" / "resnet18-5c106cde.pth")
    assert M.Trunk().load_imagenet_weights() is True
    monkeypatch.setenv("DAHITRA_RESNET18_CKPT", str(tmp_path / "missing.pth"))
    with pytest.raises(FileNotFoundError):
        M.Trunk()  .load_imagenet_weights()


def test_graphed_train_step_eager_form_on_cpu_and_shape_guard():
    """GraphedTrainStep with use_graph=False (what runs without a GPU): the flat-buffer step equals the plain loop bit for bit, unused
    parameters are frozen, and a batch of another shape is refused instead of being broadcast into the static buffers"""
    import copy
    import torch.nn.functional as F
    from dahitra_b200.train_graph import GraphedTrainStep

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Conv2d(3, 8, 3, padding=1)
            self.bn = torch.nn.BatchNorm2d(8)
            self.b = torch.nn.Conv2d(8, 2, 1)
            self.unused = torch.nn.Linear(4, 4)

        def forward(self, x):
            return self.b(F.relu(self.bn(self.a(x))))

    torch.manual_seed(3)
    net = Net()
    ref = copy.deepcopy(net).train()
    mk = lambda: (torch.randn(4, 3, 8, 8), torch.randint(0, 2, (4, 8, 8)))          # noqa: E731
    batches = [mk() for _ in range(4)]
    mk_opt = lambda ps: torch.optim.AdamW(ps, lr=1e-2, weight_decay=0.01)            # noqa: E731
    ts = GraphedTrainStep(net, F.cross_entropy, batches[0], mk_opt, use_graph=False)
    assert not ts.use_graph and len(ts.frozen) == 2 and ts.flat.numel() == sum(p.numel() for p in ts.live)
    opt = mk_opt([p for n, p in ref.named_parameters() if not n.startswith("unused")])
    for b in batches[1:]:
        l1 = float(ts.step(*b))
        opt.zero_grad(set_to_none=True)
        l2 = F.cross_entropy(ref(b[0]), b[1])
        l2.backward()
        opt.step()
        assert l1 == float(l2.detach())
    for (k, v), w in zip(net.state_dict().items(), ref.state_dict().values()):
        assert torch.equal(v, w), k
    with pytest.raises(ValueError, match="built for"):
        ts.step(batches[0][0][:1], batches[0][1][:1])
    with pytest.raises(ValueError, match="expected 2 tensors"):
        ts.step(batches[0][0])
