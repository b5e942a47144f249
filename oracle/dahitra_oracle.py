"""ORACLE — test infrastructure, NOT product code.

CPU restatement of the reference's bitemporal forward pass (nka77/DAHiTra,
``newUNetTrans`` = ``BASE_Transformer_UNet``), written as plain functional
PyTorch over a reference-layout ``state_dict``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file; the product (``dahitra_b200``) never does.

Pinning: ``oracle/pin_against_reference.py`` imports the real reference from
/root/reference (CPU, this container only), asserts this restatement reproduces
its logits and intermediate taps on identical weights/inputs, and writes the
fixtures under ``tests/golden/``.  The reference ships no golden vectors or
tests of its own, so parity is pinned by executing the reference.

Every function cites the reference lines it restates.  ``dtype`` lets the
tests evaluate the oracle in fp64 as a tie-breaker.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LN_EPS = 1e-5
BN_EPS = 1e-5

# per-level constants: (level name in keys, trunk channels, encoder heads, decoder heads, decoder depth)
# reference: models/networks.py:1176-1238 (heads / depths are literals there)
LEVELS = {5: dict(cin=256, heads=4, depth=4), 4: dict(cin=128, heads=4, depth=4),
          3: dict(cin=64, heads=8, depth=8)}
DIM = 32
DIM_HEAD = 64
SCALE = DIM ** -0.5          # models/help_funcs.py:71, models/networks.py:462 — dim**-0.5, not dim_head


def _c(sd, key, dtype):
    return sd[key].to(dtype)


# ----------------------------------------------------------------------------- trunk
def conv_bn(sd, x, conv, bn, stride, pad, dtype, relu):
    """conv (no bias) -> eval-mode BatchNorm -> optional ReLU.
    reference: models/resnet.py:57-73 / nn.BatchNorm2d eval semantics."""
    y = F.conv2d(x, _c(sd, conv + ".weight", dtype), None, stride, pad)
    g, b = _c(sd, bn + ".weight", dtype), _c(sd, bn + ".bias", dtype)
    m, v = _c(sd, bn + ".running_mean", dtype), _c(sd, bn + ".running_var", dtype)
    y = (y - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + BN_EPS) \
        * g[None, :, None, None] + b[None, :, None, None]
    return F.relu(y) if relu else y


def basic_block(sd, x, p, stride, dtype):
    """models/resnet.py:57-73."""
    y = conv_bn(sd, x, p + ".conv1", p + ".bn1", stride, 1, dtype, True)
    y = conv_bn(sd, y, p + ".conv2", p + ".bn2", 1, 1, dtype, False)
    if (p + ".downsample.0.weight") in sd:
        x = conv_bn(sd, x, p + ".downsample.0", p + ".downsample.1", stride, 0, dtype, False)
    return F.relu(y + x)


def trunk(sd, x, dtype=torch.float32):
    """ResNet_UNet.forward_single — models/networks.py:1118-1138.
    Returns the H/2, H/4, H/8 and H/16 features (the last one is computed at the
    second max-pool's resolution; layer3 has stride 1 and no dilation)."""
    x = x.to(dtype)
    x2 = conv_bn(sd, x, "resnet.conv1", "resnet.bn1", 2, 3, dtype, True)   # in-place ReLU => x_2 is post-ReLU
    p = F.max_pool2d(x2, 3, 2, 1)
    x4 = basic_block(sd, basic_block(sd, p, "resnet.layer1.0", 1, dtype), "resnet.layer1.1", 1, dtype)
    x8 = basic_block(sd, basic_block(sd, x4, "resnet.layer2.0", 2, dtype), "resnet.layer2.1", 1, dtype)
    p8 = F.max_pool2d(x8, 3, 2, 1)                                          # same maxpool module, :1128
    x16 = basic_block(sd, basic_block(sd, p8, "resnet.layer3.0", 1, dtype), "resnet.layer3.1", 1, dtype)
    return x2, x4, x8, x16


# ----------------------------------------------------------------------------- tokens
def squeeze(sd, x, k, dtype):
    """conv_squeeze_k: 1x1 conv no bias + ReLU — models/networks.py:1177-1184."""
    return F.relu(F.conv2d(x, _c(sd, f"conv_squeeze_{k}.0.weight", dtype)))


def semantic_tokens(sd, x, k, dtype):
    """_forward_semantic_tokens — models/networks.py:1273-1280.  x: (B,32,h,w)."""
    b, c, h, w = x.shape
    a = F.conv2d(x, _c(sd, f"conv_token_{k}.weight", dtype)).reshape(b, -1, h * w)
    a = torch.softmax(a, dim=-1)
    return torch.einsum("bln,bcn->blc", a, x.reshape(b, c, h * w))


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)            # biased, nn.LayerNorm
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


def mlp(sd, x, p, dtype):
    """FeedForward — models/help_funcs.py:52-63 (GELU is exact erf)."""
    h = gelu_erf(x @ _c(sd, p + ".net.0.weight", dtype).T + _c(sd, p + ".net.0.bias", dtype))
    return h @ _c(sd, p + ".net.3.weight", dtype).T + _c(sd, p + ".net.3.bias", dtype)


def _heads(t, h):
    b, n, _ = t.shape
    return t.reshape(b, n, h, -1).permute(0, 2, 1, 3)


def token_encoder(sd, tok, k, dtype, add_pos=True):
    """_forward_transformer + Transformer(depth=1) — models/networks.py:1282-1286, 434-512.
    tok: (B, 8, 32) = cat(tokens of image 1, tokens of image 2)."""
    heads = LEVELS[k]["heads"]
    if add_pos:
        tok = tok + _c(sd, f"pos_embedding_{k}", dtype)
    p = f"transformer_{k}.layers.0"
    xn = layer_norm(tok, _c(sd, p + ".0.fn.norm.weight", dtype), _c(sd, p + ".0.fn.norm.bias", dtype))
    qkv = xn @ _c(sd, p + ".0.fn.fn.to_qkv.weight", dtype).T
    q, kk, v = (_heads(t, heads) for t in qkv.chunk(3, dim=-1))
    att = torch.softmax(torch.einsum("bhid,bhjd->bhij", q, kk) * SCALE, dim=-1)
    o = torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).flatten(2)
    tok = tok + o @ _c(sd, p + ".0.fn.fn.to_out.0.weight", dtype).T + _c(sd, p + ".0.fn.fn.to_out.0.bias", dtype)
    xn = layer_norm(tok, _c(sd, p + ".1.fn.norm.weight", dtype), _c(sd, p + ".1.fn.norm.bias", dtype))
    return tok + mlp(sd, xn, p + ".1.fn.fn", dtype)


def pixel_decoder(sd, x, m, k, dtype, add_pos=True, pos_key=None):
    """_forward_transformer_decoder + TransformerDecoder — models/networks.py:1288-1295,
    models/help_funcs.py:66-114,170-186.  x: (B,32,h,w) queries, m: (B,4,32) memory tokens.
    The SAME LayerNorm normalises x and m inside each layer (PreNorm2)."""
    heads, depth = LEVELS[k]["heads"], LEVELS[k]["depth"]
    b, c, h, w = x.shape
    if add_pos:
        x = x + _c(sd, pos_key or f"pos_embedding_decoder_{k}", dtype)
    x = x.flatten(2).transpose(1, 2)                          # b (h w) c
    for l in range(depth):
        p = f"transformer_decoder_{k}.layers.{l}"
        nw, nb = _c(sd, p + ".0.fn.norm.weight", dtype), _c(sd, p + ".0.fn.norm.bias", dtype)
        xn, mn = layer_norm(x, nw, nb), layer_norm(m, nw, nb)
        q = _heads(xn @ _c(sd, p + ".0.fn.fn.to_q.weight", dtype).T, heads)
        kk = _heads(mn @ _c(sd, p + ".0.fn.fn.to_k.weight", dtype).T, heads)
        v = _heads(mn @ _c(sd, p + ".0.fn.fn.to_v.weight", dtype).T, heads)
        att = torch.softmax(torch.einsum("bhid,bhjd->bhij", q, kk) * SCALE, dim=-1)
        o = torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).flatten(2)
        x = x + o @ _c(sd, p + ".0.fn.fn.to_out.0.weight", dtype).T + _c(sd, p + ".0.fn.fn.to_out.0.bias", dtype)
        xn = layer_norm(x, _c(sd, p + ".1.fn.norm.weight", dtype), _c(sd, p + ".1.fn.norm.bias", dtype))
        x = x + mlp(sd, xn, p + ".1.fn.fn", dtype)
    return x.transpose(1, 2).reshape(b, c, h, w)


def trans_module(sd, f1, f2, k, dtype, variant="levir", taps=None):
    """_forward_trans_module.
    levir: models/networks.py:1297-1318 (3 decoder passes per level, pos-emb on every level)
    xbd  : xBD_code/zoo/model_transformer_encoding.py:385-406 (1 decoder pass on conv_decode of the
           SQUEEZE outputs; positional terms only where the ``layer`` index equals 3, i.e. on the H/16
           level, using pos_embedding_3 / pos_embedding_decoder_3 — :358-383)."""
    x1, x2 = squeeze(sd, f1, k, dtype), squeeze(sd, f2, k, dtype)
    t1, t2 = semantic_tokens(sd, x1, k, dtype), semantic_tokens(sd, x2, k, dtype)
    tok = torch.cat([t1, t2], dim=1)
    if variant == "levir":
        tok = token_encoder(sd, tok, k, dtype, add_pos=True)
    else:
        if k == 5:   # layer index 3 -> pos_embedding_3 added to the level-5 tokens
            tok = tok + _c(sd, "pos_embedding_3", dtype)
        tok = token_encoder(sd, tok, k, dtype, add_pos=False)
    t1, t2 = tok.chunk(2, dim=1)
    if taps is not None:
        taps[f"tokens_{k}"] = tok
    if variant == "levir":
        x1 = pixel_decoder(sd, x1, t1, k, dtype)
        x2 = pixel_decoder(sd, x2, t2, k, dtype)
    dtok = (t2 - t1).abs()
    dx = F.conv2d(torch.cat([x1, x2], dim=1), _c(sd, f"conv_decode_{k}.weight", dtype), None, 1, 1)
    if variant == "levir":
        out = pixel_decoder(sd, dx, dtok, k, dtype)
    else:
        out = pixel_decoder(sd, dx, dtok, k, dtype, add_pos=(k == 5), pos_key="pos_embedding_decoder_3")
    if taps is not None:
        taps[f"level_{k}"] = out
    return out


def up2(x):
    """nn.Upsample(scale_factor=2) nearest — models/networks.py:1102."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def conv_bias_relu(sd, x, p, dtype, relu=True):
    y = F.conv2d(x, _c(sd, p + ".weight", dtype), _c(sd, p + ".bias", dtype), 1, 1)
    return F.relu(y) if relu else y


def head(sd, feats1, feats2, dtype, variant="levir", taps=None):
    """BASE_Transformer_UNet.forward after the trunk — models/networks.py:1326-1357
    (xBD: model_transformer_encoding.py:415-449, same chain)."""
    a128, a64, a32, a16 = feats1
    b128, b64, b32, b16 = feats2
    o5 = up2(trans_module(sd, a16, b16, 5, dtype, variant, taps))
    o4 = trans_module(sd, a32, b32, 4, dtype, variant, taps) + o5
    o4 = conv_bias_relu(sd, up2(o4), "conv_layer4.0", dtype)
    o3 = trans_module(sd, a64, b64, 3, dtype, variant, taps) + o4
    o3 = conv_bias_relu(sd, up2(o3), "conv_layer3.0", dtype)
    x = torch.cat([a128, b128], dim=1)
    y = conv_bn(sd, x, "conv_layer2_0.0", "conv_layer2_0.1", 1, 1, dtype, True)   # help_funcs.py:7-15
    o2 = conv_bias_relu(sd, y, "conv_layer2_0.3", dtype, relu=False) + o3
    o2 = conv_bias_relu(sd, up2(o2), "conv_layer2.0", dtype)
    if taps is not None:
        taps["out_2"] = o2
    return conv_bias_relu(sd, o2, "classifier", dtype, relu=False)


def forward_levir(sd, x1, x2, dtype=torch.float32, taps=None):
    """BASE_Transformer_UNet.forward(x1, x2) — models/networks.py:1321-1357."""
    with torch.no_grad():
        f1, f2 = trunk(sd, x1, dtype), trunk(sd, x2, dtype)
        if taps is not None:
            taps["x16_a"] = f1[3]
        return head(sd, f1, f2, dtype, "levir", taps)


def forward_xbd(sd, x, dtype=torch.float32, taps=None):
    """xBD BASE_Transformer_UNet.forward(x) — model_transformer_encoding.py:409-449."""
    with torch.no_grad():
        f1, f2 = trunk(sd, x[:, :3], dtype), trunk(sd, x[:, 3:], dtype)
        return head(sd, f1, f2, dtype, "xbd", taps)
