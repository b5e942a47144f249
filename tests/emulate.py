"""Test helper: torch emulation of the ALGEBRA the CUDA kernels implement, reading the prepared weight packs
exactly as the kernels do (include/dahitra_b200.h layouts).  It lets the CPU suite check the host-side
weight preparation (BN folding, collapsed attention products, LayerNorm folding, re-layouts) against the
oracle without a GPU.  Not product code and not the oracle."""
import torch
import torch.nn.functional as F

EPS = 1e-5


def ln_hat(x):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + EPS)


def gelu(x):
    return 0.5 * x * (1 + torch.erf(x * 0.7071067811865476))


def conv_nhwc(x, w_khwc, bias, KH, stride, pad, res=None, relu=False, up=1):
    """x NHWC; w [KH*KW*Cin][Cout] as the kernels read it."""
    cin = x.shape[-1]
    cout = w_khwc.shape[1]
    w = w_khwc.reshape(KH, KH, cin, cout).permute(3, 2, 0, 1)
    xx = x.permute(0, 3, 1, 2)
    if up == 2:
        xx = xx.repeat_interleave(2, 2).repeat_interleave(2, 3)
    y = F.conv2d(xx, w, bias, stride, pad).permute(0, 2, 3, 1)
    if res is not None:
        y = y + res
    return F.relu(y) if relu else y


def unpack_enc(enc, H):
    o = 0
    def take(n):
        nonlocal o
        v = enc[o:o + n]; o += n
        return v
    d = dict(pos=take(256).view(8, 32), g1=take(32), b1n=take(32), Mqk=take(H * 1024).view(H, 32, 32),
             MvoT=take(H * 1024).view(H, 32, 32), bo=take(32), g2=take(32), b2n=take(32),
             W1t=take(1024).view(32, 32), b1=take(32), W2t=take(1024).view(32, 32), b2=take(32))
    assert o == enc.numel()
    return d


def unpack_dec_layer(dec, H, layer):
    stride = 64 + 2 * H * 1024 + 32 + 1024 + 32 + 1024 + 32
    L = dec[layer * stride:(layer + 1) * stride]
    o = 0
    def take(n):
        nonlocal o
        v = L[o:o + n]; o += n
        return v
    d = dict(g1=take(32), b1n=take(32), MqkT=take(H * 1024).view(H, 32, 32), MovT=take(H * 1024).view(H, 32, 32),
             bo=take(32), W1f=take(1024).view(32, 32), b1f=take(32), W2t=take(1024).view(32, 32), b2=take(32))
    assert o == stride
    return d


def squeeze_tokens(feat, wsq, wtok, chunk=256):
    """feat [N][npix][Cin] -> xs [N][npix][32], partials [N][nchunk][4][34]"""
    xs = F.relu(feat @ wsq)
    a = xs @ wtok                                   # [N][npix][4]
    N, npix, _ = xs.shape
    parts = []
    for c0 in range(0, npix, chunk):
        ac, xc = a[:, c0:c0 + chunk], xs[:, c0:c0 + chunk]
        m = ac.max(1).values                        # [N][4]
        e = torch.exp(ac - m[:, None])              # [N][p][4]
        s = e.sum(1)
        t = torch.einsum("npl,npc->nlc", e, xc)
        parts.append(torch.cat([m[..., None], s[..., None], t], -1))
    return xs, torch.stack(parts, 1)


def token_encoder(partials, B, enc, H, add_pos):
    m, s, t = partials[..., 0], partials[..., 1], partials[..., 2:]
    M = m.max(1, keepdim=True).values
    sc = torch.exp(m - M)
    tok = (t * sc[..., None]).sum(1) / (s * sc).sum(1)[..., None]          # [2B][4][32]
    X = torch.cat([tok[:B], tok[B:]], 1)                                     # [B][8][32]
    p = unpack_enc(enc, H)
    if add_pos:
        X = X + p["pos"]
    xn = ln_hat(X) * p["g1"] + p["b1n"]
    out = torch.zeros_like(X)
    for h in range(H):
        u = xn @ p["Mqk"][h]                                                 # [B][8][c']
        dots = u @ xn.transpose(1, 2)
        att = dots.softmax(-1)
        y = xn @ p["MvoT"][h]                                                # y[j][c] = sum_c' xn[j][c'] MvoT[c'][c]
        out = out + att @ y
    X = X + out + p["bo"]
    xn2 = ln_hat(X) * p["g2"] + p["b2n"]
    X = X + gelu(xn2 @ p["W1t"] + p["b1"]) @ p["W2t"] + p["b2"]
    return torch.stack([X[:, :4], X[:, 4:], (X[:, 4:] - X[:, :4]).abs()], 1)   # mem [B][3][4][32]


def decoder_tables(mem_call, dec, H, depth):
    """mem_call [nimg][4][32] -> list over layers of (A[nimg][32][4H], cA[nimg][4H], Bv[nimg][4H][32], bo)"""
    out = []
    for l in range(depth):
        p = unpack_dec_layer(dec, H, l)
        mn = ln_hat(mem_call) * p["g1"] + p["b1n"]                           # [n][4][32]
        a = torch.einsum("hec,nje->nhjc", p["MqkT"], mn)                     # [n][h][j][c]
        v = torch.einsum("hec,nje->nhjc", p["MovT"], mn)
        n = mem_call.shape[0]
        A = (a * p["g1"]).reshape(n, 4 * H, 32).transpose(1, 2)              # [n][c][4H]
        cA = (a * p["b1n"]).sum(-1).reshape(n, 4 * H)
        out.append((A, cA, v.reshape(n, 4 * H, 32), p["bo"], p))
    return out


def pixel_decoder(x, pos, tables, H, skip=None):
    """x [nimg][npix][32]"""
    if pos is not None:
        x = x + pos
    n = x.shape[0]
    for A, cA, Bv, bo, p in tables:
        s = ln_hat(x) @ A + cA[:, None]                                      # [n][npix][4H]
        pr = s.view(n, -1, H, 4).softmax(-1).view(n, -1, 4 * H)
        x = x + pr @ Bv + bo
        x = x + gelu(ln_hat(x) @ p["W1f"] + p["b1f"]) @ p["W2t"] + p["b2"]
    if skip is not None:
        x = x + skip
    return x


def stem(x, w, b):
    """x NCHW (N,3,H,W); w [7*7*3][64] rows ordered (r, s, ci)."""
    wt = w.reshape(7, 7, 3, 64).permute(3, 2, 0, 1)
    return F.relu(F.conv2d(x, wt, b, 2, 3)).permute(0, 2, 3, 1)


def maxpool(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)


def forward(P, x1, x2, variant="levir", nc=2):
    """Same launch sequence as dahitra_forward (csrc/forward.cu), on the prepared packs P (slot name -> tensor)."""
    B = x1.shape[0]
    c = lambda x, slot, K, s, res=None, relu=False, up=1, bias=True: conv_nhwc(
        x, P[slot + "_W"], P[slot + "_B"] if bias else None, K, s, K // 2, res, relu, up)
    F2 = stem(torch.cat([x1, x2]), P["DH_W_STEM_W"], P["DH_W_STEM_B"])
    P2 = maxpool(F2)
    t = c(P2, "DH_W_L1_0_C1", 3, 1, relu=True); t = c(t, "DH_W_L1_0_C2", 3, 1, res=P2, relu=True)
    u = c(t, "DH_W_L1_1_C1", 3, 1, relu=True); F4 = c(u, "DH_W_L1_1_C2", 3, 1, res=t, relu=True)
    a = c(F4, "DH_W_L2_0_C1", 3, 2, relu=True); d = c(F4, "DH_W_L2_0_DS", 1, 2)
    t = c(a, "DH_W_L2_0_C2", 3, 1, res=d, relu=True)
    u = c(t, "DH_W_L2_1_C1", 3, 1, relu=True); F8 = c(u, "DH_W_L2_1_C2", 3, 1, res=t, relu=True)
    P8 = maxpool(F8)
    a = c(P8, "DH_W_L3_0_C1", 3, 1, relu=True); d = c(P8, "DH_W_L3_0_DS", 1, 1)
    t = c(a, "DH_W_L3_0_C2", 3, 1, res=d, relu=True)
    u = c(t, "DH_W_L3_1_C1", 3, 1, relu=True); F16 = c(u, "DH_W_L3_1_C2", 3, 1, res=t, relu=True)
    outs, C4 = {}, None
    for i, (k, feat, H, depth) in enumerate(((5, F16, 4, 4), (4, F8, 4, 4), (3, F4, 8, 8))):
        s = f"DH_W_LV{k}_"
        n, h, w, cin = feat.shape
        xs, parts = squeeze_tokens(feat.reshape(n, h * w, cin), P[s + "SQ"], P[s + "TOK"])
        add_pos = True if variant == "levir" else (k == 5)
        mem = token_encoder(parts, B, P[s + "ENC"], H, add_pos)
        pos = P.get(s + "POS")
        skip = None if i == 0 else (outs[5].repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(B, h * w, 32)
                                    if i == 1 else C4.reshape(B, h * w, 32))
        if variant == "levir":
            tabs01 = decoder_tables(torch.cat([mem[:, 0], mem[:, 1]]), P[s + "DEC"], H, depth)
            xd = pixel_decoder(xs, pos, tabs01, H).reshape(n, h, w, 32)
            src = torch.cat([xd[:B], xd[B:]], -1)
        else:
            src = torch.cat([xs[:B], xs[B:]], -1).reshape(B, h, w, 64)
        dx = conv_nhwc(src, P[s + "DECODE"], None, 3, 1, 1)
        tabs2 = decoder_tables(mem[:, 2], P[s + "DEC"], H, depth)
        outs[k] = pixel_decoder(dx.reshape(B, h * w, 32), pos, tabs2, H, skip).reshape(B, h, w, 32)
        if i == 1:
            C4 = c(outs[4], "DH_W_CL4", 3, 1, relu=True, up=2)
    C3 = c(outs[3], "DH_W_CL3", 3, 1, relu=True, up=2)
    Y20 = c(torch.cat([F2[:B], F2[B:]], -1), "DH_W_CL20A", 3, 1, relu=True)
    O2 = c(Y20, "DH_W_CL20B", 3, 1, res=C3)
    C2 = c(O2, "DH_W_CL2", 3, 1, relu=True, up=2)
    wc = P["DH_W_CLS_W"].reshape(3, 3, nc, 32).permute(2, 3, 0, 1)
    return F.conv2d(C2.permute(0, 3, 1, 2), wc, P["DH_W_CLS_B"], 1, 1)


def train_decoder_from_tables(x, tables, heads):
    """The algebra of csrc/train_decoder.cu on its table layout (DH_TRAIN_TAB_FLOATS of include/dahitra_b200.h), with
    differentiable torch ops: x (B, 32, N) channel-planar, tables (B, depth, T) -> (B, 32, N)."""
    B, C, N = x.shape
    K = 4 * heads
    v = x.transpose(1, 2)                                               # (B, N, 32)
    for l in range(tables.shape[1]):
        t = tables[:, l]
        o = 0

        def take(n, *shape):
            nonlocal o
            r = t[:, o:o + n].reshape(B, *shape)
            o += n
            return r
        A, c0, Bv, bo = take(32 * K, 32, K), take(K, 1, K), take(32 * K, K, 32), take(32, 1, 32)
        W1, b1, W2, b2 = take(1024, 32, 32), take(32, 1, 32), take(1024, 32, 32), take(32, 1, 32)
        assert o == t.shape[1]
        d = c0 + ln_hat(v) @ A
        p = d.view(B, N, heads, 4).softmax(-1).view(B, N, K)
        v = v + bo + p @ Bv
        v = v + b2 + gelu(b1 + ln_hat(v) @ W1) @ W2
    return v.transpose(1, 2)
