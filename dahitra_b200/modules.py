"""Parameter containers for the drop-in ``newUNetTrans`` network.

These classes exist to hold tensors under EXACTLY the state_dict keys, shapes,
registration order and RNG-consumption order of the reference, so that

* checkpoints written by the reference trainer load with ``strict=True``
  (reference: models/evaluator.py:71-73, models/trainer.py:106-134), and
* ``torch.manual_seed(s); define_G(...)`` yields bit-identical weights
  (reference: models/networks.py:77-127 ``init_weights`` matches on the class
  names 'Conv' / 'Linear' / 'BatchNorm2d' during ``net.apply``).

None of them carries the arithmetic of the inference path: the native forward in
``dahitra_b200.networks`` reads the tensors and launches sm_100a kernels.  The
small ``forward`` methods below exist for the autograd (training) route: its
convolutions / BatchNorm are stock PyTorch, its pixel decoders and tokenizer run
on the native training kernels through ``PixelDecoder.train_tables`` and
``dahitra_b200.training`` (see DESIGN.md "Training step").

Key layout that is mirrored (reference file:line):
  * ResNet-18 trunk           models/resnet.py:125-204 (conv1, bn1, layer1..4, fc)
  * token encoder             models/networks.py:434-512 (Residual/PreNorm/Attention/FeedForward)
  * pixel decoder             models/help_funcs.py:26-31,43-49,66-114,170-186
  * TwoLayerConv2d            models/help_funcs.py:7-15
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- trunk
class TrunkBlock(nn.Module):
    """ResNet BasicBlock container: keys conv1/bn1/conv2/bn2/downsample.{0,1}."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        # The 1x1 projection (if any) is attached by Trunk._stage, which creates it
        # BEFORE this block's convs (RNG order of models/resnet.py:188-197) but
        # registers it after bn2 (key order of BasicBlock, models/resnet.py:47-55).
        self.downsample = None
        self.stride = stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + idt)


class Trunk(nn.Module):
    """ResNet-18 as the reference instantiates it (stride-2 layer2, stride-1
    layer3/4 because of ``replace_stride_with_dilation=[False, True, True]`` with
    dilation forced back to 1 in BasicBlock; models/resnet.py:45-46,182-204)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = self._stage(64, 64, 1)
        self.layer2 = self._stage(64, 128, 2)
        self.layer3 = self._stage(128, 256, 1)   # "dilated" => stride 1
        self.layer4 = self._stage(256, 512, 1)   # never evaluated; checkpoint ballast
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, 1000)
        # same post-construction init sweep as models/resnet.py:162-167
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    IMAGENET_FILE = "resnet18-5c106cde.pth"     # what model_urls['resnet18'] names, reference models/resnet.py:11-12

    def load_imagenet_weights(self, path=None) -> bool:
        """The reference builds its trunk as ``resnet18(pretrained=True)`` (models/networks.py:1096 -> models/resnet.py:228-233:
        ``model.load_state_dict(load_state_dict_from_url(...))``).  ``init_weights`` then re-draws every Conv / BatchNorm affine
        (networks.py:88-105), so what survives ``define_G`` are the ImageNet BatchNorm running statistics (and layer4 / fc, which no
        forward reads).  This loads the same file, strictly, WITHOUT ever downloading: ``path``, else ``$DAHITRA_RESNET18_CKPT``, else
        the torch hub cache the reference's own call would have filled.  Returns False (fresh statistics: mean 0, var 1) when no
        file is there (or ``DAHITRA_RESNET18_CKPT=none``); a path that was asked for explicitly and does not exist is an error."""
        import os
        asked = path or os.environ.get("DAHITRA_RESNET18_CKPT")
        if asked is not None and str(asked).lower() in ("", "0", "none", "off"):
            return False                                     # DAHITRA_RESNET18_CKPT=none: never look (fresh statistics)
        if asked:
            if not os.path.exists(asked):
                raise FileNotFoundError(f"dahitra_b200: resnet18 checkpoint {asked} not found")
            path = asked
        else:
            path = os.path.join(torch.hub.get_dir(), "checkpoints", self.IMAGENET_FILE)
            if not os.path.exists(path):
                return False
        self.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
        return True

    @staticmethod
    def _stage(cin: int, cout: int, stride: int) -> nn.Sequential:
        proj = None
        if stride != 1 or cin != cout:
            proj = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(cout))
        first = TrunkBlock(cin, cout, stride)
        first.downsample = proj
        return nn.Sequential(first, TrunkBlock(cout, cout, 1))


# --------------------------------------------------------------------------- token transformer
class _Res(nn.Module):
    """y = fn(x, *ctx) + x  (key: fn)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *ctx):
        return self.fn(x, *ctx) + x


class _Norm(nn.Module):
    """fn(LN(x), LN(ctx)...) with ONE LayerNorm shared by all inputs (keys: norm, fn).
    reference: models/help_funcs.py:35-49."""

    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, *ctx):
        return self.fn(self.norm(x), *[self.norm(c) for c in ctx])


class _Mlp(nn.Module):
    """Linear-GELU(erf)-Linear under keys net.0 / net.3."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(0.0),
                                 nn.Linear(hidden, dim), nn.Dropout(0.0))

    def forward(self, x):
        return self.net(x)


def _split_heads(t, h):
    b, n, _ = t.shape
    return t.view(b, n, h, -1).transpose(1, 2)


class TokenSelfAttn(nn.Module):
    """keys: to_qkv.weight, to_out.0.{weight,bias}; scale = dim**-0.5."""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        self.heads, self.scale = heads, dim ** -0.5
        self.to_qkv = nn.Linear(dim, heads * dim_head * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(heads * dim_head, dim), nn.Dropout(0.0))

    def forward(self, x):
        q, k, v = (_split_heads(t, self.heads) for t in self.to_qkv(x).chunk(3, dim=-1))
        p = (q @ k.transpose(-1, -2) * self.scale).softmax(-1)
        o = (p @ v).transpose(1, 2).flatten(2)
        return self.to_out(o)


class PixelCrossAttn(nn.Module):
    """keys: to_q/to_k/to_v.weight, to_out.0.{weight,bias}."""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        self.heads, self.scale = heads, dim ** -0.5
        inner = heads * dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(0.0))

    def forward(self, x, m):
        q, k, v = (_split_heads(t, self.heads) for t in (self.to_q(x), self.to_k(m), self.to_v(m)))
        p = (q @ k.transpose(-1, -2) * self.scale).softmax(-1)
        o = (p @ v).transpose(1, 2).flatten(2)
        return self.to_out(o)


class TokenEncoder(nn.Module):
    """keys: layers.L.0.fn.{norm,fn.*}, layers.L.1.fn.{norm,fn.net.*}."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                _Res(_Norm(dim, TokenSelfAttn(dim, heads, dim_head))),
                _Res(_Norm(dim, _Mlp(dim, mlp_dim)))]))

    def forward(self, x):
        for attn, ff in self.layers:
            x = ff(attn(x))
        return x


class PixelDecoder(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.heads = heads
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                _Res(_Norm(dim, PixelCrossAttn(dim, heads, dim_head))),
                _Res(_Norm(dim, _Mlp(dim, mlp_dim)))]))

    def forward(self, x, m):
        for attn, ff in self.layers:
            x = ff(attn(x, m))
        return x

    def forward_collapsed(self, x, m):
        """The same function (reference models/help_funcs.py:66-114,170-186) evaluated in the collapsed algebra the native
        kernel uses, with plain differentiable torch ops — the training route's default.  The 4 memory tokens are constant
        over the pixels, so per image and head the key / value projections fold into two small matrices:
            dots = LN0(x) @ (g * A) + b @ A,   A[c, (h, j)] = dim^-0.5 sum_d Wq[hd, c] k[j, hd]
            x   += softmax_j(dots) @ Bv + b_o,  Bv[(h, j), c] = sum_d Wo[c, hd] v[j, hd]
        and the LayerNorm affine of the pixel side folds into A / W1 (LN0 = normalisation without affine), so nothing of size
        (B, N, heads * 64) is ever formed and no LayerNorm weight gradient is reduced over the B * N pixel rows: gradients
        reach Wq, Wk, Wv, Wo and the LayerNorm parameters through the small matrices.  x: (B, N, 32), m: (B, 4, 32)."""
        B, N, C = x.shape
        for attn, ff in self.layers:
            norm, ca = attn.fn.norm, attn.fn.fn
            H = ca.heads
            mn = norm(m)                                                     # (B, 4, C): PreNorm2 shares the layer's LayerNorm
            k = ca.to_k(mn).view(B, -1, H, ca.to_k.out_features // H)       # (B, 4, H, D)
            v = ca.to_v(mn).view(B, -1, H, ca.to_v.out_features // H)
            D = k.shape[-1]
            A = torch.einsum("hdc,bjhd->bchj", ca.to_q.weight.view(H, D, C), k).reshape(B, C, -1) * ca.scale    # (B, C, 4H)
            Bv = torch.einsum("chd,bjhd->bhjc", ca.to_out[0].weight.view(C, H, D), v).reshape(B, -1, C)         # (B, 4H, C)
            xn = torch.nn.functional.layer_norm(x, (C,), None, None, norm.eps)
            dots = torch.baddbmm((norm.bias @ A).unsqueeze(1), xn, norm.weight[None, :, None] * A)              # (B, N, 4H)
            p = dots.view(B, N, H, -1).softmax(-1).view(B, N, -1)
            x = x + torch.baddbmm(ca.to_out[0].bias.view(1, 1, C), p, Bv)
            norm2, l1, l2 = ff.fn.norm, ff.fn.fn.net[0], ff.fn.fn.net[3]
            xn2 = torch.nn.functional.layer_norm(x, (C,), None, None, norm2.eps)
            h = torch.addmm(l1.bias + l1.weight @ norm2.bias, xn2.reshape(B * N, C), (l1.weight * norm2.weight[None, :]).t())
            x = x + torch.addmm(l2.bias, torch.nn.functional.gelu(h), l2.weight.t()).view(B, N, C)
        return x

    def train_tables(self, m):
        """Per (image, layer) tables of the native training kernels (csrc/train_decoder.cu; layout DH_TRAIN_TAB_FLOATS of
        include/dahitra_b200.h) from the parameters and the memory tokens m (B, 4, 32), with plain differentiable torch ops on
        tensors of a few KB: the same folding as ``forward_collapsed`` (k = head * 4 + token; LayerNorm affines folded into
        A / c0 and W1 / b1).  All layers are built together (the op count does not grow with the depth).  Returns (B, depth, T)."""
        B, J, C = m.shape
        L, H = len(self.layers), self.heads
        st = lambda f: torch.stack([f(attn.fn, ff.fn) for attn, ff in self.layers])          # noqa: E731
        nw, nb = st(lambda a, f: a.norm.weight), st(lambda a, f: a.norm.bias)               # (L, C)
        wq, wk, wv = (st(lambda a, f, n=n: getattr(a.fn, n).weight) for n in ("to_q", "to_k", "to_v"))   # (L, H*D, C)
        wo, bo = st(lambda a, f: a.fn.to_out[0].weight), st(lambda a, f: a.fn.to_out[0].bias)            # (L, C, H*D), (L, C)
        n2w, n2b = st(lambda a, f: f.norm.weight), st(lambda a, f: f.norm.bias)
        w1, b1 = st(lambda a, f: f.fn.net[0].weight), st(lambda a, f: f.fn.net[0].bias)
        w2, b2 = st(lambda a, f: f.fn.net[3].weight), st(lambda a, f: f.fn.net[3].bias)
        D = wq.shape[1] // H
        eps, scale = self.layers[0][0].fn.norm.eps, self.layers[0][0].fn.fn.scale
        mh = torch.nn.functional.layer_norm(m, (C,), None, None, eps)                        # PreNorm2 shares the layer's LayerNorm
        mn = mh[None] * nw[:, None, None, :] + nb[:, None, None, :]                          # (L, B, 4, C)
        k = torch.einsum("lbjc,lec->lbje", mn, wk).view(L, B, J, H, D)
        v = torch.einsum("lbjc,lec->lbje", mn, wv).view(L, B, J, H, D)
        A = torch.einsum("lhdc,lbjhd->lbchj", wq.view(L, H, D, C), k).reshape(L, B, C, H * J) * scale
        Bv = torch.einsum("lchd,lbjhd->lbhjc", wo.view(L, C, H, D), v).reshape(L, B, H * J * C)
        c0 = torch.einsum("lc,lbck->lbk", nb, A)
        shared = torch.cat([bo, (w1 * n2w[:, None, :]).transpose(1, 2).reshape(L, C * C), b1 + torch.einsum("ljc,lc->lj", w1, n2b),
                            w2.transpose(1, 2).reshape(L, C * C), b2], dim=1)
        tab = torch.cat([(nw[:, None, :, None] * A).reshape(L, B, -1), c0, Bv, shared[:, None, :].expand(L, B, -1)], dim=2)
        return tab.transpose(0, 1).contiguous()


def two_layer_head(cin: int, cout: int) -> nn.Sequential:
    """keys 0.weight, 1.*, 3.{weight,bias}  (models/help_funcs.py:7-15)."""
    return nn.Sequential(nn.Conv2d(cin, cin, 3, padding=1, bias=False), nn.BatchNorm2d(cin),
                         nn.ReLU(), nn.Conv2d(cin, cout, 3, padding=1))
