"""GPU: the native training kernels (csrc/train_decoder.cu behind dahitra_b200/training.py) against fp64 autograd of the same
function, and the whole training step's gradients against the stock-autograd route (reference
models/trainer.py:247-262, models/help_funcs.py:66-114,170-186)."""
import pytest
import torch
import torch.nn.functional as F

import emulate as E
from dahitra_b200 import modules as M
from dahitra_b200 import training as T
from dahitra_b200.networks import define_G

pytestmark = pytest.mark.gpu
DEV = "cuda"


class Args:
    net_G = "newUNetTrans"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("heads,depth,B,N", [(4, 4, 3, 256), (4, 1, 2, 1024), (8, 8, 2, 4096), (8, 2, 1, 300), (4, 2, 2, 77)])
def test_pixel_decoder_train_kernels(heads, depth, B, N):
    """forward, dL/dx and the gradient of every table entry against fp64 autograd of the same algebra (ragged N included)"""
    g = torch.Generator().manual_seed(heads * 100 + N)
    x = torch.randn(B, 32, N, generator=g)
    tab = torch.randn(B, depth, T.train_tab_floats(heads), generator=g) * 0.15
    w = torch.randn(B, 32, N, generator=g)
    xd, td = x.double().requires_grad_(), tab.double().requires_grad_()
    yd = E.train_decoder_from_tables(xd, td, heads)
    (yd * w.double()).sum().backward()
    xg, tg = x.to(DEV).requires_grad_(), tab.to(DEV).requires_grad_()
    y = T.pixel_decoder(xg, tg, heads)
    (y * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    e = (rel(y, yd), rel(xg.grad, xd.grad), rel(tg.grad, td.grad))
    print(f"[train decoder] heads {heads} depth {depth} B {B} N {N}: rel err out {e[0]:.2e} dx {e[1]:.2e} dtables {e[2]:.2e}")
    assert e[0] < 2e-6 and e[1] < 1e-5 and e[2] < 1e-5, e
    # per-entry check of the table gradient (a wrong row / column order hides in a max-norm)
    d = (tg.grad.double().cpu() - td.grad).abs()
    assert bool((d <= 1e-5 * td.grad.abs().max() + 1e-4 * td.grad.abs()).all())
    # channels_last tensors are read pixel-major without a copy: same arithmetic, same bits, same memory format out
    h = 16 if N % 16 == 0 else 1
    if h > 1:
        xc = x.view(B, 32, h, N // h).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_()
        tc = tab.to(DEV).requires_grad_()
        yc = T.pixel_decoder(xc, tc, heads)
        assert yc.is_contiguous(memory_format=torch.channels_last) and torch.equal(yc.flatten(2), y)
        (yc * w.to(DEV).view_as(yc).contiguous(memory_format=torch.channels_last)).sum().backward()
        assert torch.equal(xc.grad.flatten(2), xg.grad) and torch.equal(tc.grad, tg.grad)
    # deterministic: a second backward gives the same bits
    xg2, tg2 = x.to(DEV).requires_grad_(), tab.to(DEV).requires_grad_()
    (T.pixel_decoder(xg2, tg2, heads) * w.to(DEV)).sum().backward()
    assert torch.equal(xg2.grad, xg.grad) and torch.equal(tg2.grad, tg.grad)


@pytest.mark.parametrize("B,N", [(3, 256), (2, 1024), (2, 4096), (2, 300), (1, 77), (2, 16384)])
def test_semantic_tokens_train_kernels(B, N):
    """tokenizer forward, dL/dx and dL/dW against fp64 autograd of reference models/networks.py:1273-1280 (ragged N included;
    logits scaled so that the softmax over the pixels is far from uniform)"""
    g = torch.Generator().manual_seed(N)
    x = torch.randn(B, 32, N, generator=g).clamp_min(0)            # post-ReLU features
    w = torch.randn(4, 32, 1, 1, generator=g) * 0.5
    dt = torch.randn(B, 4, 32, generator=g)
    xd, wd = x.double().requires_grad_(), w.double().requires_grad_()
    a = torch.einsum("lc,bcn->bln", wd.view(4, 32), xd).softmax(-1)
    tokd = torch.einsum("bln,bcn->blc", a, xd)
    (tokd * dt.double()).sum().backward()
    xg, wg = x.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    tok = T.semantic_tokens(xg, wg)
    (tok * dt.to(DEV)).sum().backward()
    e = (rel(tok, tokd), rel(xg.grad, xd.grad), rel(wg.grad, wd.grad))
    print(f"[train tokens] B {B} N {N}: rel err tokens {e[0]:.2e} dx {e[1]:.2e} dW {e[2]:.2e}")
    assert e[0] < 2e-6 and e[1] < 1e-5 and e[2] < 1e-5, e
    h = 16 if N % 16 == 0 else 1
    if h > 1:                                                      # channels_last input: same bits
        xc = x.view(B, 32, h, N // h).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_()
        wc = w.to(DEV).requires_grad_()
        tc = T.semantic_tokens(xc, wc)
        assert torch.equal(tc, tok)
        (tc * dt.to(DEV)).sum().backward()
        assert xc.grad.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(xc.grad.flatten(2), xg.grad) and torch.equal(wc.grad, wg.grad)


def test_pixel_decoder_module_native_vs_stock():
    """PixelDecoder parameters + tokens: gradients through train_tables + the native kernels equal the as-written module's"""
    torch.manual_seed(5)
    dec = M.PixelDecoder(32, 4, 4, 64, 32).to(DEV)
    for p in dec.parameters():
        p.data.add_(0.1 * torch.randn_like(p))
    x = torch.randn(2, 32, 512, device=DEV)
    m = torch.randn(2, 4, 32, device=DEV)
    w = torch.randn(2, 32, 512, device=DEV)
    res = []
    for kind in ("fp64", "native", "stock"):
        d = dec.double() if kind == "fp64" else dec.float()
        dt = torch.float64 if kind == "fp64" else torch.float32
        xx, mm = x.to(dt).requires_grad_(), m.to(dt).requires_grad_()
        for p in d.parameters():
            p.grad = None
        if kind == "native":
            y = T.pixel_decoder(xx, d.train_tables(mm), d.heads)
        else:
            y = d(xx.transpose(1, 2), mm).transpose(1, 2)
        (y * w.to(dt)).sum().backward()
        res.append([y.detach()] + [t.grad.clone() for t in list(d.parameters()) + [xx, mm]])
    worst = 0.0
    for ref, nat, stk in zip(*res):
        en, es = rel(nat, ref), rel(stk, ref)
        worst = max(worst, en)
        assert en <= max(2e-5, 3 * es), (en, es, tuple(ref.shape))
    print(f"[train decoder] module gradients: worst rel err vs fp64 {worst:.2e}")


def _same_weights(net, ref, label):
    """weights / BatchNorm statistics of two training runs that differ only in HOW the same kernels were issued (graph replay vs
    eager).  Bit-identical when cuDNN picks the same algorithms in both (the usual case); otherwise fp32 rounding of a few
    steps: |a - b| <= 1e-3 max|b| + 1e-6."""
    worst = 0.0
    for (n, a), b in zip(net.state_dict().items(), ref.state_dict().values()):
        if a.dtype.is_floating_point:
            d = float((a.double() - b.double()).abs().max())
            worst = max(worst, d / max(float(b.abs().max()), 1e-30))
            assert d <= 1e-3 * float(b.abs().max()) + 1e-6, (n, d, float(b.abs().max()))
        else:
            assert torch.equal(a, b), n                         # num_batches_tracked: warm-up / capture passes left no trace
    print(f"[{label}] worst relative weight difference: {worst:.2e}")


def _within(en, es):
    """native error vs the stock route's error, both against fp64"""
    return en <= max(2e-4, 3 * es)


def _compare_with_fp64(gn, gs, g64, label):
    """Per-tensor max-norm relative error of the native-route and stock-route fp32 gradients against the fp64 gradients.
    A handful of tensors are dominated by ReLU / max-pool decisions that flip between fp32 and fp64 (discrete noise of a few
    1e-2, different for every route and seed), so the bar is statistical: the MEDIAN error of the native route is within 2x of
    the stock route's, at most 10 % of the tensors are more than 3x worse than stock, and none is off by more than 0.2."""
    errs = sorted(((rel(gn[k], g64[k]), rel(gs[k], g64[k]), k) for k in g64), reverse=True)
    en = sorted(e[0] for e in errs)
    es = sorted(e[1] for e in errs)
    med_n, med_s = en[len(en) // 2], es[len(es) // 2]
    out = [(k, a, b) for a, b, k in errs if not _within(a, b)]
    dec = [(a, b, k) for a, b, k in errs if "transformer_decoder" in k or "conv_token" in k]
    print(f"[{label}] {len(errs)} gradients vs fp64: native median {med_n:.2e} worst {en[-1]:.2e} ({errs[0][2]}); stock median {med_s:.2e} "
          f"worst {es[-1]:.2e}; {len(out)} tensors > 3x stock; decoder / tokenizer parameters ({len(dec)}): native worst {dec[0][0]:.2e}, "
          f"stock worst {max(d[1] for d in dec):.2e}")
    for k, a, b in out[:5]:
        print(f"[{label}]   > 3x stock: {k}: native {a:.2e} stock {b:.2e}")
    assert med_n <= max(5e-5, 2 * med_s), (med_n, med_s)
    assert len(out) <= 0.10 * len(errs), out[:8]
    assert en[-1] <= 0.2, errs[0]


def _grads(net, x1, x2, y, native, paired=None):
    net.native_training = native            # stock = torch ops for decoder / tokenizer and the reference's two trunk passes
    net.paired_trunk_training = native if paired is None else paired
    for p in net.parameters():
        p.grad = None
    loss = F.cross_entropy(net(x1, x2) if x2 is not None else net(x1), y)
    loss.backward()
    return float(loss), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}


def _training_routes(net, x1, x2, y, label):
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    d = lambda t: None if t is None else t.double()          # noqa: E731
    ln, gn = _grads(net, x1, x2, y, True)                    # the default training route: native kernels, paired trunk
    bufs_n = {k: v.clone() for k, v in net.named_buffers()}
    net.load_state_dict(sd0)                                 # BN running stats moved; same starting point for every route
    ls, gs = _grads(net, x1, x2, y, False)
    # BatchNorm running statistics / num_batches_tracked after one step: the paired trunk updates them per image set in the
    # reference's order (models/resnet.py:57-73 called twice per step)
    for k, v in net.named_buffers():
        assert torch.allclose(bufs_n[k].double(), v.double(), rtol=1e-5, atol=1e-7), k
    assert int(net.resnet.bn1.num_batches_tracked) == int(sd0["resnet.bn1.num_batches_tracked"]) + 2
    net.load_state_dict(sd0)
    net.double()
    l64, g64 = _grads(net, d(x1), d(x2), y, False)
    net.load_state_dict(sd0)
    l64p, g64p = _grads(net, d(x1), d(x2), y, False, paired=True)
    # the paired trunk is the same function: in fp64 (no decision flips) every gradient agrees to rounding
    assert abs(l64 - l64p) <= 1e-12 * abs(l64)
    worst = max(rel(g64p[k], g64[k]) for k in g64)
    print(f"[{label}] paired trunk vs two passes in fp64: worst rel difference {worst:.2e}")
    assert worst <= 1e-8
    assert set(gn) == set(gs) == set(g64)
    assert abs(ln - l64) <= 1e-5 * abs(l64) + 1e-6 and abs(ls - l64) <= 1e-5 * abs(l64) + 1e-6
    print(f"[{label}] loss native {ln:.6f} stock {ls:.6f} fp64 {l64:.6f}")
    _compare_with_fp64(gn, gs, g64, label)


def test_training_step_gradients_native_vs_stock():
    """One LEVIR training step (train mode: batch-statistics BN): loss and every parameter gradient of the default training
    route (native decoder / tokenizer kernels, paired trunk, channels_last) and of the stock-autograd route (torch ops, the
    reference's two trunk passes), both measured against the same network in fp64."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = define_G(Args(), gpu_ids=[0]).train()
    g = torch.Generator(device=DEV).manual_seed(7)
    x1 = torch.rand(2, 3, 256, 256, device=DEV, generator=g) * 2 - 1
    x2 = torch.rand(2, 3, 256, 256, device=DEV, generator=g) * 2 - 1
    y = (torch.rand(2, 256, 256, device=DEV, generator=g) < 0.2).long()
    _training_routes(net, x1, x2, y, "train step")


def test_training_step_xbd_variant_native_vs_stock():
    """the xBD variant's training route (one decoder pass per level on conv_decode's output; PyTorch-default init) the same way"""
    from dahitra_b200 import xbd
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    net = xbd.BASE_Transformer_UNet(3, 5, with_pos='learned').to(DEV).train()
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.rand(2, 6, 128, 160, device=DEV, generator=g) * 2 - 1
    y = torch.randint(0, 5, (2, 128, 160), device=DEV, generator=g)
    _training_routes(net, x, None, y, "train step xBD")


def test_training_step_in_cuda_graph_native():
    """forward + backward with the native kernels is capturable (no allocation-order or synchronisation surprises) and replays
    to the same gradients"""
    torch.manual_seed(0)
    net = define_G(Args(), gpu_ids=[0]).train()
    g = torch.Generator(device=DEV).manual_seed(9)
    x1 = torch.rand(2, 3, 256, 256, device=DEV, generator=g) * 2 - 1
    x2 = torch.rand(2, 3, 256, 256, device=DEV, generator=g) * 2 - 1
    y = (torch.rand(2, 256, 256, device=DEV, generator=g) < 0.2).long()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            F.cross_entropy(net(x1, x2), y).backward()
    torch.cuda.current_stream().wait_stream(side)
    live = [p for p in net.parameters() if p.grad is not None]
    eager = [p.grad.clone() for p in live]                   # second backward ACCUMULATED: eager = 2 x gradient
    for p in live:
        p.grad.zero_()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        F.cross_entropy(net(x1, x2), y).backward()
    for p in live:
        p.grad.zero_()
    gr.replay()
    gr.replay()
    torch.cuda.synchronize()
    for p, e in zip(live, eager):
        assert rel(p.grad, e) < 1e-3                         # BN running stats do not enter the train-mode forward


def test_graphed_train_step_matches_the_eager_loop():
    """dahitra_b200.train_graph.GraphedTrainStep on the real network: three AdamW steps replayed from the two CUDA graphs leave the
    same weights and BatchNorm statistics as the plain eager loop of models/trainer.py:247-262 on a copy of the module."""
    import copy
    from dahitra_b200.train_graph import GraphedTrainStep
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False       # PyTorch's defaults (earlier tests change them)
    torch.manual_seed(0)
    net = define_G(Args(), gpu_ids=[0]).train()
    ref = copy.deepcopy(net).train()
    g = torch.Generator(device=DEV).manual_seed(11)
    mk = lambda: (torch.rand(2, 3, 256, 256, device=DEV, generator=g) * 2 - 1, torch.rand(2, 3, 256, 256, device=DEV, generator=g) * 2 - 1,   # noqa: E731
                  (torch.rand(2, 256, 256, device=DEV, generator=g) < 0.2).long())
    batches = [mk() for _ in range(4)]
    # SGD with momentum: the weights stay a linear function of the gradients (Adam's first steps are sign-like and would turn
    # fp32 gradient noise on near-zero entries into +-lr differences); tools/train_step.py runs the same class with AdamW
    mk_opt = lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=0.01)          # noqa: E731
    ts = GraphedTrainStep(net, F.cross_entropy, batches[0], mk_opt)
    assert ts.use_graph and len(ts.frozen) == 48 and ts.flat.numel() == sum(p.numel() for p in ts.live)
    opt = mk_opt(ref.parameters())
    for b in batches[1:]:
        l1 = ts.step(*b).clone()
        opt.zero_grad(set_to_none=True)
        l2 = F.cross_entropy(ref(b[0], b[1]), b[2])
        l2.backward()
        opt.step()
        assert abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l2)), (float(l1), float(l2))
    torch.cuda.synchronize()
    _same_weights(net, ref, "graphed step")
    net.eval()
    with torch.no_grad():                                       # and the native inference path reads the trained weights
        y = net(batches[0][0], batches[0][1])
    assert torch.isfinite(y).all()


def test_graphed_route_inside_an_eager_loop():
    """net.graphed_training: the trainer's own loop (loss, backward(), optimizer.step() issued eagerly, models/trainer.py:247-262)
    with the network's forward / backward replayed from CUDA graphs gives the same weights and BatchNorm statistics as the
    plain eager route; another batch size falls back to the eager route; eval() inference is unaffected."""
    import copy
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False       # PyTorch's defaults (earlier tests change them)
    torch.manual_seed(0)
    net = define_G(Args(), gpu_ids=[0]).train()
    ref = copy.deepcopy(net).train()
    net.graphed_training = True
    g = torch.Generator(device=DEV).manual_seed(12)
    mk = lambda b: (torch.rand(b, 3, 256, 256, device=DEV, generator=g) * 2 - 1, torch.rand(b, 3, 256, 256, device=DEV, generator=g) * 2 - 1,   # noqa: E731
                    (torch.rand(b, 256, 256, device=DEV, generator=g) < 0.2).long())
    opts = [torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=0.01) for m in (net, ref)]
    for bsz in (2, 2, 1, 2):                                     # the third batch is smaller: eager fallback
        b = mk(bsz)
        losses = []
        for m, opt in zip((net, ref), opts):
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(m(b[0], b[1]), b[2])
            loss.backward()
            opt.step()
            losses.append(float(loss))
        assert abs(losses[0] - losses[1]) <= 1e-4 * abs(losses[1]), losses
    assert net.__dict__.get("_graphed_route") is not None and net._graphed_route.key[0][0] == (2, 3, 256, 256)
    _same_weights(net, ref, "graphed route")
    net.eval()
    with torch.no_grad():
        y = net(b[0], b[1])
    assert torch.isfinite(y).all() and copy.deepcopy(net).__dict__.get("_graphed_route") is None
