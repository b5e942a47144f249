"""Where the training step (config 4, batch 8, one GPU) spends its DEVICE time: forward + backward captured in one CUDA graph
(no Python launch overhead in the number) and replayed, with pieces of the network knocked out one at a time.  The difference to
the full step is what that piece costs (forward + backward).
   python tools/train_ablate.py > profiles/r02_train_step_ablation.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dahitra_b200.networks import define_G
from dahitra_b200 import modules as M


class Args:
    net_G = "newUNetTrans"


g = torch.Generator(device="cuda").manual_seed(100)
x1 = torch.rand(8, 3, 256, 256, device="cuda", generator=g) * 2 - 1
x2 = torch.rand(8, 3, 256, 256, device="cuda", generator=g) * 2 - 1
y = (torch.rand(8, 256, 256, device="cuda", generator=g) < 0.1).long()


def measure(patch=None, native=None):
    torch.manual_seed(0)
    net = define_G(Args(), gpu_ids=[0]).train()
    if native is not None:
        net.native_training = native
        net.channels_last_training = native          # stock = round 2's earlier route (NCHW, torch ops)
        net.paired_trunk_training = native
    undo = patch(net) if patch else None

    def fb():
        for p in net.parameters():
            p.grad = None
        F.cross_entropy(net(x1, x2), y).backward()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fb()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fb()
    for _ in range(3):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    if undo:
        undo()
    return e0.elapsed_time(e1) / 10


def no_decoder(net):
    old = M.PixelDecoder.forward_collapsed
    M.PixelDecoder.forward_collapsed = lambda self, x, m: x + 0 * m.sum()
    net.native_training = False
    return lambda: setattr(M.PixelDecoder, "forward_collapsed", old)


def no_bn(net):
    old = torch.nn.BatchNorm2d.forward
    torch.nn.BatchNorm2d.forward = lambda self, x: x * self.weight.view(1, -1, 1, 1) + self.bias.view(1, -1, 1, 1)
    return lambda: setattr(torch.nn.BatchNorm2d, "forward", old)


def eval_bn(net):
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    return None


out = {"workload": "forward + backward of the LEVIR training step, batch 8, one GPU, one CUDA graph replayed (ms, device time)"}
out["full_stock_autograd"] = measure(native=False)
try:
    out["full_native_decoder"] = measure(native=True)
except Exception as e:                                    # the native training kernels are not built
    out["full_native_decoder"] = repr(e)[:200]
out["without_pixel_decoders"] = measure(no_decoder, native=False)
out["batchnorm_as_affine_only"] = measure(no_bn, native=False)
out["batchnorm_in_eval_mode"] = measure(eval_bn, native=False)
print(json.dumps(out))
