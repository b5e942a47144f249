"""Which tensor-core sub-path contributes how much error?  B=2, U(-1,1) inputs; weights: define_G init (seed 0)
and the ill-conditioned default-scale synthetic set (oracle/synth.py seed 3).  Prints max/mean |d| against the
fp64 oracle and the strict-tolerance violations for several flag subsets."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dahitra_b200.networks import define_G          # noqa: E402
from oracle import dahitra_oracle as O              # noqa: E402
from oracle import synth                            # noqa: E402


class A:
    net_G = "newUNetTrans"


torch.manual_seed(0)
net = define_G(A(), gpu_ids=[0]).eval()
x1, x2 = synth.synth_pair(2, 256, 256, seed=2, kind="uniform")
names = {1: "conv", 2: "conv_x3", 4: "stride2", 8: "dec", 16: "stem", 32: "dec_x3", 64: "conv_v1"}
FLAGSETS = (0, 1 | 4 | 16 | 8, 1 | 4 | 16 | 8 | 32, 1 | 2 | 8 | 32, 1 | 2 | 8 | 32 | 4, 1 | 2 | 8 | 32 | 16, 1 | 2 | 8 | 32 | 4 | 16)
for wname in ("defineG", "synth3"):
    if wname == "synth3":
        net.load_state_dict({k: v.cuda() for k, v in synth.synth_state_dict({k: v.cpu() for k, v in net.state_dict().items()},
                                                                              seed=3, style="default").items()})
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    ref = O.forward_levir(sd, x1, x2, dtype=torch.float64)
    ref32 = O.forward_levir(sd, x1, x2)
    n32 = (ref32.double() - ref).abs()
    print(f"weights={wname}: fp32 CPU oracle vs fp64: max|d|={float(n32.max()):.3e} mean|d|={float(n32.mean()):.3e} "
          f"ref_absmax={float(ref.abs().max()):.3e}", flush=True)
    for flags in FLAGSETS:
        net._engine.flags = flags
        net.invalidate_native_cache()
        with torch.no_grad():
            y = net(x1.cuda(), x2.cuda()).double().cpu()
        d = (y - ref).abs()
        bad = int((d > 1e-4 + 1e-3 * ref.abs()).sum())
        agree = float((y.argmax(1) == ref.argmax(1)).float().mean())
        tag = "+".join(v for k, v in names.items() if flags & k) or "fp32"
        print(f"weights={wname} flags={flags:2d} {tag:36s} max|d|={float(d.max()):.3e} mean|d|={float(d.mean()):.3e} "
              f"outside-strict={bad}/{d.numel()} argmax_agree={agree:.6f}", flush=True)
