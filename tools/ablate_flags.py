"""Error of every shipped precision mode (dahitra_b200.engine.MODES) against the fp64 oracle.  B=2, U(-1,1) inputs; weights:
define_G init (seed 0) and the ill-conditioned default-scale synthetic set (dahitra_b200/synth.py seed 3).  Prints
max / mean |d| and the strict-tolerance (1e-4 + 1e-3 |ref|) violations per mode.
   python tools/ablate_flags.py > profiles/r02_precision_ablation.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dahitra_b200.networks import define_G          # noqa: E402
from oracle import dahitra_oracle as O              # noqa: E402
from dahitra_b200 import synth                      # noqa: E402
from dahitra_b200.engine import MODES               # noqa: E402


class A:
    net_G = "newUNetTrans"


torch.manual_seed(0)
net = define_G(A(), gpu_ids=[0]).eval()
x1, x2 = synth.synth_pair(2, 256, 256, seed=2, kind="uniform")
FLAGSETS = list(MODES.items())
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
    for mode, flags in FLAGSETS:
        net._engine.flags = flags
        with torch.no_grad():
            y = net(x1.cuda(), x2.cuda()).double().cpu()
        d = (y - ref).abs()
        bad = int((d > 1e-4 + 1e-3 * ref.abs()).sum())
        agree = float((y.argmax(1) == ref.argmax(1)).float().mean())
        tag = mode
        print(f"weights={wname} flags={flags:5d} {tag:18s} max|d|={float(d.max()):.3e} mean|d|={float(d.mean()):.3e} "
              f"outside-strict={bad}/{d.numel()} argmax_agree={agree:.6f}", flush=True)
