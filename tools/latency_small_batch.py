"""Latency of configs[0] (LEVIR 256x256, batch 4) on one GPU: eager launches vs a captured CUDA graph.
   usage: python tools/latency_small_batch.py [--mode tf32x3]   -> one JSON line"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dahitra_b200.networks import define_G


class Args:
    net_G = "newUNetTrans"


ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="tf32x3")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--extra-flags", type=int, default=0, help="OR-ed into the mode's DH_FLAG_* bitmask (32768 = PDL, 256 = one stream)")
a = ap.parse_args()
torch.manual_seed(0)
from dahitra_b200.engine import MODES
net = define_G(Args(), gpu_ids=[0]).eval().set_mode(MODES[a.mode] | a.extra_flags)
x1 = torch.rand(a.batch, 3, 256, 256, device="cuda") * 2 - 1
x2 = torch.rand(a.batch, 3, 256, 256, device="cuda") * 2 - 1


def timed(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    eager_ms = timed(lambda: net(x1, x2))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        net(x1, x2)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = net(x1, x2)
    graph_ms = timed(g.replay)
print(json.dumps(dict(workload=f"LEVIR 256x256 batch {a.batch} (configs[0])", mode=a.mode, extra_flags=a.extra_flags, eager_ms=eager_ms, graph_ms=graph_ms,
                      eager_pairs_per_s=a.batch / eager_ms * 1e3, graph_pairs_per_s=a.batch / graph_ms * 1e3)))
