"""Device time of the native training decoder kernels alone (csrc/train_decoder.cu) at the shapes of one LEVIR training step,
batch 8: per level one launch for both image sets (16 images) and one for the difference features (8 images).
   python tools/bench_train_decoder.py > profiles/r02_train_decoder_kernels.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dahitra_b200 import _lib
from dahitra_b200.training import train_tab_floats

only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None      # e.g. --only "level3 pair" (for ncu captures)
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 20
lib = _lib.load()
dev = "cuda"
rows = []
for name, B, N, heads, depth in (("level5 pair", 16, 256, 4, 4), ("level5 diff", 8, 256, 4, 4), ("level4 pair", 16, 1024, 4, 4),
                                 ("level4 diff", 8, 1024, 4, 4), ("level3 pair", 16, 4096, 8, 8), ("level3 diff", 8, 4096, 8, 8)):
    if only and name != only:
        continue
    T = train_tab_floats(heads)
    x = torch.randn(B, 32, N, device=dev)
    tab = torch.randn(B, depth, T, device=dev) * 0.15
    xs = torch.empty(depth, B, 32, N, device=dev)
    out, dx = torch.empty_like(x), torch.empty_like(x)
    nblk = lib.dahitra_pixel_decoder_train_blocks(N)
    part = torch.empty(B, nblk, depth, T, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fwd = lambda: _lib.check(lib.dahitra_pixel_decoder_train_fwd(x.data_ptr(), tab.data_ptr(), xs.data_ptr(), out.data_ptr(), B, N, heads, depth, 0, st))
    bwd = lambda: _lib.check(lib.dahitra_pixel_decoder_train_bwd(out.data_ptr(), xs.data_ptr(), tab.data_ptr(), dx.data_ptr(), part.data_ptr(), B, N, heads, depth, 0, st))
    ms = []
    for fn in (fwd, bwd):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / iters)
    K = 4 * heads
    fma_fwd = B * N * depth * (2 * 32 * K + 2 * 1024)
    rows.append(dict(call=name, images=B, pixels=N, heads=heads, depth=depth, ctas=B * nblk, fwd_ms=ms[0], bwd_ms=ms[1],
                     fwd_tflops=2 * fma_fwd / ms[0] / 1e9, bwd_tflops=2 * 3 * fma_fwd / ms[1] / 1e9,
                     note="bwd FLOPs = recomputed forward + data gradient + weight gradient = 3x forward"))
print(json.dumps(dict(kernels=rows, fwd_ms_total=sum(r["fwd_ms"] for r in rows), bwd_ms_total=sum(r["bwd_ms"] for r in rows))))
