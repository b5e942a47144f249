#!/usr/bin/env python
"""bench.py — image-pairs/sec of the newUNetTrans bitemporal forward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload levir256|xbd1024]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one forward over one batch of synthetic image pairs (weights: torch.manual_seed(0) + define_G
random init, exactly what the reference builds).  Workload = BASELINE.json configs[1]: LEVIR 256x256, 64 pairs
per GPU per step, sharded by pair with no collective (weak scaling: every rank runs 64 pairs).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "levir256": dict(H=256, W=256, pairs=64, nc=2, variant="levir",
                     desc="newUNetTrans LEVIR-CD 256x256 batched inference, 64 pairs per GPU per step"),
    "xbd1024": dict(H=1024, W=1024, pairs=8, nc=5, variant="xbd",
                    desc="xBD 1024x1024 pre/post pair 5-class forward, 8 pairs per GPU per step"),
}
DEFAULT_MODE = "tf32x3"     # == dahitra_b200.engine.DEFAULT_MODE: what a user of the module gets without configuration
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0)


class Args:
    net_G = "newUNetTrans"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                        bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"]))), "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc, self.path = None, f"/tmp/dahitra_clocks_{os.getpid()}.csv"
        try:
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            sel = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            self.cmd = ["nvidia-smi", "-i", sel, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"]
        except Exception:
            self.cmd = None

    def start(self):
        if self.cmd is None:
            return
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(self.cmd, stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.25)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- CPU legs
def cpu_port_pairs_per_sec(wl, pairs, steps, warmup):
    """The oracle port of the reference forward on the host cores (all threads).  Checker/baseline only."""
    from oracle import dahitra_oracle as O
    from oracle import synth
    from dahitra_b200.networks import define_G
    from dahitra_b200.xbd import BASE_Transformer_UNet as XNet
    torch.manual_seed(0)
    if wl["variant"] == "levir":
        sd = define_G(Args(), gpu_ids=[]).state_dict()
    else:
        sd = XNet(input_nc=3, output_nc=wl["nc"], token_len=4, resnet_stages_num=4, with_pos="learned",
                  with_decoder_pos="learned", enc_depth=1, dec_depth=8).state_dict()
    sd = {k: v.detach() for k, v in sd.items()}
    x1, x2 = synth.synth_pair(pairs, wl["H"], wl["W"], seed=1, kind="uniform")
    fn = (lambda: O.forward_levir(sd, x1, x2)) if wl["variant"] == "levir" else (lambda: O.forward_xbd(sd, torch.cat([x1, x2], 1)))
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    sec = sum(ts) / len(ts)
    return pairs / sec, sec


def run_reference(a, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    pairs = 4 if wl["variant"] == "levir" else 1
    v, sec = cpu_port_pairs_per_sec(wl, pairs, a.steps, min(a.warmup, 3))
    sample = f"{pairs} pairs per step of the {wl['pairs']}-pair workload, {a.steps} steps, torch CPU fp32, {torch.get_num_threads()} threads"
    line = dict(impl="reference", metric="image-pairs/sec", value=v, unit="pairs/s", n_gpus=a.gpus, steps=a.steps,
                warmup=a.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=wl["desc"], H=wl["H"], W=wl["W"], pairs_per_gpu=wl["pairs"], sample_pairs=pairs),
                cpu_baseline=dict(value=v, unit="pairs/s", cores=torch.get_num_threads(), kind="port", sample=sample),
                e2e=dict(value=v, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- native arm
def run_native(a, wl):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from dahitra_b200.networks import define_G
    from dahitra_b200.xbd import BASE_Transformer_UNet as XNet
    from oracle import synth   # input generator only (synthetic data)

    H, W, Bp = wl["H"], wl["W"], (a.pairs or wl["pairs"])
    torch.manual_seed(0)
    if wl["variant"] == "levir":
        net = define_G(Args(), gpu_ids=[local]).eval()
    else:
        net = XNet(input_nc=3, output_nc=wl["nc"], token_len=4, resnet_stages_num=4, with_pos="learned",
                   with_decoder_pos="learned", enc_depth=1, dec_depth=8).to(dev).eval()
    from dahitra_b200.engine import MODES
    net.set_mode(a.flags if a.flags is not None else a.mode)
    mode_name = next((k for k, v in MODES.items() if v == net._engine.flags), f"flags{net._engine.flags}")
    # rotating input sets so consecutive steps never re-read the same inputs from L2 (3 x 100 MB > 126 MB L2;
    # the ~3.5 GB of per-step intermediates stream through HBM regardless)
    nsets = 3
    sets = []
    for i in range(nsets):
        x1, x2 = synth.synth_pair(Bp, H, W, seed=100 + 10 * rank + i, kind="uniform")
        sets.append((x1.to(dev), x2.to(dev)))

    def step(i):
        x1, x2 = sets[i % nsets]
        if wl["variant"] == "levir":
            return net(x1, x2)
        return net._engine.forward_pair(net, x1, x2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(a.warmup, 3)):
            step(i)
        barrier()
        sampler = ClockSampler(dev) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(a.steps):
            step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        # ---- e2e: host (pinned) inputs -> H2D -> forward -> argmax -> D2H of the class map, every step
        hx = [(s[0].cpu().pin_memory(), s[1].cpu().pin_memory()) for s in sets]
        pred_h = torch.empty((Bp, H, W), dtype=torch.int64).pin_memory()

        def e2e_step(i):
            h1, h2 = hx[i % nsets]
            d1, d2 = h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)
            y = net(d1, d2) if wl["variant"] == "levir" else net._engine.forward_pair(net, d1, d2)
            pred_h.copy_(torch.argmax(y, dim=1), non_blocking=True)   # what models/evaluator.py:89-103 moves to the host
            torch.cuda.current_stream().synchronize()
        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            e2e_step(i)
        barrier()
        e2e_sync_sec = time.perf_counter() - t0
        # same end-to-end work through the pipelined public API (dahitra_b200.pipeline.PairPipeline): the upload of
        # batch i+1 overlaps the forward of batch i; the class map comes back as the fused uint8 argmax
        e2e_sec, e2e_d2h = e2e_sync_sec, Bp * H * W * 8
        if True:                                        # both variants: the engine takes the pre / post tensors separately
            from dahitra_b200.pipeline import PairPipeline
            pipe = PairPipeline(net, out="argmax_u8")
            for _ in pipe.run(hx[i % nsets] for i in range(3)):
                pass
            barrier()
            # steady state of one pipe.run over K + 3 batches: the clock starts when result #3 is handed out and stops at
            # result #K+3, so K uploads, K forwards and K downloads lie inside it; the run including the pipeline fill
            # (first upload not overlapped) is reported next to it
            t_fill = time.perf_counter()
            nres = 0
            for pred in pipe.run(hx[i % nsets] for i in range(a.steps + 3)):
                nres += 1
                if nres == 3:
                    t0 = time.perf_counter()
            e2e_sec = time.perf_counter() - t0
            barrier()
            e2e_fill_sec = (time.perf_counter() - t_fill) * a.steps / (a.steps + 3)
            e2e_d2h = Bp * H * W
            assert nres == a.steps + 3
            # the upload alone (what PCIe allows for fp32 inputs)
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d1, d2 = torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])
            torch.cuda.synchronize()
            h0.record()
            for i in range(3):
                d1.copy_(hx[i % nsets][0], non_blocking=True)
                d2.copy_(hx[i % nsets][1], non_blocking=True)
            h1.record()
            torch.cuda.synchronize()
            h2d_ms = h0.elapsed_time(h1) / 3
            del d1, d2
        clocks = sampler.stop() if sampler else None
        # ---- per-launch profile (after the timed regions): roofline of the dominant kernel
        prof = None
        if rank == 0:
            runs = [net._engine.profile_pair(net, *sets[i % nsets]) for i in range(3)]
            prof = [dict(name=r["name"], flops=r["flops"], bytes=r["bytes"],
                         ms=statistics.mean(x[j]["ms"] for x in runs)) for j, r in enumerate(runs[0])]
        # ---- in-run parity of the timed mode against the strict fp32 mode (same weights, same inputs), and the
        #      strict mode's own throughput for reference
        parity, strict = None, None
        if rank == 0 and wl["variant"] == "levir":
            nb = min(8, Bp)
            xa, xb = sets[0][0][:nb].contiguous(), sets[0][1][:nb].contiguous()
            y_mode = net(xa, xb).double()
            saved = net._engine.flags
            net._engine.flags = 0
            net.invalidate_native_cache()
            y_ref = net(xa, xb).double()
            d = (y_mode - y_ref).abs()
            parity = dict(against="strict fp32 mode (flags 0), same weights and inputs, %d pairs" % nb,
                          max_abs=float(d.max()), mean_abs=float(d.mean()), ref_abs_max=float(y_ref.abs().max()),
                          outside_tol=int((d > 1e-4 + 1e-3 * y_ref.abs()).sum()), tol="1e-4 + 1e-3*|ref|", elements=d.numel(),
                          argmax_agree=float((y_mode.argmax(1) == y_ref.argmax(1)).float().mean()))
            # the other shipped modes on the same workload (5 steps each): throughput + their own parity against fp32
            strict = {}
            for other in ("fp32", "tf32", "f16"):
                if MODES[other] == saved:
                    continue
                net._engine.flags = MODES[other]
                net.invalidate_native_cache()
                y_o = net(xa, xb).double()
                for i in range(2):
                    step(i)
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s0.record()
                for i in range(5):
                    step(i)
                s1.record()
                torch.cuda.synchronize()
                do = (y_o - y_ref).abs()
                strict[other] = dict(value=Bp / (s0.elapsed_time(s1) / 5 / 1e3), unit="pairs/s per GPU",
                                     max_abs_vs_fp32=float(do.max()), outside_tol=int((do > 1e-4 + 1e-3 * y_ref.abs()).sum()),
                                     argmax_agree=float((y_o.argmax(1) == y_ref.argmax(1)).float().mean()))
            net._engine.flags = saved
            net.invalidate_native_cache()
    tmax = torch.tensor([ms, e2e_sec * 1e3, e2e_sync_sec * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_sync_ms = float(tmax[0]), float(tmax[1]), float(tmax[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = load_peaks()
    ms_step = ms / a.steps
    value = world * Bp / (ms_step / 1e3)
    e2e_value = world * Bp * a.steps / (e2e_ms / 1e3)
    top = max(prof, key=lambda r: r["ms"])
    ridge = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    ai = top["flops"] / max(top["bytes"], 1.0)
    if ai > ridge:
        achieved = top["flops"] / (top["ms"] * 1e-3) / 1e12
        roof = dict(kernel=top["name"], bound="tensor", achieved=achieved, peak=peaks["bf16_tflops_sustained"],
                    unit="TFLOP/s", frac=achieved / peaks["bf16_tflops_sustained"])
        # what the tensor pipe actually executes: kind::tf32 runs at half the bf16 rate, and 3xTF32 issues three MMAs
        # per algorithmic product
        fl = net._engine.flags                              # TF32-equivalent MMAs per product (a 16-bit K=16 MMA counts 1/2)
        mult = 0.5 if (fl & 1024) else ((1.5 if (fl & 2048) else 2 if (fl & 512) else 3) if (fl & 2) else 1)
        roof["tensor_pipe"] = dict(mma_tflops=achieved * mult, mmas_per_product=mult,
                                   tf32_peak_tflops=peaks["bf16_tflops_sustained"] / 2,
                                   frac=achieved * mult / (peaks["bf16_tflops_sustained"] / 2),
                                   note="kind::tf32 MMA throughput against half the measured bf16 rate")
    else:
        achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        roof = dict(kernel=top["name"], bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=achieved / peaks["hbm_gbs"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # per-launch DRAM bytes from the committed ncu capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(mode_name, {}).get(top["name"])
    roof.update(traffic=traffic, peak_source=f"MEASURED_PEAKS.json ({peak_src}; sustained bf16 for a kernel inside a step)",
                ai_flop_per_byte=ai, launch_ms=top["ms"], share_of_step=top["ms"] / sum(r["ms"] for r in prof),
                algorithmic_flops=top["flops"], algorithmic_bytes=top["bytes"],
                whole_step=dict(algorithmic_gflop=sum(r["flops"] for r in prof) / 1e9,
                                algorithmic_mb=sum(r["bytes"] for r in prof) / 1e6,
                                profiled_ms=sum(r["ms"] for r in prof)))
    kernels = sorted(prof, key=lambda r: -r["ms"])[:6]
    if a.dump_kernels:
        os.makedirs(os.path.dirname(os.path.abspath(a.dump_kernels)), exist_ok=True)
        json.dump(dict(flags=net._engine.flags, pairs=Bp, H=H, W=W, peaks=peaks, launches=prof), open(a.dump_kernels, "w"), indent=1)
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        cp = 4 if wl["variant"] == "levir" else 1
        v, sec = cpu_port_pairs_per_sec(wl, cp, 3, 1)
        cpu = dict(value=v, unit="pairs/s", cores=torch.get_num_threads(), kind="port",
                   sample=f"oracle port (torch CPU fp32), {cp} pairs per call, mean of 3 calls after 1 warm-up ({sec:.2f} s/call)")
    line = dict(metric="image-pairs/sec", value=value, unit="pairs/s", n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"fp32": "f32", "tf32": "tf32", "tf32_fast": "tf32", "tf32x3": "x3 error-compensated (convs and stem: f16 main + bf16 corrections; decoder: 3xTF32), fp32 accumulate and storage",
                       "tf32x3_unfolded": "x3 error-compensated (convs: f16 main + bf16 corrections; stem/decoder: 3xTF32), fp32 accumulate and storage",
                       "tf32x3_tf32main": "x3 error-compensated (tf32 main + bf16 corrections), fp32 accumulate and storage",
                       "tf32x3_pure": "3xTF32 (error-compensated), fp32 accumulate and storage", "bf16": "bf16 operands, fp32 accumulate and storage"}.get(mode_name, "f32/tf32"),
                data="synthetic",
                config=dict(workload=wl["desc"], H=H, W=W, pairs_per_gpu=Bp, global_pairs_per_step=world * Bp,
                            sharding="by image pair, one process per GPU, no collective",
                            weights="torch.manual_seed(0); define_G random init", mode=mode_name, flags=net._engine.flags,
                            l2="3 rotating input sets (3x%.0f MB) + multi-GB per-step intermediates >> 126 MB L2" % (2 * Bp * 3 * H * W * 4 / 1e6)),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit="pairs/s", h2d_bytes_per_step=2 * Bp * 3 * H * W * 4,
                         d2h_bytes_per_step=e2e_d2h, ms_per_step=e2e_ms / a.steps,
                         timing="steady state of one PairPipeline.run: K results between the 3rd and the (K+3)th hand-out (wall clock)",
                         ms_per_step_including_fill=e2e_fill_sec * 1e3 / a.steps, h2d_alone_ms_per_step=h2d_ms,
                         api="dahitra_b200.pipeline.PairPipeline(net).run(pinned host batches): H2D of batch i+1 overlaps the "
                             "forward of batch i; uint8 class map D2H every step",
                         unpipelined=dict(value=world * Bp * a.steps / (e2e_sync_ms / 1e3), d2h_bytes_per_step=Bp * H * W * 8,
                                          api="net(x1.to(dev), x2.to(dev)); torch.argmax(.,1) -> pinned host int64 map; per-step sync")),
                parity=parity, other_modes=strict,
                gpu_launches=len(prof) * a.steps, launches_per_step=len(prof),
                roofline=roof, top_kernels=[dict(name=k["name"], ms=round(k["ms"], 4)) for k in kernels])
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="levir256", choices=list(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU per step (default: the workload's)")
    ap.add_argument("--mode", default=DEFAULT_MODE, help="precision mode of the native engine: fp32 | tf32 | tf32_fast | tf32x3")
    ap.add_argument("--flags", type=int, default=None, help="raw DH_FLAG_* bitmask (overrides --mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-kernels", default=None, help="write the per-launch table (name, ms, flops, bytes) to this JSON file")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    if a.impl == "reference":
        return run_reference(a, wl)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus != world and world == 1 and a.gpus > 1:
        # convenience: re-launch under torchrun when asked for N>1 without it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_native(a, wl)


if __name__ == "__main__":
    sys.exit(main())
