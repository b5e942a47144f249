"""Per-launch roofline table from a bench.py --dump-kernels file.
   usage: python tools/roofline_table.py profiles/r01_kernels_tf32x3.json [mode_name] > profiles/r01_per_kernel_roofline_tf32x3.md
bound = tensor if the launch's algorithmic intensity is above the ridge (sustained bf16 peak / HBM peak), else hbm;
achieved = algorithmic FLOPs (or bytes) / measured launch time; peaks from the file (MEASURED_PEAKS.json at bench time)."""
import json, sys

d = json.load(open(sys.argv[1]))
mode = sys.argv[2] if len(sys.argv) > 2 else f"flags {d['flags']}"
pk = d["peaks"]
tf, hbm = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
ridge = tf * 1e12 / (hbm * 1e9)
L = d["launches"]
tot = sum(l["ms"] for l in L)
print(f"# Per-launch roofline — mode `{mode}`, {d['pairs']} pairs of {d['H']}x{d['W']}, one B200\n")
print(f"Peaks: HBM {hbm:.0f} GB/s, bf16 dense {tf:.0f} TF/s sustained (ridge {ridge:.0f} FLOP/B).  Times: CUDA events around each launch "
      f"(`dahitra_forward_profiled`), sum {tot:.3f} ms (the un-profiled step is shorter: levels 4/3 overlap on side streams).\n")
print("| launch | ms | share | GFLOP | MB | AI (F/B) | bound | achieved | % of peak |")
print("|---|---:|---:|---:|---:|---:|---|---:|---:|")
for l in L:
    fl, by, ms = l["flops"], max(l["bytes"], 1.0), l["ms"]
    ai = fl / by
    if ai > ridge:
        ach, peak, unit, b = fl / (ms * 1e-3) / 1e12, tf, "TF/s", "tensor"
    else:
        ach, peak, unit, b = by / (ms * 1e-3) / 1e9, hbm, "GB/s", "hbm"
    print(f"| {l['name']} | {ms:.3f} | {100 * ms / tot:.1f} % | {fl / 1e9:.1f} | {by / 1e6:.1f} | {ai:.0f} | {b} | {ach:.0f} {unit} | {100 * ach / peak:.1f} % |")
