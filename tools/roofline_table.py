"""Per-launch roofline table from a bench.py --dump-kernels file.
   usage: python tools/roofline_table.py profiles/r02_kernels_tf32x3.json [mode_name] > profiles/r02_per_kernel_roofline_tf32x3.md
bound = tensor if the launch's algorithmic intensity is above the ridge (sustained bf16 peak / HBM peak), else hbm;
achieved = algorithmic FLOPs (or bytes) / measured launch time; peaks from the file (MEASURED_PEAKS.json at bench time).

The pixel-decoder rows are different: the kernel evaluates the COLLAPSED cross-attention + MLP (per layer four 128x32x32
products on the tensor cores, three TF32 passes each), ~9.5x fewer FLOPs than the reference's as-written q / k / v / out
projections, and it is bound by instruction issue (LayerNorm, softmax, GELU per pixel), not by the tensor pipe or HBM.  Its
row therefore reports the EXECUTED tensor-core FLOPs against the TF32 rate and, as the yardstick that bounds it, the HBM
fraction; the as-written FLOP count is listed only as the model-FLOP figure."""
import json, sys

d = json.load(open(sys.argv[1]))
mode = sys.argv[2] if len(sys.argv) > 2 else f"flags {d['flags']}"
pk = d["peaks"]
tf, hbm = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
ridge = tf * 1e12 / (hbm * 1e9)
L = d["launches"]
tot = sum(l["ms"] for l in L)
x3 = 3 if (d["flags"] & 32) else 1
print(f"# Per-launch roofline — mode `{mode}`, {d['pairs']} pairs of {d['H']}x{d['W']}, one B200\n")
print(f"Peaks: HBM {hbm:.0f} GB/s, bf16 dense {tf:.0f} TF/s sustained (ridge {ridge:.0f} FLOP/B).  Times: CUDA events around each launch "
      f"(`dahitra_forward_profiled`), sum {tot:.3f} ms (the un-profiled step is shorter: levels 4/3 overlap on side streams).\n")
print("| launch | ms | share | GFLOP | MB | AI (F/B) | bound | achieved | % of peak |")
print("|---|---:|---:|---:|---:|---:|---|---:|---:|")
dec_rows = []
for l in L:
    fl, by, ms = l["flops"], max(l["bytes"], 1.0), l["ms"]
    ai = fl / by
    if l["name"].startswith("decoder_"):
        # pixels x depth from the as-written count: per pixel per layer 2 * (32 I + 4 I + 4 I + I 32 + 2 * 32 * 32), I = 64 * heads
        heads = 8 if l["name"].startswith("decoder_3") else 4
        inner = 64 * heads
        per_px_layer = 2.0 * (32 * inner + 4 * inner + 4 * inner + inner * 32 + 2 * 32 * 32)
        px_layers = fl / per_px_layer
        executed = px_layers * 2.0 * (32 * 4 * heads + 4 * heads * 32 + 32 * 32 + 32 * 32) * x3      # four products, x3 passes
        ach = by / (ms * 1e-3) / 1e9
        dec_rows.append((l["name"], ms, fl, executed))
        print(f"| {l['name']} | {ms:.3f} | {100 * ms / tot:.1f} % | {executed / 1e9:.1f} executed ({fl / 1e9:.0f} as written) | {by / 1e6:.1f} | — | issue | "
              f"{ach:.0f} GB/s; {executed / (ms * 1e-3) / 1e12:.0f} TF/s of TF32 MMAs | {100 * ach / hbm:.1f} % of HBM; {100 * executed / (ms * 1e-3) / 1e12 / (tf / 2):.1f} % of the TF32 rate |")
        continue
    if ai > ridge:
        ach, peak, unit, b = fl / (ms * 1e-3) / 1e12, tf, "TF/s", "tensor"
    else:
        ach, peak, unit, b = by / (ms * 1e-3) / 1e9, hbm, "GB/s", "hbm"
    print(f"| {l['name']} | {ms:.3f} | {100 * ms / tot:.1f} % | {fl / 1e9:.1f} | {by / 1e6:.1f} | {ai:.0f} | {b} | {ach:.0f} {unit} | {100 * ach / peak:.1f} % |")
fl_all, by_all = sum(l["flops"] for l in L), sum(l["bytes"] for l in L)
print(f"\nWhole forward: {fl_all / 1e9:.0f} GFLOP as written, {by_all / 1e6:.0f} MB algorithmic; at the profiled {tot:.3f} ms: "
      f"{by_all / (tot * 1e-3) / 1e9:.0f} GB/s = {100 * by_all / (tot * 1e-3) / 1e9 / hbm:.1f} % of the HBM roof, "
      f"{fl_all / (tot * 1e-3) / 1e12:.0f} TF/s = {100 * fl_all / (tot * 1e-3) / 1e12 / tf:.1f} % of the bf16 rate.")
if dec_rows:
    print("\nDecoder rows: `executed` = tensor-core FLOPs the collapsed kernel issues (4 products of 128x32x32 per layer and 128-pixel tile, "
          f"x{x3} TF32 passes); the kernel is bound by instruction issue (ncu: issue-active ~57 %, tensor pipe ~25 %), see "
          "`r02_ncu_full_pixel_decoder_tc_level3.txt`.")
