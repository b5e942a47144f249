"""Per-step kernel table from a torch.profiler text dump (tools/profile_train_step.py): python tools/profile_table.py file [steps]"""
import re, sys
rows = []
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for l in open(sys.argv[1]):
    parts = re.split(r"\s{2,}", l.strip())
    if len(parts) < 10:
        continue
    name, sc, calls = parts[0], parts[6], parts[-1]

    def us(s):
        for suf, f in (("ms", 1e3), ("us", 1.0), ("s", 1e6)):
            if s.endswith(suf):
                try:
                    return float(s[:-len(suf)]) * f
                except ValueError:
                    return None
        return None
    v = us(sc)
    if not v or name.startswith(("aten::", "autograd", "Optimizer", "_PixelDecoderTrain")) or "Backward" in name:
        continue
    rows.append((v / steps, name[:120], int(calls) // steps))
rows.sort(reverse=True)
print(f"device time per step (kernels listed): {sum(r[0] for r in rows) / 1e3:.2f} ms")
for t, n, c in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    print(f"{t:9.1f} us  {c:4d} calls  {n}")
