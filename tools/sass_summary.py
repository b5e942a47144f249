"""Static evidence from the built library (no GPU needed): per kernel family, the number of instantiations, the register / shared /
local-memory (spill) figures of `cuobjdump --dump-resource-usage`, and the counts of the Blackwell-specific SASS mnemonics of
`cuobjdump -sass` (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UBLKCP = bulk copies, SYNCS = mbarrier operations, FFMA2 = packed fp32 FMA).      python tools/sass_summary.py > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dahitra_b200", "libdahitra_b200.so")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "FFMA2", "HFMA2", "FFMA", "MUFU")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def family(d):
    d = re.sub(r"^void\s+", "", d)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    return re.split(r"[<(]", d, 1)[0]


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(m.group(i)) for i in range(2, 6))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.defaultdict(collections.Counter), None
    for line in sass.split("\n"):
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for mn in MNEMONICS:
                if op == mn or (mn in ("UTCHMMA", "UTMALDG", "SYNCS", "UBLKCP", "UTCBAR", "LDTM", "STTM", "UTMASTG", "UTCQMMA") and op.startswith(mn)):
                    counts[cur][mn] += 1
    dm = demangle(sorted(set(usage) | set(counts)))
    fam = collections.defaultdict(list)
    for k in dm:
        fam[family(dm[k])].append(k)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(usage)} kernels (template instantiations) in {len(fam)} families; sm_100a SASS\n")
    cols = [m for m in MNEMONICS if any(counts[k][m] for k in counts)]
    print("| kernel family | instances | registers (min-max) | static shared (max, B) | stack / local (max, B) | SASS instructions | " + " | ".join(cols) + " |")
    print("|---|---:|---:|---:|---:|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for f in sorted(fam, key=lambda f: -sum(counts[k]["_all"] for k in fam[f])):
        ks = fam[f]
        regs = [usage[k][0] for k in ks if k in usage]
        sh = max((usage[k][2] for k in ks if k in usage), default=0)
        loc = max((max(usage[k][1], usage[k][3]) for k in ks if k in usage), default=0)
        row = [sum(counts[k][m] for k in ks) for m in cols]
        for m, v in zip(cols, row):
            tot[m] += v
        n = sum(counts[k]["_all"] for k in ks)
        tot["_all"] += n
        print(f"| `{f}` | {len(ks)} | {min(regs)}-{max(regs)} | {sh} | {loc} | {n} | " + " | ".join(str(v) if v else "" for v in row) + " |")
    print(f"| **total** | {len(usage)} | | | | {tot['_all']} | " + " | ".join(str(tot[m]) for m in cols) + " |")
    spilled = sorted((dm[k], usage[k][1], usage[k][3]) for k in usage if usage[k][1] or usage[k][3])
    print(f"\nKernels with a stack frame or local memory: {len(spilled)}")
    for d, st, lo in spilled[:40]:
        print(f"  stack {st} B, local {lo} B: {d[:160]}")


if __name__ == "__main__":
    sys.exit(main())
