#!/bin/bash
# Full-set ncu capture of ONE template instance of a kernel family: the slowest launch whose demangled name matches INSTANCE
# (a python regex) among the launches of FAMILY (the ncu -k regex on the function name) in `bench.py --steps 1`.
# Leaves text only in gpurun_out/ (raw metric table + details page; the .ncu-rep is deleted: gpurun copies back <= 64 MiB).
#   usage: tools/ncu_capture_instance.sh MODE TAG FAMILY INSTANCE OUTNAME [KEEP_REP]
MODE=${1:-tf32x3}; TAG=${2:-r02}; FAMILY=$3; INSTANCE=$4; OUT=$5; KEEP=${6:-0}
BENCH="python bench.py --steps 1 --warmup 3 --mode $MODE --no-cpu-baseline --no-parity"
mkdir -p gpurun_out
if [ ! -f gpurun_out/launches_${MODE}_$TAG.csv ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${MODE}_$TAG.csv \
      $BENCH > gpurun_out/ncu_launch_${MODE}_$TAG.log 2>&1
fi
SKIP=$(python - "$FAMILY" "$INSTANCE" gpurun_out/launches_${MODE}_$TAG.csv <<'PY'
import csv, re, sys
fam, inst, path = re.compile(sys.argv[1]), re.compile(sys.argv[2]), sys.argv[3]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
best, idx, n = -1.0, 0, 0
for r in rows[1:]:
    if fam.search(r[ki]):
        if inst.search(r[ki]):
            v = float(r[vi].replace(",", ""))
            if v > best: best, idx = v, n
        n += 1
print(idx)
PY
)
echo "[ncu_capture_instance] $INSTANCE: launch #$SKIP of family $FAMILY" | tee -a gpurun_out/ncu_capture_$TAG.log
REP=/tmp/prof_${OUT}_${MODE}_$TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$FAMILY -s $SKIP -c 1 -o $REP -f \
    $BENCH > gpurun_out/ncu_full_${OUT}_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_${OUT}_$TAG.log
python tools/ncu_summary.py $REP.ncu-rep > gpurun_out/${TAG}_ncu_full_${OUT}.txt 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_${OUT}.csv 2>/dev/null
ncu -i $REP.ncu-rep --page details > gpurun_out/${TAG}_ncu_details_${OUT}.txt 2>/dev/null
if [ "$KEEP" = "1" ]; then cp $REP.ncu-rep gpurun_out/; fi
