#!/bin/bash
# One gpurun call: (1) ncu launch list of `bench.py --mode $MODE` (gpu__time_duration per launch), (2) a full-set
# capture of the slowest launch of each kernel family named in $KERNELS (regexes, space separated).
#   usage: tools/ncu_capture.sh MODE TAG "regex1 regex2 ..."
# The launch order of bench.py is deterministic, so the index found in pass (1) addresses the same launch in (2).
MODE=${1:-tf32x3}
TAG=${2:-r01}
KERNELS=${3:-"conv_tc2_kernel pixel_decoder_tc_kernel"}
BENCH="python bench.py --steps 1 --warmup 3 --mode $MODE --no-cpu-baseline"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${MODE}_$TAG.csv \
    $BENCH > gpurun_out/ncu_launch_${MODE}_$TAG.log 2>&1
for K in $KERNELS; do
  SKIP=$(python - "$K" gpurun_out/launches_${MODE}_$TAG.csv <<'EOF'
import csv, re, sys
pat, path = re.compile(sys.argv[1]), sys.argv[2]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
best, idx, n = -1.0, 0, 0
for r in rows[1:]:
    if pat.search(r[ki]):
        v = float(r[vi].replace(",", ""))
        if v > best: best, idx = v, n
        n += 1
print(idx)
EOF
)
  echo "[ncu_capture] $K: slowest launch is #$SKIP of its family" | tee -a gpurun_out/ncu_capture_$TAG.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/prof_${K}_${MODE}_$TAG -f \
      $BENCH > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_${K}_$TAG.log
done
ls -la gpurun_out | grep $TAG
