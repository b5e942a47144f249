#!/bin/bash
# compute-sanitizer racecheck / synccheck / initcheck over one whole forward of the plain-C host (1 pair, 256x256, default mode)
mkdir -p gpurun_out
X=examples/bin
for tool in racecheck synccheck initcheck; do
  T=25; [ $tool = racecheck ] && T=40
  (timeout $T compute-sanitizer --tool $tool --error-exitcode 7 $X/dahitra_infer --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/y_$tool.bin; echo "$tool exit code $?") > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  grep -c "=========" gpurun_out/r02_sanitizer_$tool.txt; tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
