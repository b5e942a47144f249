#!/bin/bash
# compute-sanitizer memcheck over the other routes of the forward: every precision mode (strict fp32 on CUDA cores = conv_ffma /
# decoder.cu / tokens.cu; tf32 / f16 / bf16 = conv_tc2 on fp32 storage), an odd batch in the default mode, the xBD variant
mkdir -p gpurun_out
X=examples/bin
OUT=gpurun_out/r02_sanitizer_memcheck_modes.txt
: > $OUT
run() {  # label, timeout, args...
  local label=$1; local T=$2; shift; shift
  echo "##### $label: dahitra_infer $*" >> $OUT
  (timeout $T compute-sanitizer --tool memcheck --error-exitcode 7 $X/dahitra_infer "$@"; echo "exit code $?") 2>&1 | grep -v "^=========$" | tail -5 >> $OUT
}
run "fp32 (flags 0), 1 pair 256x256" 10 --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/a.bin --flags 0
run "tf32 (flags 61), 1 pair 256x256" 10 --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/b.bin --flags 61
run "f16 (flags 2109), 1 pair 256x256" 10 --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/c.bin --flags 2109
run "bf16 (flags 1085), 1 pair 256x256" 10 --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/d.bin --flags 1085
run "default mode, 3 pairs 256x320" 10 --weights $X/san_w.bin --synthetic 3x256x320
run "xBD variant, 5 classes, 1 pair 1024x1024 (default mode)" 20 --weights $X/san_w_xbd.bin --variant 1 --nc 5 --synthetic 1x1024x1024
cat $OUT
