#!/bin/bash
# The plain-C host of the C ABI on the GPU (examples/dahitra_infer.c; inputs pre-generated into examples/bin/ by the CPU side):
# plain run, its GPU tests, a 64-pair run from synthetic host images, and compute-sanitizer memcheck over one whole forward.
mkdir -p gpurun_out
X=examples/bin
(timeout 30 $X/dahitra_infer --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/y.bin --repeat 20) > gpurun_out/r02_c_host_1pair.log 2>&1
echo "c host rc=$?"; cat gpurun_out/r02_c_host_1pair.log
(timeout 80 python -m pytest tests/test_c_host.py -q -m gpu -s 2>&1 | tail -25) > gpurun_out/r02an_tests.log 2>&1
tail -12 gpurun_out/r02an_tests.log
(timeout 25 $X/dahitra_infer --weights $X/san_w.bin --synthetic 64x256x256 --repeat 20) > gpurun_out/r02_c_host_64pairs.log 2>&1
cat gpurun_out/r02_c_host_64pairs.log
(timeout 55 compute-sanitizer --tool memcheck --error-exitcode 7 $X/dahitra_infer --weights $X/san_w.bin --input $X/san_x.bin --output /tmp/y2.bin; echo "memcheck exit code $?") > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
tail -8 gpurun_out/r02_sanitizer_memcheck.txt
cmp /tmp/y.bin /tmp/y2.bin && echo "memcheck run wrote the same bytes as the plain run" >> gpurun_out/r02_sanitizer_memcheck.txt
